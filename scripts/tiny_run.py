"""Tiny end-to-end run of every kernel (for compute-sanitizer): 3 messages, detect + both packers, random keys."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tfhe_omr_b200 as omr
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
from stage_times import random_detector

det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(1)
B = 3
a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
pv = det.detect((a, b))
rp = omr.RetrievalParams(300, 2)
pay = torch.randint(0, 256, (B, 612), dtype=torch.int16, device="cuda", generator=g)
w = torch.randint(0, 257, (rp.payload_cipher_count * 2, 300), dtype=torch.int16, device="cuda", generator=g)
i1 = det.encode_pertinent_indices(rp, pv, seed=3, n_cipher=2)
p1 = det.encode_pertinent_payloads(pv, pay, rp.combination_count, 2, w)
x = torch.randint(0, 134215681, (2, 1024), dtype=torch.int32, device="cuda", generator=g)
det.ntt(1, x); det.ntt(1, x, inverse=True)
torch.cuda.synchronize()
print("tiny run ok", int(pv.tensor.sum().item()) & 0xffff, int(i1.sum().item()) & 0xffff, int(p1.sum().item()) & 0xffff)

"""Soak: the whole D = 65 536 board (detect + both digests) N times on the same inputs; every repetition must reproduce the first
one bit for bit (pertinency vector checksum and the 33 digest ciphertexts).  A missed barrier or an exchange hazard in the
throughput kernels shows up here as a differing word.   python scripts/soak_board.py [N]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import tfhe_omr_b200 as omr
from stage_times import random_detector

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10
D = 65536
det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(7)
a = torch.randint(0, 2048, (D, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (D, 7), dtype=torch.int16, device="cuda", generator=g)
pay = torch.randint(0, 256, (D, 612), dtype=torch.int16, device="cuda", generator=g)
rp = omr.RetrievalParams(D, 50)
seed = bytes(range(32))
first = None
for it in range(N):
    t0 = time.perf_counter()
    pv = det.detect((a, b))
    idx = det.encode_pertinent_indices(rp, pv, seed=11, cipher_index=0, n_cipher=rp.max_encode_indices_cipher_count)
    dig = det.encode_pertinent_payloads(pv, pay, rp.combination_count, rp.cmb_count_per_cipher, seed=seed, all_payloads_count=D)
    torch.cuda.synchronize()
    # position-weighted checksum of the pertinency vector (wrapping int64 arithmetic) + the digests themselves
    w = torch.arange(1, 4097, device="cuda", dtype=torch.int64)
    chk = (pv.tensor.reshape(-1, 4096) * w).sum(dim=1)
    cur = (chk.clone(), idx.clone(), dig.clone())
    if first is None:
        first = cur
    ok = all(torch.equal(x, y) for x, y in zip(first, cur))
    print(f"board {it}: {time.perf_counter() - t0:.2f} s, {'identical' if ok else 'DIFFERENT'}", flush=True)
    if not ok:
        bad = (first[0] != cur[0]).nonzero().flatten()[:8].tolist()
        print("first differing messages:", bad)
        sys.exit(1)
print(f"soak ok: {N} boards identical")

"""Timing breakdown of encode_pertinent_payloads at D = 65 536 (weights from the seed, payload upload, kernel) on a synthetic
pertinency vector: used to tell first-call allocation cost from steady state."""
import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
import numpy as np, torch
import tfhe_omr_b200 as omr
from stage_times import random_detector
det = random_detector()
D = 65536
rp = omr.RetrievalParams(D, 50)
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randint(0, 2048, (2048, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (2048, 7), dtype=torch.int16, device="cuda", generator=g)
pv_small = det.detect((a, b))
# fake a full-size pertinency vector by repeating
t = pv_small.tensor.repeat(32, 1, 1).contiguous()

print(type(pv_small), t.shape)

payloads = np.random.default_rng(0).integers(0, 256, (D, 612), dtype=np.uint16)
seed = bytes(range(32))
def T(label, f, n=3):
    for i in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); print(f"{label} run {i}: {(time.perf_counter()-t0)*1e3:.2f} ms")
    return r
w = T("seeded_weights", lambda: det.seeded_weights(seed, rp.combination_count, rp.cmb_count_per_cipher, D))
pd = T("payload h2d", lambda: torch.from_numpy(payloads.view(np.int16)).to("cuda"))
class PV: pass
pvo = pv_small.__class__(t, 0) if True else None
T("encode (device payloads, device weights)", lambda: det.encode_pertinent_payloads(pvo, pd, rp.combination_count, rp.cmb_count_per_cipher, w))
T("encode (numpy payloads, seed)", lambda: det.encode_pertinent_payloads(pvo, payloads, rp.combination_count, rp.cmb_count_per_cipher, seed=seed, all_payloads_count=D))

"""One pass of every hot-path kernel at a given batch (random keys; the path is data-oblivious), for ncu:
detect (K1, sum7, K2, K3, K4) + index digest + payload digest.  No warm-up launches of the big kernels, so
`ncu -k regex:<kernel> -c 1` captures the one launch at the requested shape."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import tfhe_omr_b200 as omr
from stage_times import random_detector

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--board", type=int, default=65536)
ap.add_argument("--tensor-core-ks", action="store_true", help="opt into the CUTLASS int8 GEMM key switch (default: hand-written CUDA-core kernels)")
args = ap.parse_args()
det = random_detector()
if args.tensor_core_ks:
    det.set_tensor_core_key_switch(True)
B = args.batch
g = torch.Generator(device="cuda"); g.manual_seed(B)
a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
pay = torch.randint(0, 256, (B, 612), dtype=torch.int16, device="cuda", generator=g)
rp = omr.RetrievalParams(args.board, min(50, args.board))
w = torch.randint(0, 257, (rp.payload_cipher_count * rp.cmb_count_per_cipher, args.board), dtype=torch.int16, device="cuda", generator=g)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); s.record()
pv, t = det.detect_with_time_info((a, b))
i1 = det.encode_pertinent_indices(rp, pv, seed=3, n_cipher=rp.max_encode_indices_cipher_count)
p1 = det.encode_pertinent_payloads(pv, pay, rp.combination_count, rp.cmb_count_per_cipher, w)
e.record(); torch.cuda.synchronize()
print(f"profile driver ok: batch {B}, {s.elapsed_time(e):.1f} ms, l1 {t.total_first_level_bootstrapping_time:.1f} l2 {t.total_second_level_bootstrapping_time:.1f} "
      f"trace {t.total_trace_time:.1f} ms, checks {int(pv.tensor.sum().item()) & 0xffff} {int(i1.sum().item()) & 0xffff} {int(p1.sum().item()) & 0xffff}")

"""Key-switch time per batch size for the three shapes (split CUDA-core, CUDA-core, tensor-core GEMM): OMR_KS_GEMM_MIN=1 forces
the GEMM, OMR_KS_GEMM=0 disables it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector, timed
det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(1)
rl = torch.randint(0, 134215681, (16384, 2, 1024), dtype=torch.int32, device="cuda", generator=g)
row = []
for B in (1, 8, 32, 64, 128, 256, 512, 1024, 2048, 8192, 16384):
    det.key_switch(rl[:B]); torch.cuda.synchronize()
    row.append((B, round(min(timed(lambda: det.key_switch(rl[:B]))[0] for _ in range(3)), 3)))
print(os.environ.get("OMR_KS_GEMM", "1"), os.environ.get("OMR_KS_GEMM_MIN", "default"), row)

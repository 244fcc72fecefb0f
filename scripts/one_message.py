"""detect on ONE message with random keys (latency shapes): the target of the round-1 ncu capture of the small-batch kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector, timed
det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randint(0, 2048, (1, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (1, 7), dtype=torch.int16, device="cuda", generator=g)
det.detect((a, b)); torch.cuda.synchronize()
ms, pv = timed(lambda: det.detect((a, b)))
print(f"one message: {ms:.3f} ms, checksum {int(pv.tensor.sum().item()) & 0xffffff}")

"""A/B timing of the level-2 blind rotation (and level 1) for experimental builds: OMR_B200_LIB=<so> python scripts/ab_l2.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector, timed
det = random_detector()
B = 2368
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
lw = torch.randint(0, 4096, (B, 671), dtype=torch.int32, device="cuda", generator=g)
out = det.second_level_blind_rotate(lw); torch.cuda.synchronize()
t2 = [timed(lambda: det.second_level_blind_rotate(lw))[0] for _ in range(3)]
t1 = [timed(lambda: det.first_level_blind_rotate(a, b))[0] for _ in range(2)]
o1 = det.first_level_blind_rotate(a, b); torch.cuda.synchronize()
print(os.environ.get("OMR_B200_LIB", "default"), "l2", [round(x, 1) for x in t2], "l1", [round(x, 1) for x in t1], "checksum", int(out.sum().item()) & 0xffffff, int(o1.to(torch.int64).sum().item()) & 0xffffff)

"""Repeat the L2 blind rotation on the same inputs and print every duration (variance hunt)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector, timed
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(1)
lwe = torch.randint(0, 4096, (B, 671), dtype=torch.int32, device="cuda", generator=g)
a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
out = []
for i in range(8):
    t, _ = timed(lambda: det.second_level_blind_rotate(lwe)); out.append(round(t, 1))
print("l2 back-to-back:", out)
out = []
for i in range(4):
    t1, r = timed(lambda: det.first_level_blind_rotate(a, b))
    t2, _ = timed(lambda: det.second_level_blind_rotate(lwe)); out.append((round(t1, 1), round(t2, 1)))
print("l1 then l2:", out)

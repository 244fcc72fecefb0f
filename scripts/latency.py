"""Single-message and small-batch detect latency (BASELINE.json configs[0]); stage times and whole-call time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector, timed
det = random_detector()
g = torch.Generator(device="cuda"); g.manual_seed(1)
for B in (1, 2, 8, 16, 20, 21, 22, 24, 25, 64, 148, 149):
    a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
    b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
    det.detect((a, b)); torch.cuda.synchronize()
    t1, l1 = timed(lambda: det.first_level_blind_rotate(a, b))
    t2, ks = timed(lambda: det.key_switch(l1))
    t3, l2 = timed(lambda: det.second_level_blind_rotate(ks))
    t4, _ = timed(lambda: det.trace(l2))
    tt = min(timed(lambda: det.detect((a, b)))[0] for _ in range(3))
    print(f"B={B:4d}: l1 {t1:7.2f}  ks {t2:5.2f}  l2 {t3:7.2f}  trace {t4:5.2f}  detect {tt:7.2f} ms", flush=True)

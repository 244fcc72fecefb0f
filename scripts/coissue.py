"""Experiment: do the integer level-1 kernel and the FP64 level-2 kernel overlap when one CTA of each shares an SM?
Run with OMR_L1_HALF=1: 147 level-1 CTAs (256 threads, 4 blind rotations each) and 148 level-2 CTAs (256 threads), one wave
each, so that every SM holds one CTA of each kernel when both run."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from stage_times import random_detector
det = random_detector()
det.set_latency_shapes(False)
B1, B = 84, 148
g = torch.Generator(device="cuda"); g.manual_seed(1)
a = torch.randint(0, 2048, (B1, 512), dtype=torch.int16, device="cuda", generator=g)
b = torch.randint(0, 2048, (B1, 7), dtype=torch.int16, device="cuda", generator=g)
lw = torch.randint(0, 4096, (B, 671), dtype=torch.int32, device="cuda", generator=g)
sA, sB = torch.cuda.Stream(), torch.cuda.Stream()
def run(l1, l2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sA.wait_event(e0); sB.wait_event(e0)
    if l1:
        with torch.cuda.stream(sA): det.first_level_blind_rotate(a, b)
    if l2:
        with torch.cuda.stream(sB): det.second_level_blind_rotate(lw)
    torch.cuda.current_stream().wait_stream(sA); torch.cuda.current_stream().wait_stream(sB)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
run(True, True)
for _ in range(2):
    print(f"l1 alone {run(True, False):7.2f}  l2 alone {run(False, True):7.2f}  both {run(True, True):7.2f} ms", flush=True)

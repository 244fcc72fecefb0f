"""Per-stage device timing at a given batch (config 5 of BASELINE.json: bootstrap-only microbench).
Random keys/clues: the path is data-oblivious integer arithmetic."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tfhe_omr_b200 as omr
from tfhe_omr_b200.detector import BSK1_SHAPE, KSK_SHAPE, BSK2_SHAPE, TRACE_SHAPE

Q1, Q2 = 134215681, 1125899906826241


def random_detector(device=0, seed=0):
    g = torch.Generator(device=f"cuda:{device}"); g.manual_seed(seed)
    dev = f"cuda:{device}"
    bsk1 = torch.randint(0, Q1, BSK1_SHAPE, dtype=torch.int32, device=dev, generator=g)
    ksk = torch.randint(0, Q1, KSK_SHAPE, dtype=torch.int32, device=dev, generator=g)
    bsk2 = torch.randint(0, Q2, BSK2_SHAPE, dtype=torch.int64, device=dev, generator=g)
    trk = torch.randint(0, Q2, TRACE_SHAPE, dtype=torch.int64, device=dev, generator=g)
    return omr.Detector(omr.DetectionKey(bsk1, ksk, bsk2, trk), device=device)


def timed(fn, reps=1):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(reps):
        out = fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, nargs="+", default=[148, 592]); ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    det = random_detector()
    for B in args.batch:
        g = torch.Generator(device="cuda"); g.manual_seed(B)
        a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
        b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
        det.trace(det.second_level_blind_rotate(det.key_switch(det.first_level_blind_rotate(a, b)))); torch.cuda.synchronize()   # warm-up: scratch buffers of this batch size exist
        t1, l1 = timed(lambda: det.first_level_blind_rotate(a, b), args.reps)
        t2, ks = timed(lambda: det.key_switch(l1), args.reps)
        t3, l2 = timed(lambda: det.second_level_blind_rotate(ks), args.reps)
        t4, _ = timed(lambda: det.trace(l2), args.reps)
        tot = t1 + t2 + t3 + t4
        print(json.dumps({"batch": B, "l1_ms": round(t1, 2), "ks_ms": round(t2, 2), "l2_ms": round(t3, 2), "trace_ms": round(t4, 2),
                          "msgs_per_s": round(B / tot * 1e3, 1), "us_per_msg": round(tot / B * 1e3, 1)}), flush=True)

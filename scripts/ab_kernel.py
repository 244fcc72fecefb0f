"""A/B timing of one stage kernel across builds of the library: for each OMR_B200_LIB given on the command line (a '-' = the in-tree
build) a fresh process times the stage at the given batches, several repetitions each, and prints min / median device ms.
    python scripts/ab_kernel.py --stage l2 --batch 296 2368 16384 --reps 5 -- build_variants/before/libomr_before.so -"""
import argparse, json, os, statistics, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(args):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import torch
    from stage_times import random_detector
    det = random_detector()
    out = {}
    for B in args.batch:
        g = torch.Generator(device="cuda"); g.manual_seed(B)
        a = torch.randint(0, 2048, (B, 512), dtype=torch.int16, device="cuda", generator=g)
        b = torch.randint(0, 2048, (B, 7), dtype=torch.int16, device="cuda", generator=g)
        lw = torch.randint(0, 4096, (B, 671), dtype=torch.int32, device="cuda", generator=g)
        rl = torch.randint(0, 134215681, (B, 2, 1024), dtype=torch.int32, device="cuda", generator=g)
        r2 = torch.randint(0, 1125899906826241, (B, 2, 2048), dtype=torch.int64, device="cuda", generator=g)
        fn = {"l1": lambda: det.first_level_blind_rotate(a, b), "ks": lambda: det.key_switch(rl), "l2": lambda: det.second_level_blind_rotate(lw),
              "trace": lambda: det.trace(r2)}[args.stage]
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        out[B] = {"min": round(min(ts), 2), "median": round(statistics.median(ts), 2), "max": round(max(ts), 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="l2", choices=["l1", "ks", "l2", "trace"])
    ap.add_argument("--batch", type=int, nargs="+", default=[296, 2368, 16384])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("libs", nargs="*")
    args = ap.parse_args()
    if args.child:
        child(args); sys.exit(0)
    for rnd in range(args.rounds):
        for lib in args.libs:
            env = dict(os.environ)
            if lib != "-":
                env["OMR_B200_LIB"] = os.path.abspath(lib)
            else:
                env.pop("OMR_B200_LIB", None)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--stage", args.stage, "--reps", str(args.reps), "--batch"] +
                               [str(b) for b in args.batch], env=env, capture_output=True, text=True)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            print(f"round {rnd} {args.stage} {lib:45s} {line[-1] if line else r.stderr[-400:]}", flush=True)

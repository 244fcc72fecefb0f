#!/usr/bin/env python
"""omr_core/examples/omr.rs on the GPU: the end-to-end run of the reference's driver — key generation, clue generation for a
board with up to 50 planted messages, detection, index / payload digests, decode, verification — with the same stage timings the
reference logs (examples/omr.rs:125-137,159-172,179-208,215-220), through the product API only (no CPU oracle, no CPU fallback).

    python examples/omr.py --payload-count 65536          (-p; the reference's default board, README.md:99-125)
    python examples/omr.py -p 1                           (BASELINE.json configs[0]: single-message latency)

What differs from the Rust driver: there is no --thread-count (the batch is the parallelism), the detection key and the clues are
made on the GPU from a seeded ChaCha12 stream, and the recipient's secrets / clue public key — which never leave the recipient —
are sampled with numpy."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

Q2 = 1125899906826241


def recipient(np, seed):
    """SecretKeyPack::new + generate_clue_key (key_gen/secret.rs:46-106): binary s0, s2, ternary z1, z2; clue public key (pa, pa*s0 + e)"""
    rng = np.random.default_rng(seed)
    s0 = rng.integers(0, 2, 512, dtype=np.int32); z1 = rng.integers(-1, 2, 1024, dtype=np.int32)
    s2 = rng.integers(0, 2, 670, dtype=np.int32); z2 = rng.integers(-1, 2, 2048, dtype=np.int32)
    pa = rng.integers(0, 2048, 512, dtype=np.int64)
    full = np.convolve(pa, s0.astype(np.int64))
    prod = full[:512].copy(); prod[:511] -= full[512:]
    pb = (prod + np.rint(rng.normal(0.0, 0.8293, 512)).astype(np.int64)) % 2048
    return (s0, z1, s2, z2), (pa.astype(np.uint16), pb.astype(np.uint16))


def main(argv=None):
    import numpy as np
    import torch
    import tfhe_omr_b200 as omr
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-p", "--payload-count", type=int, default=1 << 16)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--no-warm-up", action="store_true",
                    help="skip the silent first pass at D = 8 that creates the CUDA context and loads the kernels (a cold process adds "
                         "0.3-0.9 s of one-time module loading to the first call of every stage; the reference binary has no such cost)")
    args = ap.parse_args(argv)
    if not args.no_warm_up and args.payload_count > 8:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            rc = main(["--payload-count", "8", "--device", str(args.device), "--seed", str(args.seed + 17), "--no-warm-up"])
        if rc:
            return rc
    D = max(1, args.payload_count)                           # the reference clamps to >= 1 (examples/omr.rs:47-65)
    pertinent_count = min(D, 50)                             # examples/omr.rs:103-107
    torch.cuda.set_device(args.device)
    rng = np.random.default_rng(args.seed)

    def stamp(label, t0, per=None):
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"{label}: {dt * 1e3:.3f} ms" + (f"  ({dt * 1e3 / per:.4f} ms each)" if per else ""))
        return dt

    t0 = time.perf_counter()
    secrets, clue_key = recipient(np, args.seed)             # the recipient and a decoy (examples/omr.rs:72-83)
    _, decoy_key = recipient(np, args.seed + 1)
    det = omr.Detector.generate(secrets, rng.bytes(32), device=args.device)
    stamp("key generation (recipient secrets on the host, detection key on the GPU)", t0)
    print(f"detection key size: {det.detect_key_size() / 2**20:.1f} MiB on the device")

    pertinent = np.zeros(D, bool); pertinent[:pertinent_count] = True; rng.shuffle(pertinent)      # examples/omr.rs:103-113
    planted = np.flatnonzero(pertinent)
    t0 = time.perf_counter()
    a, b = det.gen_clues(decoy_key, D, seed=rng.bytes(32))
    pa, pb = det.gen_clues(clue_key, D, seed=rng.bytes(32))
    sel = torch.from_numpy(planted).to(a.device)
    a[sel], b[sel] = pa[sel], pb[sel]
    stamp("gen clues time", t0, D)
    t0 = time.perf_counter()
    payloads = rng.integers(0, 256, (D, 612), dtype=np.uint16)                                     # Payload::random (payload.rs:26-38)
    stamp("gen payloads time", t0)

    t0 = time.perf_counter()
    pv, times = det.detect_with_time_info((a, b))
    dt = stamp("detect time", t0)
    print(f"detect time per message: {dt * 1e3 / D:.4f} ms   ({D / dt:.1f} messages/s; first level {times.total_first_level_bootstrapping_time:.1f} ms, "
          f"second level {times.total_second_level_bootstrapping_time:.1f} ms, trace {times.total_trace_time:.1f} ms on the device)")

    rp = omr.RetrievalParams(D, pertinent_count)
    n_idx = rp.max_encode_indices_cipher_count
    t0 = time.perf_counter()
    idx = det.encode_pertinent_indices(rp, pv, seed=int(rng.integers(0, 2**63)), cipher_index=0, n_cipher=n_idx)
    stamp("encode indices times", t0, n_idx)
    seed = rng.bytes(32)                                                                          # let seed = rng.gen() (examples/omr.rs:194)
    t0 = time.perf_counter()
    pay = det.encode_pertinent_payloads(pv, payloads, rp.combination_count, rp.cmb_count_per_cipher, seed=seed, all_payloads_count=D)
    stamp("encode pertinent payloads time", t0)

    z2 = secrets[3].astype(np.int64)
    z2n = torch.from_numpy(np.where(z2 < 0, Q2 + z2, z2).astype(np.uint64).view(np.int64)).to(a.device).reshape(1, 2048)
    det.ntt(2, z2n)                                                                               # NttRlweSecretKey::from_coeff_secret_key
    t0 = time.perf_counter()
    indices, solved = omr.Retriever(det, rp, z2n.reshape(-1)).decode_digest_host(idx, pay, seed=seed)
    stamp("decode time", t0)

    ok = indices == [int(p) for p in planted]
    for i, p in zip(indices, solved):                                                             # examples/omr.rs:222-232
        if not np.array_equal(payloads[i], p):
            ok = False
            print(f"Fail {i}\nDifferent count: {int((payloads[i] != p).sum())}")
    print("All done" if ok else "retrieval FAILED")
    det.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

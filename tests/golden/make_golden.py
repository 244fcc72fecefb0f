"""Generates tests/golden/golden_v1.npz from the CPU oracle (run in the build container: `python tests/golden/make_golden.py`).

The reference ships no golden vectors for this path and cannot be run here (Rust + un-vendored Primus-fhe), so these
fixtures pin the ORACLE's conventions against regressions and give the GPU tests an
oracle-independent target.  Inputs are integer-only (numpy PCG64 streams, uniformly random key blobs), so the
fixture does not depend on libm."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O  # noqa: E402

KEY_SEED, CLUE_SEED = 424242, 77


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


def inputs():
    rng = np.random.default_rng(CLUE_SEED)
    a = rng.integers(0, 2048, (2, 512), dtype=np.uint16); b = rng.integers(0, 2048, (2, 7), dtype=np.uint16)
    a[1, ::3] = 0                                    # exercises the a_i == 0 skip
    x1 = rng.integers(0, O.Q1, (3, 1024), dtype=np.uint32); x2 = rng.integers(0, O.Q2, (3, 2048), dtype=np.uint64)
    payloads = rng.integers(0, 256, (2, 612), dtype=np.uint16)
    weights = rng.integers(0, 257, (4, 300), dtype=np.uint16)
    return a, b, x1, x2, payloads, weights


def compute():
    a, b, x1, x2, payloads, weights = inputs()
    kp = O.KeyPack(blobs=O.random_key_blobs(KEY_SEED))
    out = {}
    l1 = kp.l1(a, b, threads=2); ks = kp.keyswitch(l1); l2 = kp.l2(ks, threads=2); tr = kp.trace(l2, threads=2)
    out["l1_sha"], out["l1_head"] = sha(l1), l1[:, :, :8].copy()
    out["ks"] = ks
    out["l2_sha"], out["l2_head"] = sha(l2), l2[:, :, :8].copy()
    out["pv_sha"], out["pv_head"] = sha(tr), tr[:, :, :8].copy()
    f1, f2 = x1.copy(), x2.copy()
    O.lib().orc_ntt1_forward(O.ptr(f1), 3); O.lib().orc_ntt2_forward(O.ptr(f2), 3)
    out["ntt1_sha"], out["ntt1_head"], out["ntt2_sha"], out["ntt2_head"] = sha(f1), f1[:, :8].copy(), sha(f2), f2[:, :8].copy()
    # digests over the two pertinency ciphertexts as messages 256, 257 of a D = 300 board (two base-257 digits)
    idx = np.stack([O.encode_indices(300, 2, tr, 256, 0xFEED, c) for c in range(2)])
    pay = O.encode_payloads(tr, payloads, 256, weights, 2)
    out["idx_sha"], out["idx_head"], out["pay_sha"], out["pay_head"] = sha(idx), idx[:, :, :8].copy(), sha(pay), pay[:, :, :8].copy()
    out["buckets"] = np.array([O.lib().orc_bucket_of(0xFEED, c, m, s) for c in range(2) for m in (0, 1, 256, 65535) for s in range(5)], np.uint32)
    out["chacha12_w"] = O.chacha12_weights(bytes(range(32)), 64)
    return out


# ---- v2: the counter-based (ChaCha12) generators — detection key and clues — for a fixed secret and seed -------------------
KG_SECRET_SEED, KG_SEED, CLUE_CB_SEED = 0x4F4D520003, bytes(range(50, 82)), bytes(range(90, 122))


def inputs_v2():
    kp = O.KeyPack(seed=KG_SECRET_SEED)
    msgs = np.random.default_rng(9).integers(0, 8, (3, 7), dtype=np.uint8)
    return kp, msgs


def compute_v2():
    kp, msgs = inputs_v2()
    ref = O.KeyPack(cb_from=kp, cb_seed=KG_SEED)
    out = {}
    for name, arr in (("bsk1", ref.bsk1), ("ksk", ref.ksk), ("bsk2", ref.bsk2), ("trace", ref.trk)):
        out[name + "_sha"] = sha(arr)
        out[name + "_head"] = np.ascontiguousarray(arr).reshape(-1)[:16].copy()
        out[name + "_tail"] = np.ascontiguousarray(arr).reshape(-1)[-16:].copy()
    a, b = kp.gen_clues_cb(CLUE_CB_SEED, 3, index0=70000, msgs=msgs)
    out["clue_a"], out["clue_b"] = a, b
    out["secrets"] = np.concatenate([x.astype(np.int32) for x in kp.secrets()])
    pa, pb = kp.clue_key()
    out["clue_key"] = np.stack([pa, pb])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **compute())
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **compute_v2())
    print("wrote golden_v1.npz, golden_v2.npz")

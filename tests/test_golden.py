"""Golden fixtures (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).
CPU: the oracle reproduces them.  GPU (-m gpu): the CUDA path reproduces them without the oracle in the loop."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

GOLD = dict(np.load(os.path.join(HERE, "golden", "golden_v1.npz")))


def test_oracle_reproduces_golden():
    out = G.compute()
    assert set(out) == set(GOLD)
    for k, v in out.items():
        assert np.array_equal(v, GOLD[k]), k


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    import torch
    import oracle as O
    import tfhe_omr_b200 as omr
    a, b, x1, x2, payloads, weights = G.inputs()
    bsk1, ksk, bsk2, trk = O.random_key_blobs(G.KEY_SEED)          # numpy PCG64 only; no oracle computation
    det = omr.Detector(omr.DetectionKey(bsk1, ksk, bsk2, trk), device=0)
    dev = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x).view(dt)).cuda()
    l1 = det.first_level_blind_rotate(dev(a, np.int16), dev(b, np.int16))
    ks = det.key_switch(l1)
    l2 = det.second_level_blind_rotate(ks)
    pv = det.trace(l2.clone())
    torch.cuda.synchronize()
    host = lambda t, dt: t.cpu().numpy().view(dt)
    assert np.array_equal(G.sha(host(l1, np.uint32)), GOLD["l1_sha"])
    assert np.array_equal(host(ks, np.uint32), GOLD["ks"])
    assert np.array_equal(G.sha(host(l2, np.uint64)), GOLD["l2_sha"])
    assert np.array_equal(G.sha(host(pv, np.uint64)), GOLD["pv_sha"])
    f1, f2 = dev(x1, np.int32), dev(x2, np.int64)
    det.ntt(1, f1); det.ntt(2, f2); torch.cuda.synchronize()
    assert np.array_equal(G.sha(host(f1, np.uint32)), GOLD["ntt1_sha"]) and np.array_equal(G.sha(host(f2, np.uint64)), GOLD["ntt2_sha"])
    rp = omr.RetrievalParams(300, 2)
    pvv = omr.PertinencyVector(pv, index0=256)
    idx = det.encode_pertinent_indices(rp, pvv, seed=0xFEED, cipher_index=0, n_cipher=2)
    pay = det.encode_pertinent_payloads(pvv, payloads, 4, 2, weights)
    torch.cuda.synchronize()
    assert np.array_equal(G.sha(host(idx, np.uint64)), GOLD["idx_sha"])
    assert np.array_equal(G.sha(host(pay, np.uint64)), GOLD["pay_sha"])
    # coefficient-domain key upload (OMR_KEYS_COEFF, the Rust-shim path): INTT the ring keys on the GPU, re-create
    kb1, kb2, kt = dev(bsk1, np.int32).reshape(-1, 1024), dev(bsk2, np.int64).reshape(-1, 2048), dev(trk, np.int64).reshape(-1, 2048)
    det.ntt(1, kb1, inverse=True); det.ntt(2, kb2, inverse=True); det.ntt(2, kt, inverse=True); torch.cuda.synchronize()
    det2 = omr.Detector(omr.DetectionKey(kb1, dev(ksk, np.int32), kb2, kt, coeff_domain=True), device=0)
    pv2 = det2.detect((dev(a, np.int16), dev(b, np.int16)))
    assert np.array_equal(G.sha(pv2.to_host()), GOLD["pv_sha"])


GOLD2 = dict(np.load(os.path.join(HERE, "golden", "golden_v2.npz")))


def test_oracle_reproduces_golden_v2():
    """the counter-based generators (ChaCha12 streams, integer Gaussian tables) are pinned against silent convention drift"""
    out = G.compute_v2()
    assert set(out) == set(GOLD2)
    for k, v in out.items():
        assert np.array_equal(v, GOLD2[k]), k


@pytest.mark.gpu
def test_gpu_reproduces_golden_v2():
    """GPU detection-key generation and clue generation reproduce the fixtures from the stored secrets / clue key alone (no oracle)"""
    import hashlib
    import tfhe_omr_b200 as omr
    sec = GOLD2["secrets"]
    secrets = (sec[:512], sec[512:1536], sec[1536:2206], sec[2206:])
    det = omr.Detector.generate(secrets, G.KG_SEED, device=0, want_keys=True)
    dk = det.detection_key
    for name, arr in (("bsk1", dk.bsk1), ("ksk", dk.ksk), ("bsk2", dk.bsk2), ("trace", dk.trace)):
        assert np.array_equal(np.frombuffer(hashlib.sha256(np.ascontiguousarray(arr).tobytes()).digest(), np.uint8), GOLD2[name + "_sha"]), name
        assert np.array_equal(arr.reshape(-1)[:16], GOLD2[name + "_head"]) and np.array_equal(arr.reshape(-1)[-16:], GOLD2[name + "_tail"])
    msgs = np.random.default_rng(9).integers(0, 8, (3, 7), dtype=np.uint8)
    a, b = det.gen_clues((GOLD2["clue_key"][0], GOLD2["clue_key"][1]), 3, seed=G.CLUE_CB_SEED, index0=70000, msgs=msgs)
    assert np.array_equal(a.cpu().numpy().view(np.uint16), GOLD2["clue_a"]) and np.array_equal(b.cpu().numpy().view(np.uint16), GOLD2["clue_b"])
    det.close()

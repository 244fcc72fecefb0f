"""Reference test vectors: if `tests/golden/ref_v1/*.omrb` exist — written by the reference's own CPU
path through `ffi/omr-b200-sys/examples/dump_vectors.rs` — the CUDA library (through the C ABI) and the CPU oracle must
reproduce every stage boundary, the pertinency vector and the payload digest word for word.  Until a machine with cargo has
produced them the two `test_reference_vectors_*` tests are skipped, and the same checker is exercised on vectors written in
the same format from the oracle, which proves the plumbing (coefficient-domain key upload, blobs, stage entry points) but says
nothing about Primus-fhe."""
import glob
import os

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "tests", "golden", "ref_v1")
HOWTO = ("no reference vectors: on a machine with cargo run `cargo run --release --example dump_vectors -- /tmp/ref_v1 4` in "
         "ffi/omr-b200-sys and copy /tmp/ref_v1/*.omrb to tests/golden/ref_v1/")
FILES = ("detection_key", "clues", "rlwe1", "lwe2", "rlwe2", "pertinency_vector", "payloads", "payload_digest")


def _load(d):
    from tfhe_omr_b200 import blobs
    out = {}
    for name in FILES:
        kind, arrs, hdr = blobs.load(os.path.join(d, name + ".omrb"))
        out[name] = (arrs, hdr)
    return out


def _ntt_native_keys(arrs, hdr):
    """detection-key arrays -> this library's NTT ordering (the oracle only takes that form)"""
    bsk1, ksk, bsk2, trk = (np.array(arrs[k]) for k in ("bsk1", "ksk", "bsk2", "trace"))
    if hdr["domain"] == 1:
        O.lib().orc_ntt1_forward(O.ptr(bsk1), bsk1.size // O.N1)
        O.lib().orc_ntt2_forward(O.ptr(bsk2), bsk2.size // O.N2)
        O.lib().orc_ntt2_forward(O.ptr(trk), trk.size // O.N2)
    return bsk1, ksk, bsk2, trk


def _coeff2(x):
    y = np.ascontiguousarray(x, np.uint64).copy().reshape(-1, O.N2)
    O.lib().orc_ntt2_inverse(O.ptr(y), y.shape[0])
    return y.reshape(np.shape(x))


def _digest_setup(v):
    (pay, _), (dig, dh) = v["payloads"], v["payload_digest"]
    n = pay["payloads"].shape[0]
    rp = O.retrieval_params(n, -(-n // 2))
    seed = bytes([int(dh["aux"]) & 0xFF] * 32)
    return n, rp, seed, np.array(pay["payloads"]), np.array(dig["ct"])


def check_oracle(d):
    v = _load(d)
    kp = O.KeyPack(blobs=_ntt_native_keys(*v["detection_key"]))
    a, b = np.array(v["clues"][0]["a"]), np.array(v["clues"][0]["b"])
    l1 = kp.l1(a, b)
    assert np.array_equal(l1, v["rlwe1"][0]["ct"]), "first-level blind rotations + sum (detector.rs:553-557)"
    ks = kp.keyswitch(l1)
    assert np.array_equal(ks, v["lwe2"][0]["ct"]), "key switch + modulus switch + offset (detector.rs:560-596)"
    l2 = kp.l2(ks)
    assert np.array_equal(l2, v["rlwe2"][0]["ct"]), "second-level blind rotation (detector.rs:599-624)"
    pv = kp.trace(l2)
    assert v["pertinency_vector"][1]["domain"] == 1
    assert np.array_equal(_coeff2(pv), v["pertinency_vector"][0]["pv"]), "N^-1, trace, to NTT (detector.rs:626-639)"
    n, rp, seed, payloads, want = _digest_setup(v)
    cc, n_cipher = rp["combination_count"], rp["payload_cipher_count"]
    w = np.zeros((n_cipher * 2, n), np.uint16)
    w[:cc] = O.chacha12_weights(seed, cc * n).reshape(cc, n)
    got = O.encode_payloads(pv, payloads, 0, w, n_cipher)
    assert np.array_equal(_coeff2(got), want), "encode_pertinent_payloads (detector.rs:341-453)"


def check_gpu(d):
    import torch
    import tfhe_omr_b200 as omr
    v = _load(d)
    det = omr.Detector.from_blob(os.path.join(d, "detection_key.omrb"), device=0)
    a, b = np.array(v["clues"][0]["a"]), np.array(v["clues"][0]["b"])
    da = torch.from_numpy(a.view(np.int16)).cuda(); db = torch.from_numpy(b.view(np.int16)).cuda()
    l1 = det.first_level_blind_rotate(da, db)
    assert np.array_equal(l1.cpu().numpy().view(np.uint32), v["rlwe1"][0]["ct"]), "first-level blind rotations + sum"
    ks = det.key_switch(l1)
    assert np.array_equal(ks.cpu().numpy().view(np.uint32), v["lwe2"][0]["ct"]), "key switch + modulus switch + offset"
    l2 = det.second_level_blind_rotate(ks)
    assert np.array_equal(l2.cpu().numpy().view(np.uint64), v["rlwe2"][0]["ct"]), "second-level blind rotation"
    det.set_output_domain(True)
    pv = det.detect_host(a, b, global_index0=0, want_pv=True)
    assert np.array_equal(pv, v["pertinency_vector"][0]["pv"]), "N^-1, trace, to NTT"
    n, rp, seed, payloads, want = _digest_setup(v)
    got = det.encode_payloads_seeded_host(payloads, seed, n, rp["combination_count"], 2)
    assert np.array_equal(got, want), "encode_pertinent_payloads"
    det.close()


def _write_oracle_vectors(d, n=3, coeff_keys=True):
    """vectors in the dump_vectors.rs format, but made by the ORACLE (plumbing check only)"""
    from tfhe_omr_b200 import blobs
    kp = O.KeyPack(seed=0x4F4D520001)
    decoy = O.KeyPack(seed=0x4F4D520002, sender_only=True)
    keys = [np.array(k) for k in (kp.bsk1, kp.ksk, kp.bsk2, kp.trk)]
    if coeff_keys:
        O.lib().orc_ntt1_inverse(O.ptr(keys[0]), keys[0].size // O.N1)
        O.lib().orc_ntt2_inverse(O.ptr(keys[2]), keys[2].size // O.N2)
        O.lib().orc_ntt2_inverse(O.ptr(keys[3]), keys[3].size // O.N2)
    blobs.dump(os.path.join(d, "detection_key.omrb"), "detection_key", dict(zip(("bsk1", "ksk", "bsk2", "trace"), keys)), domain=1 if coeff_keys else 0)
    a, b = decoy.gen_clues(11, n, threads=4)
    pa, pb = kp.gen_clues(12, n, threads=4)
    a[::2], b[::2] = pa[::2], pb[::2]
    blobs.dump(os.path.join(d, "clues.omrb"), "clues", {"a": a, "b": b}, count=n)
    l1 = kp.l1(a, b); ks = kp.keyswitch(l1); l2 = kp.l2(ks); pv = kp.trace(l2)
    blobs.dump(os.path.join(d, "rlwe1.omrb"), "rlwe1", {"ct": l1}, count=n, domain=1)
    blobs.dump(os.path.join(d, "lwe2.omrb"), "lwe2", {"ct": ks}, count=n)
    blobs.dump(os.path.join(d, "rlwe2.omrb"), "rlwe2", {"ct": l2}, count=n, domain=1)
    blobs.dump(os.path.join(d, "pertinency_vector.omrb"), "pertinency_vector", {"pv": _coeff2(pv)}, count=n, domain=1)
    payloads = np.random.default_rng(1).integers(0, 256, (n, O.PAYLOAD_LEN), dtype=np.uint16)
    blobs.dump(os.path.join(d, "payloads.omrb"), "payloads", {"payloads": payloads}, count=n)
    rp = O.retrieval_params(n, -(-n // 2)); cc, nc = rp["combination_count"], rp["payload_cipher_count"]
    w = np.zeros((nc * 2, n), np.uint16); w[:cc] = O.chacha12_weights(bytes([0x5A] * 32), cc * n).reshape(cc, n)
    dig = O.encode_payloads(pv, payloads, 0, w, nc)
    blobs.dump(os.path.join(d, "payload_digest.omrb"), "digest", {"ct": _coeff2(dig)}, count=nc, aux=0x5A, domain=1)


def _have_ref():
    return all(os.path.exists(os.path.join(REF_DIR, f + ".omrb")) for f in FILES)


def test_reference_vectors_oracle():
    if not _have_ref():
        pytest.skip(HOWTO)
    check_oracle(REF_DIR)


@pytest.mark.gpu
def test_reference_vectors_gpu():
    if not _have_ref():
        pytest.skip(HOWTO)
    check_gpu(REF_DIR)


def test_vector_checker_on_oracle_made_vectors(tmp_path):
    """the checker itself (blob format, coefficient-domain keys, stage order) on vectors the oracle wrote in the same format"""
    _write_oracle_vectors(str(tmp_path), n=2)
    check_oracle(str(tmp_path))
    raw = bytearray(open(tmp_path / "lwe2.omrb", "rb").read()); raw[70] ^= 1; open(tmp_path / "lwe2.omrb", "wb").write(raw)
    with pytest.raises(AssertionError, match="key switch"):
        check_oracle(str(tmp_path))                                    # a single flipped bit is caught, and attributed to its stage


@pytest.mark.gpu
def test_vector_checker_gpu_on_oracle_made_vectors(tmp_path):
    """the same through the C ABI: omr_ctx_create_from_blob with coefficient-domain keys, stage entry points, OMR_OUT_COEFF"""
    _write_oracle_vectors(str(tmp_path), n=3)
    check_gpu(str(tmp_path))

"""Oracle unit tests (CPU): constants and known-answer values the reference's parameters pin, the only
golden table of the reference (INV_MOD_257, matrix.rs:28-41), NTT vs schoolbook, gadget reconstruction, ChaCha."""
import numpy as np
import pytest

import oracle as O


def test_constants_kat():
    # SURVEY.md A.1 (derived from parameters/mod.rs:16-22,39-105; detector.rs:462-465,489-495; secret.rs:167-168)
    assert O.const("Q1") == 134215681 == 2**27 - 2047
    assert O.const("Q2") == 1125899906826241 == 2**50 - 16383
    assert O.const("PSI1") == 4073518 == pow(7, (O.Q1 - 1) // 2048, O.Q1)
    assert O.const("PSI2") == 765727830662934 == pow(11, (O.Q2 - 1) // 4096, O.Q2)
    assert pow(O.const("PSI1"), 1024, O.Q1) == O.Q1 - 1 and pow(O.const("PSI2"), 2048, O.Q2) == O.Q2 - 1
    assert O.const("N2_INV") == 1125350151012361 and (O.const("N2_INV") * 2048) % O.Q2 == 1
    assert O.const("LUT1_SCALE") == 4194240 == ((O.Q1 >> 4) + 1) >> 1
    assert O.const("LUT2_SCALE") == 4380933489596 == (2 * O.Q2 + 257) // (2 * 257)
    assert O.const("BSK1_ELEMS") * 4 == 32 * 2**20 and O.const("BSK2_ELEMS") * 8 == int(251.25 * 2**20)


def test_inv_mod_257_table():
    """the reference's INV_MOD_257 (matrix.rs:28-41) equals the Fermat inverses"""
    L = O.lib()
    assert L.orc_inv_mod_257(0) == 0
    for i in range(1, 257):
        assert L.orc_inv_mod_257(i) == pow(i, 255, 257)
        assert (L.orc_inv_mod_257(i) * i) % 257 == 1


def test_luts():
    """detector.rs:457-503 + lut.rs:12-27"""
    l1 = np.zeros(1024, np.uint32); l2 = np.zeros(2048, np.uint64)
    O.lib().orc_lut1(O.ptr(l1)); O.lib().orc_lut2(O.ptr(l2))
    s1 = 4194240
    assert (l1[:128] == s1).all() and (l1[128:896] == 0).all() and (l1[896:] == O.Q1 - s1).all()
    s2 = 4380933489596
    assert (l2[1728:1856] == s2).all() and l2[:1728].sum() == 0 and l2[1856:].sum() == 0


@pytest.mark.parametrize("level", [1, 2])
def test_ntt_vs_schoolbook_and_roundtrip(level):
    rng = np.random.default_rng(level)
    L = O.lib()
    if level == 1:
        n, q, dt, f, i, nc = 1024, O.Q1, np.uint32, L.orc_ntt1_forward, L.orc_ntt1_inverse, L.orc_negacyclic1
    else:
        n, q, dt, f, i, nc = 2048, O.Q2, np.uint64, L.orc_ntt2_forward, L.orc_ntt2_inverse, L.orc_negacyclic2
    a = rng.integers(0, q, n, dtype=dt); b = rng.integers(0, q, n, dtype=dt)
    c = np.zeros(n, dt); nc(O.ptr(a), O.ptr(b), O.ptr(c))
    fa, fb = a.copy(), b.copy(); f(O.ptr(fa), 1); f(O.ptr(fb), 1)
    assert (fa < q).all()
    prod = np.array([(int(x) * int(y)) % q for x, y in zip(fa, fb)], dtype=dt)
    i(O.ptr(prod), 1)
    assert np.array_equal(prod, c)
    back = fa.copy(); i(O.ptr(back), 1)
    assert np.array_equal(back, a)
    # evaluation-order convention: out[k] = a(psi^(2*brv(k)+1))  
    psi = O.const("PSI1" if level == 1 else "PSI2"); bits = n.bit_length() - 1
    mono = np.zeros(n, dt); mono[1] = 1; f(O.ptr(mono), 1)           # a(X) = X
    for k in (0, 1, 5, n - 1):
        brv = int(format(k, f"0{bits}b")[::-1], 2)
        assert int(mono[k]) == pow(psi, 2 * brv + 1, q)


@pytest.mark.parametrize("which,q,logb,levels,drop", [(0, O.Q1, 5, 4, 7), (1, O.Q1, 1, 27, 0), (2, O.Q2, 7, 6, 8), (3, O.Q2, 2, 25, 0)])
def test_gadget_reconstruction(which, q, logb, levels, drop):
    """SUM d_j B^j 2^drop = x + eps (mod q), |eps| <= 2^(drop-1); digits balanced, top absorbs."""
    rng = np.random.default_rng(which)
    xs = np.concatenate([rng.integers(0, q, 2000, dtype=np.uint64), np.array([0, 1, q - 1, q // 2, q // 2 + 1, q // 2 - 1], np.uint64)])
    d = np.zeros((len(xs), levels), np.int64)
    O.lib().orc_decompose(which, O.ptr(xs), len(xs), O.ptr(d))
    B = 1 << logb
    assert (d[:, :-1] >= -B // 2).all() and (d[:, :-1] < B // 2).all()
    assert (np.abs(d[:, -1]) <= B // 2 + 1).all()
    for x, row in zip(xs, d):
        rec = sum(int(v) << (drop + logb * j) for j, v in enumerate(row))
        v = int(x) if int(x) <= q // 2 else int(x) - q
        assert abs(rec - v) <= (1 << (drop - 1) if drop else 0)


def test_reduce128_special_form():
    rng = np.random.default_rng(9)
    for _ in range(2000):
        hi = int(rng.integers(0, 2**49)); lo = int(rng.integers(0, 2**63)) * 2 + int(rng.integers(0, 2))
        assert O.lib().orc_reduce128_q2(hi, lo) == ((hi << 64) | lo) % O.Q2 == O.lib().orc_mod128_q2(hi, lo)


CHACHA12_ZERO_KEY_BLOCK0_HEAD = "9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"


def test_chacha_core_against_chacha20_vector():
    """ChaCha20 block 0, zero key / nonce (the original ChaCha test vector, also RFC 7539 §2.3 structure);
    the weight stream uses the same core with 12 rounds (rand 0.8 StdRng)."""
    key = np.zeros(8, np.uint32); out = np.zeros(16, np.uint32)
    O.lib().orc_chacha_block(O.ptr(key), 0, 0, 20, O.ptr(out))
    assert out.tobytes()[:16].hex() == "76b8e0ada0f13d90405d6ae55386bd28"
    # the 12-round core the weight stream actually uses (rand 0.8 StdRng = ChaCha12): published zero-key / zero-nonce
    # test vector of ChaCha12 (and of ChaCha8 for good measure)
    O.lib().orc_chacha_block(O.ptr(key), 0, 0, 12, O.ptr(out))
    assert out.tobytes()[:32].hex() == CHACHA12_ZERO_KEY_BLOCK0_HEAD
    O.lib().orc_chacha_block(O.ptr(key), 0, 0, 8, O.ptr(out))
    assert out.tobytes()[:32].hex() == "3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
    # the first draws for the all-zero seed follow from that block: Uniform(0, 257) takes the high word of u32 * 257
    words = np.frombuffer(bytes.fromhex(CHACHA12_ZERO_KEY_BLOCK0_HEAD), dtype="<u4").astype(np.uint64)
    assert np.array_equal(O.chacha12_weights(bytes(32), 8), ((words * 257) >> 32).astype(np.uint16))
    w = O.chacha12_weights(bytes(range(32)), 5000)
    assert w.max() <= 256 and w.min() == 0 and abs(float(w.mean()) - 128) < 4
    assert np.array_equal(w[:100], O.chacha12_weights(bytes(range(32)), 100))


def test_retrieval_params_table():
    """SURVEY A.6 table (retrieval_params.rs:50-106 with the constants of secret.rs:196-203)."""
    exp = {1: (2, 260, 7, 3, 6, 3), 256: (2, 260, 7, 3, 55, 28), 4096: (3, 390, 5, 5, 55, 28), 65536: (3, 390, 5, 5, 55, 28)}
    for D, e in exp.items():
        rp = O.retrieval_params(D, min(D, 50))
        assert tuple(rp.values()) == e, (D, rp)


def test_bucket_hash_uniform():
    b = np.array([O.lib().orc_bucket_of(5, 1, m, 2) for m in range(20000)])
    assert b.min() == 0 and b.max() == 129
    assert np.bincount(b, minlength=130).min() > 100


def test_production_cmux_equals_plain_cmux():
    """the allocation-free / lazy / special-form CMux used for the CPU baseline computes exactly the plain one"""
    kp = O.KeyPack(blobs=O.random_key_blobs(5))
    rng = np.random.default_rng(8)
    for a in (1, 777, 1024, 2047):
        acc = rng.integers(0, O.Q1, (2, 1024), dtype=np.uint32)
        assert np.array_equal(kp.cmux1(acc, a, 3), kp.cmux1_simple(acc, a, 3))
    for a in (1, 1500, 2048, 4095):
        acc = rng.integers(0, O.Q2, (2, 2048), dtype=np.uint64)
        assert np.array_equal(kp.cmux2(acc, a, 5), kp.cmux2_simple(acc, a, 5))

"""GPU tests of the drop-in boundary itself: coefficient-domain outputs (OMR_OUT_COEFF), the Detector
accessors, argument validation of the host-buffer calls, thread safety and device hygiene."""
import threading

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _intt2(x):
    """oracle inverse NTT (scaled) of every polynomial of a u64 array [..., 2048]"""
    y = np.ascontiguousarray(x, np.uint64).copy().reshape(-1, O.N2)
    O.lib().orc_ntt2_inverse(O.ptr(y), y.shape[0])
    return y.reshape(x.shape)


@pytest.fixture(scope="module")
def small(keypack, decoy):
    rng = np.random.default_rng(77)
    a, b = decoy.gen_clues(31, 12, threads=8)
    pa, pb = keypack.gen_clues(32, 3, threads=8)
    planted = np.array([1, 5, 10])
    a[planted], b[planted] = pa, pb
    payloads = rng.integers(0, 256, (12, O.PAYLOAD_LEN), dtype=np.uint16)
    return planted, a, b, payloads


def test_coefficient_domain_outputs(detector, keypack, small):
    """OMR_OUT_COEFF: what the host-buffer calls return is the inverse transform of the NTT-native result (so a Rust shim can
    re-transform with Primus-fhe's own table, SURVEY §8b risk 1), and omr_decode_digest accepts the same form back."""
    import tfhe_omr_b200 as omr
    planted, a, b, payloads = small
    n, D = len(a), 300
    rp = omr.RetrievalParams(D, len(planted))
    seed = bytes(range(32))
    detector.pv_reset(); detector.set_output_domain(False)
    pv_ntt = detector.detect_host(a, b, global_index0=100, want_pv=True)
    idx_ntt = detector.encode_indices_host(rp, 11, 0, rp.max_encode_indices_cipher_count)
    pay_ntt = detector.encode_payloads_seeded_host(payloads, seed, D, rp.combination_count, 2)
    try:
        detector.pv_reset(); detector.set_output_domain(True)
        pv_c = detector.detect_host(a, b, global_index0=100, want_pv=True)
        idx_c = detector.encode_indices_host(rp, 11, 0, rp.max_encode_indices_cipher_count)
        pay_c = detector.encode_payloads_seeded_host(payloads, seed, D, rp.combination_count, 2)
        assert np.array_equal(pv_c, _intt2(pv_ntt))
        assert np.array_equal(idx_c, _intt2(idx_ntt)) and np.array_equal(pay_c, _intt2(pay_ntt))
        # a coefficient-domain pertinency ciphertext decrypts with the plain secret: b - a*z2 = Delta * [1,0,...] (omd.rs:45-58)
        s0, z1, s2, z2 = keypack.secrets()
        z2c = np.where(z2 < 0, O.Q2 + z2.astype(np.int64), z2.astype(np.int64)).astype(np.uint64)
        for m in (1, 0):
            az = np.zeros(O.N2, np.uint64)
            O.lib().orc_negacyclic2(O.ptr(np.ascontiguousarray(pv_c[m, 0])), O.ptr(z2c), O.ptr(az))
            ph = (pv_c[m, 1].astype(object) - az.astype(object)) % O.Q2
            dec = np.array([((2 * 257 * int(c) + O.Q2) // (2 * O.Q2)) % 257 for c in ph])
            assert dec[0] == (1 if m in planted else 0) and not dec[1:].any()
        # the recipient side in the same convention: coefficient-form secret and digests
        ret = omr.Retriever(detector, rp, z2c)
        found, solved = ret.decode_digest_host(idx_c, pay_c, seed=seed)
        assert found == [100 + int(p) for p in planted]
        assert np.array_equal(solved, payloads[planted])
    finally:
        detector.set_output_domain(False); detector.pv_reset()
    with pytest.raises(omr.OmrError):
        detector._ck(detector.L.omr_set_output_domain(detector.h, 7))


def test_lut_accessors_match_reference_layout(detector):
    """Detector::first_level_lut / second_level_lut (detector.rs:117-132, 457-503; lut.rs:12-27)"""
    l1, l2 = detector.first_level_lut(), detector.second_level_lut()
    r1 = np.zeros(O.N1, np.uint32); r2 = np.zeros(O.N2, np.uint64)
    O.lib().orc_lut1(O.ptr(r1)); O.lib().orc_lut2(O.ptr(r2))
    assert np.array_equal(l1, r1) and np.array_equal(l2, r2)
    s1 = ((O.Q1 >> 4) + 1) >> 1
    assert l1[0] == s1 and l1[127] == s1 and l1[128] == 0 and l1[1023] == O.Q1 - s1 and l1[895] == 0      # [s,0 x6,-s] in chunks of 128
    assert set(np.nonzero(l2)[0]) == set(range(1728, 1856)) and l2[1728] == (2 * O.Q2 + 257) // (2 * 257)


def test_weight_rows_are_validated_and_zero_padded(detector, small):
    """omr_encode_payloads takes the number of weight rows: the natural [combination_count][D] matrix (55 rows, odd) gives the
    same digest as the zero-padded 56-row one, and nothing is read past the caller's buffer."""
    import tfhe_omr_b200 as omr
    _, a, b, payloads = small
    D = 65536
    rp = omr.RetrievalParams(D, 50)
    assert rp.combination_count == 55 and rp.payload_cipher_count == 28
    w55 = np.random.default_rng(3).integers(0, 257, (55, D), dtype=np.uint16)
    w56 = np.concatenate([w55, np.zeros((1, D), np.uint16)])
    detector.pv_reset()
    detector.detect_host(a, b, global_index0=0)
    p55 = detector.encode_payloads_host(payloads, w55, 55, 2)
    p56 = detector.encode_payloads_host(payloads, w56, 55, 2)
    assert np.array_equal(p55, p56)
    with pytest.raises(omr.OmrError):
        detector.encode_payloads_host(payloads, w55[:54], 55, 2)
    with pytest.raises(omr.OmrError):
        detector.encode_payloads_host(payloads, np.concatenate([w56, w56[:1]]), 55, 2)
    rows = w55.ctypes.data
    st = detector.L.omr_encode_payloads(detector.h, payloads.ctypes.data, len(payloads), rows, 0, D, 28, 2, p55.ctypes.data)
    assert st != 0
    detector.pv_reset()


def test_host_buffer_calls_are_thread_safe(detector, small):
    """the header promises that host-buffer calls may come from any thread: concurrent encode calls on one context share
    one digest buffer, so each must hold the lock until its result is copied out"""
    import tfhe_omr_b200 as omr
    _, a, b, payloads = small
    rp = omr.RetrievalParams(300, 3)
    detector.pv_reset()
    detector.detect_host(a, b, global_index0=0)
    seeds = list(range(40, 52))
    want = {s: detector.encode_indices_host(rp, s, 0, rp.max_encode_indices_cipher_count) for s in seeds}
    got, errs = {}, []

    def work(s):
        try:
            for _ in range(3):
                got[s] = detector.encode_indices_host(rp, s, 0, rp.max_encode_indices_cipher_count)
        except Exception as e:             # noqa: BLE001
            errs.append(e)

    ths = [threading.Thread(target=work, args=(s,)) for s in seeds]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert not errs
    for s in seeds:
        assert np.array_equal(got[s], want[s]), s
    detector.pv_reset()


def test_calls_restore_the_current_device(detector, small):
    import torch
    _, a, b, _ = small
    before = torch.cuda.current_device()
    detector.pv_reset(); detector.detect_host(a[:1], b[:1]); detector.pv_reset()
    detector.detect((a[:1], b[:1]))
    assert torch.cuda.current_device() == before
    if torch.cuda.device_count() > 1:                              # a context on another GPU must not move torch's device
        import tfhe_omr_b200 as omr
        kp = O.random_key_blobs(1)
        d1 = omr.Detector(omr.DetectionKey(*kp), device=1)
        d1.detect_host(a[:1], b[:1])
        assert torch.cuda.current_device() == before
        d1.close()

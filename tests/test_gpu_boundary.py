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
        # the reference's signatures hand the pertinency vector back in (detector.rs:223-227): load it, pack again
        detector.pv_load(pv_c, global_index0=100)
        assert np.array_equal(detector.encode_indices_host(rp, 11, 0, rp.max_encode_indices_cipher_count), idx_c)
        assert np.array_equal(detector.encode_payloads_seeded_host(payloads, seed, D, rp.combination_count, 2), pay_c)
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


def test_stream_push_equals_one_shot(detector, keypack, decoy):
    """an ingest loop with arbitrary push sizes (crossing the 16 384-message staging chunk) builds the same
    resident digest, word for word, as detect + encode over the same messages in one shot — and it decodes."""
    import tfhe_omr_b200 as omr
    n, D, index0 = 16500, 20000, 1000
    rng = np.random.default_rng(8)
    a, b = decoy.gen_clues(41, n, threads=16)
    planted = np.sort(rng.choice(n, 5, replace=False))
    pa, pb = keypack.gen_clues(42, len(planted), threads=8)
    a[planted], b[planted] = pa, pb
    payloads = rng.integers(0, 256, (n, O.PAYLOAD_LEN), dtype=np.uint16)
    rp = omr.RetrievalParams(D, len(planted))
    seed = bytes(range(7, 39))
    detector.pv_reset()
    detector.detect_host(a, b, global_index0=index0)
    want = np.concatenate([detector.encode_indices_host(rp, 0x5EED, 0, rp.max_encode_indices_cipher_count),
                           detector.encode_payloads_seeded_host(payloads, seed, D, rp.combination_count, rp.cmb_count_per_cipher)])
    detector.pv_reset()
    detector.stream_begin(rp, 0x5EED, seed, global_index0=index0)
    got0, n0 = detector.stream_snapshot()
    assert n0 == 0 and not got0.any()
    lo = 0
    for sz in (1, 0, 7, 16385, 2, 105):
        detector.stream_push(a[lo:lo + sz], b[lo:lo + sz], payloads[lo:lo + sz]); lo += sz
    assert lo == n
    got, cnt = detector.stream_snapshot()
    assert cnt == n and np.array_equal(got, want)
    with pytest.raises(omr.OmrError):                                   # the board is full after D - index0 messages
        detector.stream_push(a[:D - index0 - n + 1], b[:D - index0 - n + 1], payloads[:D - index0 - n + 1])
    detector.stream_end()
    with pytest.raises(omr.OmrError):
        detector.stream_push(a[:1], b[:1], payloads[:1])
    s0, z1, s2, z2 = keypack.secrets()
    z2n = np.where(z2 < 0, O.Q2 + z2.astype(np.int64), z2.astype(np.int64)).astype(np.uint64)
    O.lib().orc_ntt2_forward(O.ptr(z2n), 1)
    n_idx = rp.max_encode_indices_cipher_count
    found, solved = omr.Retriever(detector, rp, z2n).decode_digest_host(got[:n_idx], got[n_idx:], seed=seed)
    assert found == [index0 + int(p) for p in planted] and np.array_equal(solved, payloads[planted])


def test_digest_allreduce_single_rank(detector):
    """K7 through the C ABI with the library's own communicator (one rank: the sum is the identity, the mod-q2 pass is not)"""
    import torch
    import tfhe_omr_b200 as omr
    try:
        uid = detector.comm_unique_id()
    except omr.OmrError as e:
        pytest.skip(f"libnccl not loadable here: {e}")
    detector.comm_init(1, 0, uid)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    x = torch.randint(0, 2**62, (33, 2, 2048), dtype=torch.int64, device="cuda", generator=g)
    want = (x.cpu().numpy().view(np.uint64) % np.uint64(O.Q2))
    detector.digest_allreduce(x); torch.cuda.synchronize()
    assert np.array_equal(x.cpu().numpy().view(np.uint64), want)
    detector.comm_destroy()
    with pytest.raises(omr.OmrError):
        detector.digest_allreduce(x)


def test_gpu_detection_key_generation(keypack, decoy):
    """SURVEY §8f.4 / secret.rs:118-178: the detection key made on the GPU from the recipient's secrets and a 32-byte seed is,
    word for word, the oracle's counter-based key (BSK1, KSK, BSK2, trace key), and a detector made that way passes the
    reference's omd assertions (omd.rs:45-58) on clues of the same recipient."""
    import tfhe_omr_b200 as omr
    seed = bytes(range(200, 232))
    det = omr.Detector.generate(keypack.secrets(), seed, device=0, want_keys=True)
    ref = O.KeyPack(cb_from=keypack, cb_seed=seed)
    dk = det.detection_key
    for name, got, want in (("bsk1", dk.bsk1, ref.bsk1), ("ksk", dk.ksk, ref.ksk), ("bsk2", dk.bsk2, ref.bsk2), ("trace", dk.trace, ref.trk)):
        assert got.shape == want.shape and np.array_equal(got, want), name
    assert dk.bsk1.max() < O.Q1 and dk.ksk.max() < O.Q1 and dk.bsk2.max() < O.Q2 and dk.trace.max() < O.Q2
    a0, b0 = keypack.gen_clues(901, 2)
    a1, b1 = decoy.gen_clues(902, 2)
    a, b = np.concatenate([a0, a1]), np.concatenate([b0, b1])
    pv = det.detect((a, b)).to_host()
    assert np.array_equal(pv, ref.detect(a, b, threads=4))                 # same key on both sides -> same ciphertexts
    dec = [keypack.decrypt_decode(pv[i]) for i in range(4)]
    assert all(d[0] == 1 and not d[1:].any() for d in dec[:2]) and all(not d.any() for d in dec[2:])
    with pytest.raises(omr.OmrError):                                      # secrets are validated
        s0, z1, s2, z2 = keypack.secrets()
        omr.Detector.generate((s0 + 2, z1, s2, z2), seed, device=0)
    det.close()


def test_omr_example_end_to_end(capsys):
    """examples/omr.py — the reference's examples/omr.rs driver on the GPU — at a small board, product API only"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("omr_example", os.path.join(root, "examples", "omr.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    assert mod.main(["-p", "300"]) == 0
    out = capsys.readouterr().out
    assert "All done" in out and "detect time per message" in out and "decode time" in out
    assert mod.main(["-p", "1"]) == 0                        # BASELINE.json configs[0]: one message, three index ciphertexts

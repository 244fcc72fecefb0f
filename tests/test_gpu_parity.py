"""GPU parity tests: every stage of the CUDA path against the CPU oracle on the same inputs, bit-exact, through the
C ABI (libomr_b200.so).  Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


def _dev(x, dtype):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x).view(dtype)).cuda()


def _mixed_clues(keypack, decoy, n, pertinent_idx, seed=11):
    a, b = decoy.gen_clues(seed, n)
    for i in pertinent_idx:
        ai, bi = keypack.gen_clues(seed + 1, 1, index0=i)
        a[i], b[i] = ai[0], bi[0]
    return a, b


def test_ntt_forward_inverse_vs_oracle(detector):
    import torch
    rng = np.random.default_rng(1)
    x1 = rng.integers(0, O.Q1, (5, O.N1), dtype=np.uint32)
    x2 = rng.integers(0, O.Q2, (5, O.N2), dtype=np.uint64)
    for level, x, dt, fwd, inv in ((1, x1, np.int32, "orc_ntt1_forward", "orc_ntt1_inverse"), (2, x2, np.int64, "orc_ntt2_forward", "orc_ntt2_inverse")):
        ref = x.copy(); getattr(O.lib(), fwd)(O.ptr(ref), ref.shape[0])
        d = _dev(x, dt)
        detector.ntt(level, d); torch.cuda.synchronize()
        got = d.cpu().numpy().view(x.dtype)
        assert np.array_equal(got, ref)
        detector.ntt(level, d, inverse=True); torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(x.dtype), x)


@pytest.fixture(params=["latency", "throughput"])
def shape(request, detector):
    """Small batches run the latency launch shapes by default; every stage is checked in both (omr_set_latency_shapes)."""
    detector.set_latency_shapes(request.param == "latency")
    yield request.param
    detector.set_latency_shapes(True)


def test_stages_bit_exact(detector, keypack, decoy, shape):
    import torch
    a, b = _mixed_clues(keypack, decoy, 3, [1])
    da, db = _dev(a, np.int16), _dev(b, np.int16)
    l1 = detector.first_level_blind_rotate(da, db); torch.cuda.synchronize()
    ref_l1 = keypack.l1(a, b)
    assert np.array_equal(l1.cpu().numpy().view(np.uint32), ref_l1)
    ks = detector.key_switch(l1); torch.cuda.synchronize()
    ref_ks = keypack.keyswitch(ref_l1)
    assert np.array_equal(ks.cpu().numpy().view(np.uint32), ref_ks)
    l2 = detector.second_level_blind_rotate(ks); torch.cuda.synchronize()
    ref_l2 = keypack.l2(ref_ks)
    assert np.array_equal(l2.cpu().numpy().view(np.uint64), ref_l2)
    tr = detector.trace(l2.clone()); torch.cuda.synchronize()
    ref_tr = keypack.trace(ref_l2)
    assert np.array_equal(tr.cpu().numpy().view(np.uint64), ref_tr)
    # whole pipeline in one call
    pv = detector.detect((a, b))
    assert np.array_equal(pv.to_host(), ref_tr)


def test_launch_shapes_agree(detector, keypack):
    """Latency and throughput shapes of every stage give identical words on random (not clue-shaped) inputs, including
    batch sizes around the switch-over points (level 1: one rotation per CTA in one or several waves / four / six per CTA;
    level 2: clusters up to ~44 messages, 512-thread CTAs, 256-thread CTAs; key switch: split rows up to 256 messages)."""
    import torch
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2048, (200, 512), dtype=np.uint16); b = rng.integers(0, 2048, (200, 7), dtype=np.uint16)
    rl = rng.integers(0, O.Q1, (256, 2, O.N1), dtype=np.uint32)
    lw = rng.integers(0, 4096, (300, 671), dtype=np.uint32)
    got = {}
    assert detector.key_switch_path() == "cuda-core"        # compare the (default) CUDA-core key-switch shapes with each other
    for lat in (True, False):
        detector.set_latency_shapes(lat)
        l1 = [detector.first_level_blind_rotate(_dev(a[:n], np.int16), _dev(b[:n], np.int16)) for n in (21, 22, 43, 100, 200)]
        ks = [detector.key_switch(_dev(rl[:n], np.int32)) for n in (1, 17, 256)]
        l2 = [detector.second_level_blind_rotate(_dev(lw[:n], np.int32)) for n in (1, 24, 47, 148, 280, 300)]
        torch.cuda.synchronize()
        got[lat] = [x.cpu().numpy() for x in l1 + ks + l2]
    detector.set_latency_shapes(True)
    for x, y in zip(got[True], got[False]):
        assert np.array_equal(x, y)
    # key switch of one message against the oracle on a random (full-range) ciphertext
    first_ks = 5                                                # position of key_switch(rl[:1]) after the five level-1 outputs
    assert np.array_equal(got[True][first_ks].view(np.uint32).reshape(1, -1)[:, :671], keypack.keyswitch(rl[:1]))


def test_small_batch_exchanges_repeatable(detector):
    """The small-batch kernels exchange partial sums between groups / CTAs (shared memory in level 1, a double-buffered
    global scratch around a cluster barrier in level 2, integer atomics in the key switch): 12 repetitions at one and two
    waves of clusters must reproduce the throughput kernels' words every time (compute-sanitizer is not available on the
    pool; a missed barrier shows up here as a flaky word)."""
    import torch
    rng = np.random.default_rng(9)
    a = rng.integers(0, 2048, (3, 512), dtype=np.uint16); b = rng.integers(0, 2048, (3, 7), dtype=np.uint16)
    lw = rng.integers(0, 4096, (44, 671), dtype=np.uint32)
    rl = rng.integers(0, O.Q1, (40, 2, O.N1), dtype=np.uint32)
    da, db, dlw, drl = _dev(a, np.int16), _dev(b, np.int16), _dev(lw, np.int32), _dev(rl, np.int32)
    assert detector.key_switch_path() == "cuda-core"        # the CUDA-core key switch with its integer atomics is the one under test
    detector.set_latency_shapes(False)
    want = [detector.first_level_blind_rotate(da, db), detector.second_level_blind_rotate(dlw[:22]), detector.second_level_blind_rotate(dlw),
            detector.key_switch(drl)]
    torch.cuda.synchronize()
    detector.set_latency_shapes(True)
    for _ in range(12):
        got = [detector.first_level_blind_rotate(da, db), detector.second_level_blind_rotate(dlw[:22]), detector.second_level_blind_rotate(dlw),
               detector.key_switch(drl)]
        torch.cuda.synchronize()
        for x, y in zip(got, want):
            assert torch.equal(x, y)


def test_extreme_inputs_bit_exact(detector, keypack, shape):
    """Degenerate and saturated inputs through every stage: all-zero clues (every monomial is X^0: the skip path of the
    level-2 kernels, the not-skipped identity CMux of level 1), all-maximal values, a single non-zero coefficient, and
    non-canonical words (bits above the modulus are ignored, as `CmLwe<u16>` / `Lwe<u32>` values are canonical upstream)."""
    import torch
    a = np.zeros((4, 512), np.uint16); b = np.zeros((4, 7), np.uint16)
    a[1] = 2047; b[1] = 2047
    a[2, 0] = 1; b[2, 3] = 1024
    a[3, ::2] = 2047; a[3, 1::2] = 1; b[3] = np.arange(7) * 293
    ref_l1 = keypack.l1(a, b)
    l1 = detector.first_level_blind_rotate(_dev(a, np.int16), _dev(b, np.int16)); torch.cuda.synchronize()
    assert np.array_equal(l1.cpu().numpy().view(np.uint32), ref_l1)
    noisy = detector.first_level_blind_rotate(_dev(a | 0xF800, np.int16), _dev(b | 0x8000, np.int16)); torch.cuda.synchronize()
    assert np.array_equal(noisy.cpu().numpy(), l1.cpu().numpy())
    lw = np.zeros((4, 671), np.uint32)
    lw[1] = 4095
    lw[2, 669] = 2048; lw[2, 670] = 4095
    lw[3, ::3] = 1; lw[3, 670] = 2047
    ref_l2 = keypack.l2(lw)
    l2 = detector.second_level_blind_rotate(_dev(lw, np.int32)); torch.cuda.synchronize()
    assert np.array_equal(l2.cpu().numpy().view(np.uint64), ref_l2)
    l2n = detector.second_level_blind_rotate(_dev(lw | 0xFFFFF000, np.int32)); torch.cuda.synchronize()
    assert np.array_equal(l2n.cpu().numpy(), l2.cpu().numpy())
    # key switch of saturated accumulators (every coefficient q1 - 1, then 0) and the trace of the level-2 outputs
    rl = np.full((2, 2, O.N1), O.Q1 - 1, np.uint32); rl[1] = 0
    ks = detector.key_switch(_dev(rl, np.int32)); torch.cuda.synchronize()
    assert np.array_equal(ks.cpu().numpy().view(np.uint32), keypack.keyswitch(rl))
    tr = detector.trace(l2.clone()); torch.cuda.synchronize()
    assert np.array_equal(tr.cpu().numpy().view(np.uint64), keypack.trace(ref_l2))


def test_weights_from_seed_match_reference_stream(detector):
    """detector.rs:376-387 / retriever.rs:215-226: the combination weights are StdRng::from_seed(seed) + Uniform(0, 257); the
    GPU stream equals the oracle's restatement of rand 0.8 / rand_chacha 0.3 for ragged lengths around the 16-word ChaCha
    block, for a board-sized matrix, and on the strictly in-order path that a rejected draw would take."""
    import torch
    for seed, rows, cols in ((bytes(range(32)), 1, 1), (bytes(range(32)), 1, 15), (bytes(32), 1, 17), (bytes([7] * 32), 3, 1000),
                             (bytes(range(1, 33)), 55, 4096)):
        ref = O.chacha12_weights(seed, rows * cols).reshape(rows, cols)
        got = detector.weights_from_seed(seed, rows, cols); torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().view(np.uint16), ref)
        assert ref.max() < 257
    # without the oracle: the all-zero seed's first draws follow from the published ChaCha12 zero-key block
    words = np.frombuffer(bytes.fromhex("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"), dtype="<u4").astype(np.uint64)
    got0 = detector.weights_from_seed(bytes(32), 1, 8); torch.cuda.synchronize()
    assert np.array_equal(got0.cpu().numpy().view(np.uint16)[0], ((words * 257) >> 32).astype(np.uint16))
    ordered = detector.weights_from_seed(bytes([7] * 32), 3, 1000, in_order=True); torch.cuda.synchronize()
    assert np.array_equal(ordered.cpu().numpy().view(np.uint16), O.chacha12_weights(bytes([7] * 32), 3000).reshape(3, 1000))
    with pytest.raises(Exception):
        detector.weights_from_seed(b"short", 1, 1)


def test_tensor_core_key_switch_is_exact(detector, keypack):
    """The opt-in tensor-core key switch (digits x key limbs as an int8 GEMM, int32 accumulation, limbs recombined mod q1):
    its words must equal the default CUDA-core kernels' (split rows for <= 256 messages, one CTA per 16 messages above) on
    full-range random ciphertexts at ragged batch sizes, and the oracle's on a sample — including saturated rows."""
    import torch
    rng = np.random.default_rng(21)
    rl = rng.integers(0, O.Q1, (1030, 2, O.N1), dtype=np.uint32)
    rl[5] = O.Q1 - 1; rl[6] = 0; rl[7, 0] = (O.Q1 - 1) // 2; rl[8, 0] = (O.Q1 + 1) // 2      # extreme digits: all -1 / 0 / +-max
    d = _dev(rl, np.int32)
    sizes = (1, 9, 256, 257, 1030)
    assert detector.key_switch_path() == "cuda-core"                                        # the hand-written kernels are the default
    cc = [detector.key_switch(d[:n]) for n in sizes]; torch.cuda.synchronize()
    try:
        detector.set_tensor_core_key_switch(True)                                           # opt-in CUTLASS int8 GEMM
        assert detector.key_switch_path() == "tensor-core"
        tc = [detector.key_switch(d[:n]) for n in sizes]; torch.cuda.synchronize()
    finally:
        detector.set_tensor_core_key_switch(False)
    for x, y in zip(tc, cc):
        assert torch.equal(x, y)
    sample = np.array([0, 5, 6, 7, 8, 511, 1023, 1029])
    assert np.array_equal(tc[-1].cpu().numpy().view(np.uint32)[sample], keypack.keyswitch(rl[sample]))


def test_omd_acceptance(detector, keypack, decoy):
    """omr_core/examples/omd.rs:45-58: pertinent -> [1,0,...,0], non-pertinent -> all 0."""
    a, b = _mixed_clues(keypack, decoy, 2, [0], seed=21)
    pv = detector.detect((a, b)).to_host()
    d0, d1 = keypack.decrypt_decode(pv[0]), keypack.decrypt_decode(pv[1])
    assert d0[0] == 1 and not d0[1:].any()
    assert not d1.any()


def test_digest_bit_exact_and_decode(detector, keypack, decoy):
    """examples/omr.rs / omr_time_analyze2.rs:220-240 at D = 48: digest bit-exact vs oracle packing of the same
    pertinency vector, and the decoded set / payloads equal the planted ones."""
    import tfhe_omr_b200 as omr
    D, pert = 48, [3, 17, 40]
    a, b = _mixed_clues(keypack, decoy, D, pert, seed=31)
    pv = detector.detect((a, b))
    pvh = pv.to_host()
    rng = np.random.default_rng(5)
    payloads = rng.integers(0, 256, (D, O.PAYLOAD_LEN), dtype=np.uint16)
    rp = omr.RetrievalParams(D, len(pert))
    ref_rp = O.retrieval_params(D, len(pert))
    assert rp.max_encode_indices_cipher_count == ref_rp["max_encode_indices_cipher_count"]
    ncomb = rp.combination_count
    nciph = rp.payload_cipher_count
    weights = np.zeros((nciph * 2, D), np.uint16)
    weights[:ncomb] = O.chacha12_weights(bytes(range(32)), ncomb * D).reshape(ncomb, D)
    seed = 0xABCDEF
    idx = detector.encode_pertinent_indices(rp, pv, seed=seed, cipher_index=0, n_cipher=rp.max_encode_indices_cipher_count)
    pay = detector.encode_pertinent_payloads(pv, payloads, ncomb, 2, weights)
    idx_h = idx.cpu().numpy().view(np.uint64); pay_h = pay.cpu().numpy().view(np.uint64)
    for c in range(rp.max_encode_indices_cipher_count):
        assert np.array_equal(idx_h[c], O.encode_indices(D, len(pert), pvh, 0, seed, c))
    assert np.array_equal(pay_h, O.encode_payloads(pvh, payloads, 0, weights, nciph))
    st, found, solved = keypack.decode_digest(D, len(pert), idx_h, pay_h, weights)
    assert st == 0 and list(found) == pert
    for i, p in zip(found, solved):
        assert np.array_equal(p, payloads[i])

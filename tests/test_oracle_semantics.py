"""The reference's own acceptance assertions, run on the oracle (CPU): omd.rs:45-58 and omr_time_analyze2.rs:220-240,
plus the stage-boundary check values of SURVEY.md A.9."""
import numpy as np

import oracle as O


def test_omd_and_stage_phases(keypack, decoy):
    a0, b0 = keypack.gen_clues(7, 1); a1, b1 = decoy.gen_clues(8, 1)
    assert (keypack.decrypt_clue(a0[0], b0[0]) == 0).all()                       # clue of seven 0's (clue.rs:32)
    a, b = np.concatenate([a0, a1]), np.concatenate([b0, b1])
    l1 = keypack.l1(a, b, threads=2)
    k = [keypack.phase_l1(l1[i]) * 32 / O.Q1 for i in range(2)]
    plain = keypack.decrypt_clue(a1[0], b1[0])
    expect_k = int((plain == 0).sum()) - int((plain == 4).sum())                # +s per 0, -s per 4 (A.5 step 2)
    assert abs(k[0] - 7) < 0.2 and abs(((k[1] - expect_k + 16) % 32) - 16) < 0.2
    ks = keypack.keyswitch(l1)
    assert (ks < 4096).all()
    inter = [keypack.phase_lwe2(ks[i]) / 128 for i in range(2)]
    assert abs(inter[0] - 14) < 0.3 and abs(inter[1] - (7 + expect_k)) < 0.3     # A.9: 13.95 / 7.03
    pv = keypack.trace(keypack.l2(ks, threads=2), threads=2)
    d0, d1 = keypack.decrypt_decode(pv[0]), keypack.decrypt_decode(pv[1])
    assert d0[0] == 1 and not d0[1:].any()                                        # omd.rs:52-53
    assert not d1.any()                                                           # omd.rs:58
    assert np.array_equal(pv, keypack.detect(a, b, threads=2))                    # stages compose to detect


def test_end_to_end_retrieval_small(keypack, decoy):
    """omr_time_analyze2.rs:220-240 at D = 8 with 2 pertinent messages."""
    D, pert = 8, [2, 5]
    a, b = decoy.gen_clues(31, D)
    for i in pert:
        ai, bi = keypack.gen_clues(32, 1, index0=i); a[i], b[i] = ai[0], bi[0]
    pv = keypack.detect(a, b, threads=8)
    rng = np.random.default_rng(5)
    payloads = rng.integers(0, 256, (D, O.PAYLOAD_LEN), dtype=np.uint16)
    rp = O.retrieval_params(D, len(pert))
    nc, ncomb = rp["payload_cipher_count"], rp["combination_count"]
    weights = np.zeros((nc * 2, D), np.uint16)
    weights[:ncomb] = O.chacha12_weights(bytes(range(32)), ncomb * D).reshape(ncomb, D)
    idx = np.stack([O.encode_indices(D, len(pert), pv, 0, 77, c) for c in range(rp["max_encode_indices_cipher_count"])])
    pay = O.encode_payloads(pv, payloads, 0, weights, nc)
    st, found, solved = keypack.decode_digest(D, len(pert), idx, pay, weights)
    assert st == 0 and list(found) == pert
    for i, p in zip(found, solved):
        assert np.array_equal(p, payloads[i])
    # sharding invariance of the packing (the cross-GPU sum): shards [0,3) + [3,8) == whole
    for c in range(2):
        parts = (O.encode_indices(D, len(pert), pv[:3], 0, 77, c).astype(object) + O.encode_indices(D, len(pert), pv[3:], 3, 77, c).astype(object)) % O.Q2
        assert np.array_equal(parts.astype(np.uint64), idx[c])
    parts = (O.encode_payloads(pv[:3], payloads[:3], 0, weights, nc).astype(object) + O.encode_payloads(pv[3:], payloads[3:], 3, weights, nc).astype(object)) % O.Q2
    assert np.array_equal(parts.astype(np.uint64), pay)


def test_singular_matrix_is_reported(keypack):
    """OmrError::InvertibleMatrix (error.rs:4-8, matrix.rs:181-183): all-zero weights cannot be solved."""
    D = 4
    a, b = keypack.gen_clues(41, 1)
    pv = np.zeros((D, 2, 2048), np.uint64); pv[1] = keypack.detect(a, b, threads=1)[0]
    rp = O.retrieval_params(D, 1)
    weights = np.zeros((rp["payload_cipher_count"] * 2, D), np.uint16)
    idx = np.stack([O.encode_indices(D, 1, pv, 0, 3, c) for c in range(rp["max_encode_indices_cipher_count"])])
    pay = O.encode_payloads(pv, np.ones((D, O.PAYLOAD_LEN), np.uint16), 0, weights, rp["payload_cipher_count"])
    st, found, _ = keypack.decode_digest(D, 1, idx, pay, weights)
    assert list(found) == [1] and st == 1

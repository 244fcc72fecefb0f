"""Wire-format round trips: versioned flat blobs for clues, pertinency vectors, digests, payloads, keys."""
import numpy as np
import pytest


def test_roundtrip_and_validation(tmp_path):
    from tfhe_omr_b200 import blobs
    rng = np.random.default_rng(0)
    a = rng.integers(0, 2048, (5, 512), dtype=np.uint16); b = rng.integers(0, 2048, (5, 7), dtype=np.uint16)
    p = str(tmp_path / "clues.omrb")
    blobs.dump(p, "clues", {"a": a, "b": b}, count=5, index0=1000)
    kind, arrs, hdr = blobs.load(p)
    assert kind == "clues" and hdr == {"version": 1, "count": 5, "index0": 1000, "aux": 0}
    assert np.array_equal(arrs["a"], a) and np.array_equal(arrs["b"], b)
    pv = rng.integers(0, 2**50, (3, 2, 2048), dtype=np.uint64)
    p2 = str(tmp_path / "pv.omrb")
    blobs.dump(p2, "pertinency_vector", {"pv": pv}, count=3, index0=7)
    kind, arrs, hdr = blobs.load(p2, mmap=True)
    assert kind == "pertinency_vector" and np.array_equal(np.asarray(arrs["pv"]), pv)
    with pytest.raises(ValueError):
        blobs.dump(p2, "digest", {"ct": pv[:, :1]}, count=3)                 # wrong shape
    raw = bytearray(open(p, "rb").read())
    raw[0] ^= 1
    bad = str(tmp_path / "bad.omrb"); open(bad, "wb").write(raw)
    with pytest.raises(ValueError):
        blobs.load(bad)                                                       # bad magic
    open(bad, "wb").write(open(p, "rb").read()[:-10])
    with pytest.raises(ValueError):
        blobs.load(bad)                                                       # truncated


def test_detection_key_blob_layout_matches_abi(tmp_path):
    """the key blob is exactly the four arrays of omr_key_blobs, in order"""
    from tfhe_omr_b200 import blobs
    from tfhe_omr_b200.detector import BSK1_SHAPE, KSK_SHAPE, BSK2_SHAPE, TRACE_SHAPE
    fields = dict((n, s) for n, _, s in blobs.KINDS[1][1])
    assert (fields["bsk1"], fields["ksk"], fields["bsk2"], fields["trace"]) == (BSK1_SHAPE, KSK_SHAPE, BSK2_SHAPE, TRACE_SHAPE)

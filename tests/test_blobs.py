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
    assert kind == "clues" and hdr == {"version": 1, "count": 5, "index0": 1000, "aux": 0, "domain": 0}
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


def _c_write(L, _l, path, kind, count, index0, aux, domain, arrays):
    import ctypes as C
    ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return L.omr_blob_write(path.encode(), kind, count, index0, aux, domain, ptrs, len(arrays))


def _c_read(L, _l, path):
    import ctypes as C
    h = _l.BlobHeader()
    st = L.omr_blob_read_header(path.encode(), C.byref(h))
    if st:
        return st, None, None
    n = L.omr_blob_field_count(h.kind)
    bufs = [np.empty(L.omr_blob_field_bytes(h.kind, i, h.count), np.uint8) for i in range(n)]
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
    st = L.omr_blob_read(path.encode(), C.byref(h), ptrs, n)
    return st, h, bufs


def test_python_c_and_oracle_agree_on_the_container(tmp_path):
    """the Rust shim, the oracle and the GPU library interoperate — blobs written by tfhe_omr_b200.blobs load in
    libomr_b200.so (omr_blob_read) and in the oracle (orc_blob_read), and blobs written by either C side load in Python."""
    import oracle as O
    from tfhe_omr_b200 import _lib, blobs
    L = _lib.load()
    rng = np.random.default_rng(5)
    cases = {
        "clues": ({"a": rng.integers(0, 2048, (4, 512), dtype=np.uint16), "b": rng.integers(0, 2048, (4, 7), dtype=np.uint16)}, 4, 9, 0, 0),
        "pertinency_vector": ({"pv": rng.integers(0, 2**50, (2, 2, 2048), dtype=np.uint64)}, 2, 65000, 0, 1),
        "digest": ({"ct": rng.integers(0, 2**50, (3, 2, 2048), dtype=np.uint64)}, 3, 0, 1, 0),
        "payloads": ({"payloads": rng.integers(0, 256, (5, 612), dtype=np.uint16)}, 5, 0, 0, 0),
        "secret_key": ({"s0": rng.integers(0, 2, 512, dtype=np.int32), "z1": rng.integers(-1, 2, 1024, dtype=np.int32),
                        "s2": rng.integers(0, 2, 670, dtype=np.int32), "z2": rng.integers(-1, 2, 2048, dtype=np.int32)}, 0, 0, 0, 0),
        "rlwe1": ({"ct": rng.integers(0, 2**27, (2, 2, 1024), dtype=np.uint32)}, 2, 0, 0, 1),
        "lwe2": ({"ct": rng.integers(0, 4096, (2, 671), dtype=np.uint32)}, 2, 0, 0, 0),
        "rlwe2": ({"ct": rng.integers(0, 2**50, (1, 2, 2048), dtype=np.uint64)}, 1, 0, 0, 1),
        "clue_key": ({"pa": rng.integers(0, 2048, 512, dtype=np.uint16), "pb": rng.integers(0, 2048, 512, dtype=np.uint16)}, 0, 0, 0, 0),
    }
    for kind, (arrs, count, index0, aux, domain) in cases.items():
        kid = blobs._BY_NAME[kind]
        order = [n for n, _, _ in blobs.KINDS[kid][1]]
        p_py, p_c, p_o = (str(tmp_path / f"{kind}.{w}.omrb") for w in ("py", "c", "orc"))
        blobs.dump(p_py, kind, arrs, count=count, index0=index0, aux=aux, domain=domain)
        # Python -> C
        st, h, bufs = _c_read(L, _lib, p_py)
        assert st == 0 and (h.kind, h.count, h.index0, h.aux, h.domain) == (kid, count, index0, aux, domain)
        for name, buf in zip(order, bufs):
            assert buf.tobytes() == np.ascontiguousarray(arrs[name]).tobytes(), (kind, name)
        # C -> Python, byte-identical files
        assert _c_write(L, _lib, p_c, kid, count, index0, aux, domain, [np.ascontiguousarray(arrs[n]) for n in order]) == 0
        assert open(p_c, "rb").read() == open(p_py, "rb").read()
        k2, a2, h2 = blobs.load(p_c)
        assert k2 == kind and h2["domain"] == domain and all(np.array_equal(a2[n], arrs[n]) for n in order)
        # oracle reads the Python file and writes an identical one
        hdr = np.zeros(8, np.uint64)
        assert O.lib().orc_blob_read(p_py.encode(), O.ptr(hdr), None, 0) == 0
        assert list(hdr[:7]) == [1, kid, count, index0, aux, sum(b.nbytes for b in bufs), domain]
        payload = np.zeros(int(hdr[5]), np.uint8)
        assert O.lib().orc_blob_read(p_py.encode(), O.ptr(hdr), O.ptr(payload), payload.nbytes) == 0
        assert payload.tobytes() == b"".join(np.ascontiguousarray(arrs[n]).tobytes() for n in order)
        assert O.lib().orc_blob_write(p_o.encode(), kid, count, index0, aux, domain, O.ptr(payload), payload.nbytes) == 0
        assert open(p_o, "rb").read() == open(p_py, "rb").read()
    # validation on the C side: bad magic, truncation, trailing bytes, wrong array count, unknown kind
    good = open(str(tmp_path / "clues.py.omrb"), "rb").read()
    for name, data in (("magic", b"X" + good[1:]), ("trunc", good[:-3]), ("trail", good + b"\0"), ("kind", good[:12] + b"\x63\0\0\0" + good[16:])):
        p = str(tmp_path / f"bad_{name}.omrb"); open(p, "wb").write(data)
        st, _, _ = _c_read(L, _lib, p)
        assert st != 0 and L.omr_last_error(None), name
        if name != "kind":
            with pytest.raises(ValueError):
                blobs.load(p)
    assert _c_write(L, _lib, str(tmp_path / "x.omrb"), 2, 1, 0, 0, 0, [np.zeros(512, np.uint16)]) != 0      # clues need two arrays
    assert L.omr_blob_field_count(99) == 0 and L.omr_blob_field_bytes(1, 3, 0) == 11 * 25 * 2 * 2048 * 8


def test_create_from_blob_rejects_wrong_kind_without_touching_the_gpu(tmp_path):
    import ctypes as C
    from tfhe_omr_b200 import _lib, blobs
    L = _lib.load()
    p = str(tmp_path / "pay.omrb")
    blobs.dump(p, "payloads", {"payloads": np.zeros((1, 612), np.uint16)}, count=1)
    h = C.c_void_p()
    assert L.omr_ctx_create_from_blob(0, p.encode(), C.byref(h)) == _lib.OMR_ERR_INVALID and not h.value
    assert b"not a detection-key blob" in L.omr_last_error(None)

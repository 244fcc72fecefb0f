"""CPU-side checks of the C-ABI library and the host mirror: it loads, exports every symbol include/omr_b200.h
declares, its host-only functions agree with the oracle, and it fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from tfhe_omr_b200 import _lib
    return _lib, _lib.load()


def test_header_symbols_all_exported():
    _l, L = _lib()
    hdr = open(os.path.join(ROOT, "include", "omr_b200.h")).read()
    declared = set(re.findall(r"\b(omr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_l.EXPORTS), declared ^ set(_l.EXPORTS)
    for s in declared:
        assert hasattr(L, s), s


def test_retrieval_params_match_reference_layout():
    _l, L = _lib()
    import tfhe_omr_b200 as omr
    for D in (1, 2, 50, 256, 257, 258, 4096, 65536, 66049, 66050):
        pert = min(D, 50)
        c = _l.RetrievalParamsC()
        assert L.omr_retrieval_params_init(D, pert, C.byref(c)) == 0
        ref = O.retrieval_params(D, pert)
        py = omr.RetrievalParams(D, pert)
        for k in ("slots_per_bucket", "slots_per_segment", "segment_per_cipher", "max_encode_indices_cipher_count", "combination_count"):
            assert getattr(c, k) == ref[k] == getattr(py, k), (D, k)
        assert py.payload_cipher_count == ref["payload_cipher_count"]


def test_create_fails_loudly_without_gpu_or_with_bad_args():
    import torch
    _l, L = _lib()
    h = C.c_void_p()
    assert L.omr_ctx_create(0, None, C.byref(h)) == _l.OMR_ERR_INVALID
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    dummy = np.zeros(16, np.uint64)
    blobs = _l.KeyBlobs(dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, dummy.ctypes.data, 0)
    st = L.omr_ctx_create(0, C.byref(blobs), C.byref(h))
    assert st == _l.OMR_ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.omr_last_error(None)
    import tfhe_omr_b200 as omr
    with pytest.raises(omr.OmrError):
        omr.Detector(omr.DetectionKey(dummy, dummy, dummy, dummy))


def test_product_never_imports_oracle():
    """the product path must not route through oracle/ (nor any CPU fallback)"""
    pkg = os.path.join(ROOT, "tfhe-omr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"import\s+oracle", r"from\s+oracle", r"#include\s*[\"<][^\n]*oracle", r"libomr_oracle", r"oracle/", r"orc_[a-z]"):
                    assert not re.search(pat, src), (os.path.join(dirpath, f), pat)

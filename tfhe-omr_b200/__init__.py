"""tfhe_omr_b200 — B200 (sm_100a) implementation of InstantOMR's detection hot path behind the reference's
`Detector` API (omr_core/src/lib.rs:21-31).  Kernels live in csrc/ and are reached through the C ABI of
include/omr_b200.h (lib/libomr_b200.so); this package is the Python host-side mirror used by tests and bench.py."""
from .params import OmrParameters, RetrievalParams, PAYLOAD_LENGTH
from .detector import Detector, DetectionKey, DetectTimeInfo, PertinencyVector, OmrError
from .retriever import Retriever, InvertibleMatrix, solve_matrix_mod_257

__all__ = ["OmrParameters", "RetrievalParams", "PAYLOAD_LENGTH", "Detector", "DetectionKey", "DetectTimeInfo",
           "PertinencyVector", "OmrError", "Retriever", "InvertibleMatrix", "solve_matrix_mod_257"]

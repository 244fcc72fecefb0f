"""Host part of the recipient side (tfhe_omr_b200.retriever): the mod-257 solver of matrix.rs:164-247."""
import numpy as np
import pytest


def test_solver_recovers_planted_solution_and_detects_singularity():
    import tfhe_omr_b200 as omr
    rng = np.random.default_rng(0)
    for rows, cols in ((55, 50), (6, 1), (8, 3)):
        M = rng.integers(0, 257, (rows, cols))
        X = rng.integers(0, 257, (cols, 612))
        Y = (M @ X) % 257
        got = omr.solve_matrix_mod_257(M, Y)
        assert got.shape == (cols, 612) and np.array_equal(got, X)
    with pytest.raises(omr.InvertibleMatrix):                   # OmrError::InvertibleMatrix (error.rs:4-8)
        omr.solve_matrix_mod_257(np.zeros((6, 2), np.int64), np.ones((6, 612), np.int64))
    M = rng.integers(0, 257, (7, 3)); M[:, 2] = (M[:, 0] * 5 + M[:, 1]) % 257     # rank deficient
    with pytest.raises(omr.InvertibleMatrix):
        omr.solve_matrix_mod_257(M, rng.integers(0, 257, (7, 612)))
    with pytest.raises(omr.InvertibleMatrix):                   # fewer rows than columns (matrix.rs:171)
        omr.solve_matrix_mod_257(np.ones((2, 3), np.int64), np.ones((2, 612), np.int64))

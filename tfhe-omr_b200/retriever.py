"""Host-side mirror of the reference's `Retriever` (omr_core/src/retriever.rs:26-387) and of its mod-257 solver
(omr_core/src/matrix.rs:164-247) — SURVEY.md §8f.1.

Decryption, inverse NTT and the exact-integer decode of every slot run on the GPU (omr_decrypt_decode_device); the bucket
scan, the weight-matrix lookup and the Gaussian elimination over a 55 x 50 system are host work (microseconds in the
reference, `README.md:125`).  Errors follow the reference: `OmrError` with `InvertibleMatrix` (error.rs:4-8) when the
combination matrix is singular.
"""
import numpy as np

from . import _lib
from .detector import OmrError, N2
from .params import PAYLOAD_LENGTH

P = 257


class InvertibleMatrix(OmrError):
    """OmrError::InvertibleMatrix (error.rs:4-8)."""

    def __init__(self):
        RuntimeError.__init__(self, "matrix is not invertible")
        self.status = _lib.OMR_ERR_INVALID


def solve_matrix_mod_257(matrix, payloads):
    """solve_matrix_mod_257 (matrix.rs:164-247): Gaussian elimination over Z_257 on (matrix [rows][cols], payloads
    [rows][612]); returns the first `cols` payload rows.  First non-zero pivot, row swap, normalise, eliminate below, then
    back-substitute — the same order of operations as the reference."""
    m = np.array(matrix, dtype=np.int64) % P
    pl = np.array(payloads, dtype=np.int64) % P
    rows, cols = m.shape
    if rows < cols:
        raise InvertibleMatrix()
    for i in range(cols):
        nz = np.nonzero(m[i:, i])[0]
        if nz.size == 0:
            raise InvertibleMatrix()                       # matrix.rs:181-183
        pick = i + int(nz[0])
        if pick != i:
            m[[i, pick]] = m[[pick, i]]; pl[[i, pick]] = pl[[pick, i]]
        v = int(m[i, i])
        if v != 1:
            inv = pow(v, P - 2, P)                         # INV_MOD_257[v] (matrix.rs:28-41)
            m[i, i:] = m[i, i:] * inv % P
            pl[i] = pl[i] * inv % P
        if i == cols - 1:
            break
        c = m[i + 1:, i].copy()
        sel = np.nonzero(c)[0]
        if sel.size:
            m[i + 1 + sel, i:] = (m[i + 1 + sel, i:] - c[sel, None] * m[i, i:]) % P
            pl[i + 1 + sel] = (pl[i + 1 + sel] - c[sel, None] * pl[i]) % P
    for ic in range(cols - 1, 0, -1):
        c = m[:ic, ic].copy()
        sel = np.nonzero(c)[0]
        if sel.size:
            pl[sel] = (pl[sel] - c[sel, None] * pl[ic]) % P
            m[sel, ic] = 0
    return pl[:cols].astype(np.uint16)


def _to_host_u64(x):
    if hasattr(x, "data_ptr"):
        return x.cpu().numpy().view(np.uint64)
    return np.asarray(x).view(np.uint64) if np.asarray(x).dtype != np.uint64 else np.asarray(x)


class Retriever:
    """Retriever<F> (retriever.rs:26-61): holds the retrieval layout and the recipient's NTT-domain secret z2."""

    def __init__(self, detector, params, z2_ntt):
        import torch
        self.detector = detector
        self.params = params
        if not hasattr(z2_ntt, "data_ptr"):
            z2_ntt = torch.from_numpy(np.ascontiguousarray(z2_ntt, np.uint64).view(np.int64)).to(f"cuda:{detector.device}")
        if z2_ntt.numel() != N2:
            raise OmrError(_lib.OMR_ERR_INVALID, "z2_ntt must have 2048 coefficients")
        self.key = z2_ntt
        self.pertinent_indices_set = set()

    def _slots(self, cts):
        import torch
        if not hasattr(cts, "data_ptr"):
            cts = torch.from_numpy(np.ascontiguousarray(cts, np.uint64).view(np.int64)).to(self.key.device)
        return self.detector.decrypt_decode(self.key, cts.reshape(-1, 2, N2)).cpu().numpy().view(np.uint16)

    def decode_pertinent_indices(self, encoded_indices):
        """decode_pertinent_indices (retriever.rs:63-130) for one or several index ciphertexts; returns True when the set
        has reached pertinent_count (the reference's Ok(..))."""
        rp = self.params
        w, S = rp.slots_per_bucket, rp.slots_per_segment
        for dec in self._slots(encoded_indices):
            seg = dec[:rp.segment_per_cipher * S].reshape(rp.segment_per_cipher, rp.bucket_count_per_segment, w).astype(np.int64)
            hit = seg[..., w - 1] == 1                                   # a bucket counts only if its flag slot decodes to exactly 1
            idx = np.zeros(hit.shape, np.int64)
            for k in range(w - 2, -1, -1):                               # fold from the most significant digit (:113-116)
                idx = idx * rp.index_modulus + seg[..., k]
            self.pertinent_indices_set.update(int(v) for v in idx[hit])
            if len(self.pertinent_indices_set) == rp.pertinent_count:
                return True
        return len(self.pertinent_indices_set) == rp.pertinent_count

    def decode_combined_payloads(self, combinations):
        """decode_combined_payloads (retriever.rs:318-362): [combination_count][612]."""
        rp = self.params
        dec = self._slots(combinations)
        out = np.zeros((rp.combination_count, PAYLOAD_LENGTH), np.uint16)
        for c in range(dec.shape[0]):
            for j in range(rp.cmb_count_per_cipher):
                r = c * rp.cmb_count_per_cipher + j
                if r < rp.combination_count:
                    out[r] = dec[c, j * PAYLOAD_LENGTH:(j + 1) * PAYLOAD_LENGTH]
        return out

    def _weights(self, weights, seed):
        if weights is not None:
            return weights
        if seed is None:
            raise OmrError(_lib.OMR_ERR_INVALID, "either weights or the 32-byte seed is required")
        rp = self.params                                     # regenerate the matrix from the seed (retriever.rs:215-226)
        return self.detector.weights_from_seed(seed, rp.combination_count, rp.all_payloads_count).cpu().numpy().view(np.uint16)

    def decode_digest_host(self, encode_pertinent_indices, encode_pertinent_payloads, weights=None, seed=None):
        """decode_digest (retriever.rs:188-260) through the C ABI (omr_decode_digest): host arrays in, GPU decrypt/decode,
        bucket scan and mod-257 solver in the library.  Returns (sorted indices, payloads [len(indices)][612]); raises
        InvertibleMatrix when the combination matrix is singular."""
        import ctypes as C
        rp = self.params
        idx = np.ascontiguousarray(_to_host_u64(encode_pertinent_indices)).reshape(-1, 2, N2)
        pay = np.ascontiguousarray(_to_host_u64(encode_pertinent_payloads)).reshape(-1, 2, N2)
        w = np.ascontiguousarray(self._weights(weights, seed), np.uint16)
        if w.ndim != 2 or w.shape[0] < rp.combination_count:
            raise OmrError(_lib.OMR_ERR_INVALID, "weights must be [combination_count][all_payloads_count]")
        key = np.ascontiguousarray(self.key.cpu().numpy().view(np.uint64))
        out_idx = np.zeros(max(1, rp.pertinent_count), np.uint64)
        out_pay = np.zeros((max(1, rp.pertinent_count), PAYLOAD_LENGTH), np.uint16)
        n_found = C.c_uint32(0)
        rpc = rp.to_c()
        det = self.detector
        st = det.L.omr_decode_digest(det.h, C.byref(rpc), key.ctypes.data, idx.ctypes.data, idx.shape[0], pay.ctypes.data, pay.shape[0],
                                     w.ctypes.data, w.shape[1], out_idx.ctypes.data, C.byref(n_found), out_pay.ctypes.data)
        if st != _lib.OMR_OK:
            msg = (det.L.omr_last_error(det.h) or b"").decode()
            if "not invertible" in msg:
                raise InvertibleMatrix()
            raise OmrError(st, msg)
        n = int(n_found.value)
        self.pertinent_indices_set.update(int(v) for v in out_idx[:n])
        return [int(v) for v in out_idx[:n]], out_pay[:n].copy()

    def decode_digest(self, encode_pertinent_indices, encode_pertinent_payloads, weights=None, seed=None):
        """decode_digest (retriever.rs:188-260).  `weights` [combination_count (or more)][D] is the matrix the reference
        regenerates from the 32-byte seed (retriever.rs:215-226); pass either the matrix or the `seed` itself.
        Returns (sorted indices, payloads [len(indices)][612]); raises InvertibleMatrix."""
        rp = self.params
        weights = self._weights(weights, seed)
        for ct in encode_pertinent_indices:                              # ciphertexts are consumed until the set is full (:200-204)
            if self.decode_pertinent_indices(ct[None] if ct.ndim == 2 else ct):
                break
        indices = sorted(self.pertinent_indices_set)
        if not indices:
            return [], np.zeros((0, PAYLOAD_LENGTH), np.uint16)
        weights = np.asarray(weights)
        if any(i >= rp.all_payloads_count for i in indices):
            raise OmrError(_lib.OMR_ERR_INVALID, "decoded index outside the board")     # the reference would panic on the lookup
        matrix = weights[:rp.combination_count][:, indices]
        combined = self.decode_combined_payloads(encode_pertinent_payloads)
        return indices, solve_matrix_mod_257(matrix, combined)

"""Host-side mirror of the reference's `Detector` (omr_core/src/detector.rs:35-453) over libomr_b200.so.

Same names and argument meaning as the Rust API; the differences are the ones a GPU back end forces:
  * `detect` takes a BATCH of clues (it replaces `clues_list.par_iter().map(|c| detector.detect(c))`,
    examples/omr.rs:160-164) and returns a device-resident `PertinencyVector`;
  * `encode_pertinent_payloads` takes the weights explicitly instead of an RNG (the caller draws them exactly as
    detector.rs:376-387 does; `chacha` weights need the Rust side or its clone);
  * `encode_pertinent_indices` takes a seed: the reference uses thread_rng (detector.rs:262).
torch is used for device memory / streams only; every computation is a kernel of libomr_b200.so.  Without the
library or without a CUDA device this module raises — there is no CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .params import OmrParameters, RetrievalParams, PAYLOAD_LENGTH

N1, N2, CLUE_N, CLUE_COUNT, LWE2_N = 1024, 2048, 512, 7, 670
BSK1_SHAPE, KSK_SHAPE, BSK2_SHAPE, TRACE_SHAPE = (512, 8, 2, 1024), (1024, 27, 671), (670, 12, 2, 2048), (11, 25, 2, 2048)


class OmrError(RuntimeError):
    """Errors of the detector side.  The reference panics/asserts here (detector.rs:236,511); OmrError::
    InvertibleMatrix (error.rs:4-8) belongs to the recipient side."""

    def __init__(self, status, message):
        super().__init__(f"omr_b200 status {status}: {message}")
        self.status = status


@dataclass
class DetectTimeInfo:
    """DetectTimeInfo / DetectTimeInfoPerMessage (detector.rs:43-80): device milliseconds, summed over a batch."""
    total_detect_time: float = 0.0
    total_first_level_bootstrapping_time: float = 0.0
    total_second_level_bootstrapping_time: float = 0.0
    total_trace_time: float = 0.0

    def __add__(self, rhs):
        return DetectTimeInfo(self.total_detect_time + rhs.total_detect_time,
                              self.total_first_level_bootstrapping_time + rhs.total_first_level_bootstrapping_time,
                              self.total_second_level_bootstrapping_time + rhs.total_second_level_bootstrapping_time,
                              self.total_trace_time + rhs.total_trace_time)


class DetectionKey:
    """DetectionKey (key_gen/detection.rs:9-16) flattened to the four blobs of omr_key_blobs.
    Arrays may be numpy (host) or torch CUDA tensors (all four on the same device)."""

    def __init__(self, bsk1, ksk, bsk2, trace, coeff_domain=False):
        self.bsk1, self.ksk, self.bsk2, self.trace = bsk1, ksk, bsk2, trace
        self.coeff_domain = coeff_domain
        self.params = OmrParameters()

    def size(self):
        return int(np.prod(BSK1_SHAPE)) * 4 + int(np.prod(KSK_SHAPE)) * 4 + int(np.prod(BSK2_SHAPE)) * 8 + int(np.prod(TRACE_SHAPE)) * 8


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise OmrError(_lib.OMR_ERR_CUDA, "no CUDA device visible: tfhe_omr_b200 has no CPU fallback")
    return torch


class PertinencyVector:
    """`Vec<NttRlweCiphertext<SecondLevelField>>` kept in HBM: int64 tensor [count][2][2048] (bit pattern = u64)."""

    def __init__(self, tensor, index0=0):
        self.tensor = tensor
        self.index0 = index0

    def __len__(self):
        return self.tensor.shape[0]

    def to_host(self):
        return self.tensor.cpu().numpy().view(np.uint64)


class Detector:
    """Detector (detector.rs:35-39) bound to one GPU."""

    def __init__(self, detection_key, device=None):
        self.L = _lib.load()
        torch = _torch()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.detection_key = detection_key
        blobs = _lib.KeyBlobs()
        on_device = hasattr(detection_key.bsk1, "data_ptr")
        keep = []
        for name, shape, dt in (("bsk1", BSK1_SHAPE, np.uint32), ("ksk", KSK_SHAPE, np.uint32),
                                ("bsk2", BSK2_SHAPE, np.uint64), ("trace", TRACE_SHAPE, np.uint64)):
            arr = getattr(detection_key, name)
            if on_device:
                if arr.numel() * arr.element_size() != int(np.prod(shape)) * np.dtype(dt).itemsize or not arr.is_contiguous():
                    raise OmrError(_lib.OMR_ERR_INVALID, f"{name}: wrong size or not contiguous")
                setattr(blobs, name, arr.data_ptr())
            else:
                arr = np.ascontiguousarray(arr, dt)
                if arr.size != int(np.prod(shape)):
                    raise OmrError(_lib.OMR_ERR_INVALID, f"{name}: expected shape {shape}")
                keep.append(arr)
                setattr(blobs, name, arr.ctypes.data)
        blobs.flags = _lib.KEYS_COEFF if detection_key.coeff_domain else _lib.KEYS_NTT_NATIVE
        h = C.c_void_p()
        create = self.L.omr_ctx_create_device_keys if on_device else self.L.omr_ctx_create
        st = create(self.device, C.byref(blobs), C.byref(h))
        if st != _lib.OMR_OK:
            raise OmrError(st, (self.L.omr_last_error(None) or b"").decode())
        self.h = h          # (omr_ctx_create_device_keys drains the device before copying: the tensors may come from any torch stream)

    @classmethod
    def from_blob(cls, path, device=None):
        """Detector::new from a detection-key blob on disk (omr_ctx_create_from_blob; the blob's domain selects the key flags)"""
        torch = _torch()
        self = cls.__new__(cls)
        self.L = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        st = self.L.omr_ctx_create_from_blob(self.device, str(path).encode(), C.byref(h))
        if st != _lib.OMR_OK:
            raise OmrError(st, (self.L.omr_last_error(None) or b"").decode())
        self.h = h
        self.detection_key = None
        return self

    @classmethod
    def generate(cls, secrets, seed, device=None, want_keys=False):
        """SecretKeyPack::generate_detector (key_gen/secret.rs:118-187) with the key material made on the GPU
        (omr_generate_detector).  secrets = (s0 [512] binary, z1 [1024] ternary, s2 [670] binary, z2 [2048] ternary) int32 arrays,
        seed = 32 bytes from the caller's CSPRNG.  want_keys=True also returns the flat detection key as a DetectionKey of numpy
        arrays (what the recipient ships to a detector)."""
        torch = _torch()
        self = cls.__new__(cls)
        self.L = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        seed = bytes(seed)
        if len(seed) != 32:
            raise OmrError(_lib.OMR_ERR_INVALID, "the seed is 32 bytes")
        arrs = [np.ascontiguousarray(x, np.int32) for x in secrets]
        if [a.size for a in arrs] != [CLUE_N, N1, LWE2_N, N2]:
            raise OmrError(_lib.OMR_ERR_INVALID, "secrets must be (s0[512], z1[1024], s2[670], z2[2048])")
        sk = _lib.SecretKey(*[a.ctypes.data for a in arrs])
        dk, blobs = None, None
        if want_keys:
            dk = DetectionKey(np.empty(BSK1_SHAPE, np.uint32), np.empty(KSK_SHAPE, np.uint32), np.empty(BSK2_SHAPE, np.uint64), np.empty(TRACE_SHAPE, np.uint64))
            blobs = _lib.KeyBlobs(dk.bsk1.ctypes.data, dk.ksk.ctypes.data, dk.bsk2.ctypes.data, dk.trace.ctypes.data, _lib.KEYS_NTT_NATIVE)
        h = C.c_void_p()
        st = self.L.omr_generate_detector(self.device, C.byref(sk), seed, C.byref(blobs) if blobs is not None else None, C.byref(h))
        if st != _lib.OMR_OK:
            raise OmrError(st, (self.L.omr_last_error(None) or b"").decode())
        self.h = h
        self.detection_key = dk
        return self

    def close(self):
        if getattr(self, "h", None):
            self.L.omr_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st != _lib.OMR_OK:
            raise OmrError(st, (self.L.omr_last_error(self.h) or b"").decode())

    # -- accessors (detector.rs:112-132) --
    def detect_key_size(self):
        return int(self.L.omr_detect_key_size(self.h))

    def launch_count(self):
        return int(self.L.omr_launch_count(self.h))

    def first_level_lut(self):
        """Detector::first_level_lut (detector.rs:117-123): u32[1024] mod q1, coefficient form"""
        out = np.empty(N1, np.uint32); self._ck(self.L.omr_first_level_lut(self.h, out.ctypes.data)); return out

    def second_level_lut(self):
        """Detector::second_level_lut (detector.rs:126-132): u64[2048] mod q2, coefficient form"""
        out = np.empty(N2, np.uint64); self._ck(self.L.omr_second_level_lut(self.h, out.ctypes.data)); return out

    def set_output_domain(self, coeff):
        """omr_set_output_domain: host-buffer ciphertexts in coefficient form (True; what a Rust shim uses so that nothing depends
        on Primus-fhe's NTT ordering) or in this library's NTT ordering (False, default)"""
        self._ck(self.L.omr_set_output_domain(self.h, _lib.OUT_COEFF if coeff else _lib.OUT_NTT_NATIVE))

    def key_switch_path(self):
        """'cuda-core' (hand-written kernels, default) or 'tensor-core' (opt-in CUTLASS int8 GEMM)"""
        return "tensor-core" if self.L.omr_key_switch_path(self.h) else "cuda-core"

    def weights_from_seed(self, seed, rows, cols, in_order=False):
        """The reference's combination weights (detector.rs:376-387): StdRng::from_seed(seed) + Uniform(0, 257), rows x cols
        draws in stream order, generated on the GPU (omr_weights_from_seed_device).  Returns a CUDA int16 tensor."""
        torch = _torch()
        seed = bytes(seed)
        if len(seed) != 32:
            raise OmrError(_lib.OMR_ERR_INVALID, "the seed is 32 bytes")
        out = torch.empty((rows, cols), dtype=torch.int16, device=f"cuda:{self.device}")
        self._ck(self.L.omr_weights_from_seed_device(self.h, seed, rows * cols, out.data_ptr(), 1 if in_order else 0, self._stream()))
        return out

    def set_tensor_core_key_switch(self, enable):
        """omr_set_tensor_core_key_switch: the key switch as an int8 tensor-core GEMM (default when built in) or on CUDA cores."""
        self._ck(self.L.omr_set_tensor_core_key_switch(self.h, 1 if enable else 0))

    def set_latency_shapes(self, enable):
        """omr_set_latency_shapes: small batches use the latency launch shapes (default) or the throughput shapes."""
        self._ck(self.L.omr_set_latency_shapes(self.h, 1 if enable else 0))

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    # -- detect (detector.rs:135-166) ------------------------------------------------------------------------------
    def detect(self, clues, index0=0, times=None):
        """clues = (a [B][512] u16, b [B][7] u16) as numpy arrays (host; copied inside) or torch CUDA int16 tensors.
        Returns a PertinencyVector.  Wrong clue shape raises (the reference asserts the clue count, detector.rs:511)."""
        torch = _torch()
        a, b = clues
        if hasattr(a, "data_ptr"):
            da, db = a, b
        else:
            a = np.ascontiguousarray(a, np.uint16).reshape(-1, CLUE_N)
            b = np.ascontiguousarray(b, np.uint16)
            if b.size != a.shape[0] * CLUE_COUNT:
                raise OmrError(_lib.OMR_ERR_INVALID, "Invalid clue count.")
            da = torch.from_numpy(a.view(np.int16)).to(f"cuda:{self.device}")
            db = torch.from_numpy(b.reshape(-1, CLUE_COUNT).view(np.int16)).to(f"cuda:{self.device}")
        B = da.shape[0]
        if da.numel() != B * CLUE_N or db.numel() != B * CLUE_COUNT:
            raise OmrError(_lib.OMR_ERR_INVALID, "Invalid clue count.")
        pv = torch.empty((B, 2, N2), dtype=torch.int64, device=f"cuda:{self.device}")
        st_times = _lib.StageTimes() if times is not None else None
        self._ck(self.L.omr_detect_batch_device(self.h, da.data_ptr(), db.data_ptr(), B, pv.data_ptr(), self._stream(),
                                                C.byref(st_times) if st_times is not None else None))
        if times is not None:
            times.total_detect_time += st_times.detect_ms
            times.total_first_level_bootstrapping_time += st_times.first_level_bootstrapping_ms
            times.total_second_level_bootstrapping_time += st_times.second_level_bootstrapping_ms
            times.total_trace_time += st_times.trace_ms
        return PertinencyVector(pv, index0)

    def detect_with_time_info(self, clues, index0=0):
        """detect_with_time_info (detector.rs:169-221)."""
        t = DetectTimeInfo()
        pv = self.detect(clues, index0, times=t)
        return pv, t

    # -- digest packing ----------------------------------------------------------------------------------------------
    def encode_pertinent_indices(self, retrieval_params, pertinency_vector, seed=0, cipher_index=0, n_cipher=1, out=None):
        """encode_pertinent_indices (detector.rs:223-339).  The reference calls it max_encode_indices_cipher_count
        times (examples/omr.rs:180-183); pass cipher_index / n_cipher to build several ciphertexts in one launch.
        Returns an int64 CUDA tensor [n_cipher][2][2048] = this GPU's digest contribution (already mod q2)."""
        torch = _torch()
        if retrieval_params.polynomial_size != N2:
            raise OmrError(_lib.OMR_ERR_INVALID, "polynomial_size != ntt dimension")      # detector.rs:236
        pv = pertinency_vector
        if out is None:
            out = torch.empty((n_cipher, 2, N2), dtype=torch.int64, device=pv.tensor.device)
        rp = retrieval_params.to_c()
        self._ck(self.L.omr_encode_indices_device(self.h, C.byref(rp), pv.tensor.data_ptr(), len(pv), pv.index0, seed,
                                                  cipher_index, n_cipher, out.data_ptr(), self._stream()))
        return out

    def seeded_weights(self, seed, combination_count, cmb_count_per_cipher, all_payloads_count):
        """[ceil(cc / per) * per][D] weight matrix of the reference for a 32-byte seed: the first combination_count rows are
        the ChaCha12 stream (weights_from_seed), the unused tail rows stay zero (detector.rs:370-371)."""
        torch = _torch()
        rows = -(-combination_count // cmb_count_per_cipher) * cmb_count_per_cipher
        w = torch.zeros((rows, all_payloads_count), dtype=torch.int16, device=f"cuda:{self.device}")
        w[:combination_count] = self.weights_from_seed(seed, combination_count, all_payloads_count)
        return w

    def encode_pertinent_payloads(self, pertinency_vector, payloads, combination_count, cmb_count_per_cipher, weights=None, out=None,
                                  seed=None, all_payloads_count=None):
        """encode_pertinent_payloads (detector.rs:341-453).  payloads [count][612] u16, weights [rows][D] u16 with
        rows >= ceil(combination_count / cmb_count_per_cipher) * cmb_count_per_cipher (unused tail rows zero,
        detector.rs:370-371), column = global message index.  numpy or CUDA tensors.  Instead of `weights`, the reference's
        32-byte `seed` (the StdRng the reference passes in) and the board size `all_payloads_count` may be given."""
        torch = _torch()
        pv = pertinency_vector
        dev = pv.tensor.device
        n_cipher = -(-combination_count // cmb_count_per_cipher)
        if weights is None:
            if seed is None or all_payloads_count is None:
                raise OmrError(_lib.OMR_ERR_INVALID, "either weights or (seed, all_payloads_count) is required")
            weights = self.seeded_weights(seed, combination_count, cmb_count_per_cipher, all_payloads_count)
        if not hasattr(payloads, "data_ptr"):
            payloads = torch.from_numpy(np.ascontiguousarray(payloads, np.uint16).reshape(-1, PAYLOAD_LENGTH).view(np.int16)).to(dev)
        if not hasattr(weights, "data_ptr"):
            weights = torch.from_numpy(np.ascontiguousarray(weights, np.uint16).view(np.int16)).to(dev)
        if payloads.shape[0] != len(pv) or weights.shape[0] < n_cipher * cmb_count_per_cipher:
            raise OmrError(_lib.OMR_ERR_INVALID, "payload / weight shape mismatch")
        if out is None:
            out = torch.empty((n_cipher, 2, N2), dtype=torch.int64, device=dev)
        self._ck(self.L.omr_encode_payloads_device(self.h, pv.tensor.data_ptr(), payloads.data_ptr(), len(pv), pv.index0,
                                                   weights.data_ptr(), weights.shape[1], n_cipher, cmb_count_per_cipher,
                                                   out.data_ptr(), self._stream()))
        return out

    def digest_reduce_mod(self, digest):
        """mod-q2 reduction after the cross-GPU sum of partial digests (the rayon reduce of detector.rs:333-336,445-448)."""
        self._ck(self.L.omr_digest_reduce_mod(self.h, digest.data_ptr(), digest.numel(), self._stream()))
        return digest

    def digest_accumulate(self, running, part):
        """streaming: running = (running + part) mod q2 (both canonical int64 CUDA tensors of the same shape)"""
        if running.shape != part.shape:
            raise OmrError(_lib.OMR_ERR_INVALID, "digest shapes differ")
        self._ck(self.L.omr_digest_add_mod(self.h, running.data_ptr(), part.data_ptr(), running.numel(), self._stream()))
        return running

    def digest_allreduce(self, digest, comm=None):
        """K7 (omr_digest_allreduce): in-place sum over all ranks of the partial digests, mod q2, on the current stream.
        comm = an ncclComm_t as an int (e.g. torch's ProcessGroupNCCL._comm_ptr()), or None for the communicator made by comm_init."""
        n_cipher = digest.numel() // (2 * N2)
        self._ck(self.L.omr_digest_allreduce(self.h, C.c_void_p(comm) if comm else None, digest.data_ptr(), n_cipher, self._stream()))
        return digest

    def comm_unique_id(self):
        """rank 0: the 128-byte NCCL id to hand to every rank (omr_comm_unique_id)"""
        buf = (C.c_uint8 * 128)()
        st = self.L.omr_comm_unique_id(buf)
        if st != _lib.OMR_OK:
            raise OmrError(st, (self.L.omr_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, n_ranks, rank, unique_id):
        """every rank: join the library's own NCCL communicator (omr_comm_init)"""
        if len(unique_id) != 128:
            raise OmrError(_lib.OMR_ERR_INVALID, "the NCCL id is 128 bytes")
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.omr_comm_init(self.h, n_ranks, rank, buf))

    def comm_destroy(self):
        self._ck(self.L.omr_comm_destroy(self.h))

    # -- streaming ingest (README.md:9; omr_stream_*) -----------------------------------------------------------------------
    def stream_begin(self, retrieval_params, index_seed, weight_seed, global_index0=0):
        weight_seed = bytes(weight_seed)
        if len(weight_seed) != 32:
            raise OmrError(_lib.OMR_ERR_INVALID, "the seed is 32 bytes")
        rp = retrieval_params.to_c()
        self._stream_cts = retrieval_params.max_encode_indices_cipher_count + retrieval_params.payload_cipher_count
        self._ck(self.L.omr_stream_begin(self.h, C.byref(rp), index_seed, weight_seed, global_index0))

    def stream_push(self, a, b, payloads):
        """detect the next messages (host arrays) and fold them into the resident running digest; returns once the inputs are staged"""
        a = np.ascontiguousarray(a, np.uint16).reshape(-1, CLUE_N); b = np.ascontiguousarray(b, np.uint16).reshape(-1, CLUE_COUNT)
        payloads = np.ascontiguousarray(payloads, np.uint16).reshape(-1, PAYLOAD_LENGTH)
        if not (a.shape[0] == b.shape[0] == payloads.shape[0]):
            raise OmrError(_lib.OMR_ERR_INVALID, "Invalid clue count.")
        self._ck(self.L.omr_stream_push(self.h, a.ctypes.data, b.ctypes.data, payloads.ctypes.data, a.shape[0]))

    def stream_snapshot(self):
        """(running digest [n_index + n_payload][2][2048] u64 in the output domain, messages folded in so far)"""
        out = np.empty((self._stream_cts, 2, N2), np.uint64); n = C.c_uint64(0)
        self._ck(self.L.omr_stream_snapshot(self.h, out.ctypes.data, C.byref(n)))
        return out, int(n.value)

    def stream_end(self):
        self._ck(self.L.omr_stream_end(self.h))

    def decrypt_decode(self, z2_ntt, cts):
        """recipient side: decoded slots (values mod 257) of NTT-domain RLWE ciphertexts, int16 CUDA tensor [n][2048]"""
        torch = _torch()
        out = torch.empty((cts.shape[0], N2), dtype=torch.int16, device=cts.device)
        self._ck(self.L.omr_decrypt_decode_device(self.h, z2_ntt.data_ptr(), cts.data_ptr(), cts.shape[0], out.data_ptr(), self._stream()))
        return out

    def gen_clues(self, clue_key, count, seed, index0=0, msgs=None):
        """sender side (Sender::gen_clues, sender.rs:27-30) batched on the GPU: clue_key = (pa, pb) u16[512] each; seed = the 32
        bytes that key the ChaCha12 stream all randomness comes from (the reference asks for a CryptoRng, clue.rs:27-30; an int
        is widened little-endian, for tests).  Returns CUDA int16 tensors (a [count][512], b [count][7])."""
        torch = _torch()
        seed = seed.to_bytes(32, "little") if isinstance(seed, int) else bytes(seed)
        if len(seed) != 32:
            raise OmrError(_lib.OMR_ERR_INVALID, "the seed is 32 bytes")
        dev = f"cuda:{self.device}"
        pa, pb = (torch.from_numpy(np.ascontiguousarray(k, np.uint16).view(np.int16)).to(dev) if not hasattr(k, "data_ptr") else k for k in clue_key)
        if pa.numel() != CLUE_N or pb.numel() != CLUE_N:
            raise OmrError(_lib.OMR_ERR_INVALID, "clue key must be two polynomials of 512 coefficients")
        a = torch.empty((count, CLUE_N), dtype=torch.int16, device=dev); b = torch.empty((count, CLUE_COUNT), dtype=torch.int16, device=dev)
        dm = None
        if msgs is not None:
            dm = torch.from_numpy(np.ascontiguousarray(msgs, np.uint8).reshape(count, CLUE_COUNT)).to(dev)
        self._ck(self.L.omr_gen_clues_device(self.h, pa.data_ptr(), pb.data_ptr(), seed, index0, count,
                                             dm.data_ptr() if dm is not None else None, a.data_ptr(), b.data_ptr(), self._stream()))
        return a, b

    # -- host-buffer ("e2e") path: what the Rust shim binds ------------------------------------------------------------
    def pv_reset(self):
        self._ck(self.L.omr_pv_reset(self.h))

    def pv_load(self, pv, global_index0=0):
        """omr_pv_load: replace the resident store by host pertinency ciphertexts [count][2][2048] (output domain of the context)"""
        pv = np.ascontiguousarray(pv, np.uint64).reshape(-1, 2, N2)
        self._ck(self.L.omr_pv_load(self.h, pv.ctypes.data, pv.shape[0], global_index0))

    def detect_host(self, a, b, global_index0=0, want_pv=False):
        a = np.ascontiguousarray(a, np.uint16).reshape(-1, CLUE_N); b = np.ascontiguousarray(b, np.uint16).reshape(-1, CLUE_COUNT)
        if a.shape[0] != b.shape[0]:
            raise OmrError(_lib.OMR_ERR_INVALID, "Invalid clue count.")
        pv = np.empty((a.shape[0], 2, N2), np.uint64) if want_pv else None
        self._ck(self.L.omr_detect_batch(self.h, a.ctypes.data, b.ctypes.data, a.shape[0], global_index0,
                                         pv.ctypes.data if want_pv else None, None))
        return pv

    def encode_indices_host(self, retrieval_params, seed, cipher_index=0, n_cipher=1):
        out = np.empty((n_cipher, 2, N2), np.uint64)
        rp = retrieval_params.to_c()
        self._ck(self.L.omr_encode_indices(self.h, C.byref(rp), seed, cipher_index, n_cipher, out.ctypes.data))
        return out

    def encode_payloads_seeded_host(self, payloads, seed, all_payloads_count, combination_count, cmb_count_per_cipher):
        payloads = np.ascontiguousarray(payloads, np.uint16).reshape(-1, PAYLOAD_LENGTH)
        n_cipher = -(-combination_count // cmb_count_per_cipher)
        out = np.empty((n_cipher, 2, N2), np.uint64)
        self._ck(self.L.omr_encode_payloads_seeded(self.h, payloads.ctypes.data, payloads.shape[0], bytes(seed), all_payloads_count,
                                                   combination_count, cmb_count_per_cipher, out.ctypes.data))
        return out

    def encode_payloads_host(self, payloads, weights, combination_count, cmb_count_per_cipher):
        """weights [rows][D] with combination_count <= rows <= ceil(cc / per) * per: the natural [combination_count][D] matrix
        (what Retriever._weights returns) is accepted, the library zero-fills the missing tail rows (detector.rs:370-371)."""
        payloads = np.ascontiguousarray(payloads, np.uint16).reshape(-1, PAYLOAD_LENGTH); weights = np.ascontiguousarray(weights, np.uint16)
        n_cipher = -(-combination_count // cmb_count_per_cipher)
        if weights.ndim != 2 or not (combination_count <= weights.shape[0] <= n_cipher * cmb_count_per_cipher):
            raise OmrError(_lib.OMR_ERR_INVALID, "weights must be [rows][D] with combination_count <= rows <= n_cipher * cmb_count_per_cipher")
        out = np.empty((n_cipher, 2, N2), np.uint64)
        self._ck(self.L.omr_encode_payloads(self.h, payloads.ctypes.data, payloads.shape[0], weights.ctypes.data, weights.shape[0], weights.shape[1],
                                            n_cipher, cmb_count_per_cipher, out.ctypes.data))
        return out

    # -- stages (benches/two_level_bs.rs:47-145) on CUDA tensors ---------------------------------------------------------
    def first_level_blind_rotate(self, da, db):
        torch = _torch(); B = da.shape[0]
        out = torch.empty((B, 2, N1), dtype=torch.int32, device=da.device)
        self._ck(self.L.omr_l1_blind_rotate_device(self.h, da.data_ptr(), db.data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    def key_switch(self, rlwe):
        torch = _torch(); B = rlwe.shape[0]
        out = torch.empty((B, LWE2_N + 1), dtype=torch.int32, device=rlwe.device)
        self._ck(self.L.omr_keyswitch_device(self.h, rlwe.data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    def second_level_blind_rotate(self, lwe):
        torch = _torch(); B = lwe.shape[0]
        out = torch.empty((B, 2, N2), dtype=torch.int64, device=lwe.device)
        self._ck(self.L.omr_l2_blind_rotate_device(self.h, lwe.data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    def trace(self, rlwe):
        self._ck(self.L.omr_trace_device(self.h, rlwe.data_ptr(), rlwe.shape[0], self._stream()))
        return rlwe

    def mulmod_peak(self, level, iters=20000):
        """Measured peak of the register-only Shoup butterfly loop (mulmods/s) — the integer roofline denominator."""
        v = C.c_double(0.0)
        self._ck(self.L.omr_mulmod_peak(self.h, level, iters, C.byref(v)))
        return v.value

    def ntt(self, level, data, inverse=False):
        fn = self.L.omr_ntt_inverse_device if inverse else self.L.omr_ntt_forward_device
        self._ck(fn(self.h, level, data.data_ptr(), data.shape[0], self._stream()))
        return data

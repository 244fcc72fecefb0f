"""ctypes binding of the CPU oracle (oracle/libomr_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (tfhe_omr_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ORACLE_DIR = os.path.join(_ROOT, "oracle")

N1, N2, CLUE_N, CLUE_COUNT, LWE2_N, KSK_STRIDE, PAYLOAD_LEN = 1024, 2048, 512, 7, 670, 671, 612
Q1 = 134215681
Q2 = 1125899906826241
P = 257
BSK1_SHAPE = (512, 8, 2, 1024)
KSK_SHAPE = (1024, 27, 671)
BSK2_SHAPE = (670, 12, 2, 2048)
TRK_SHAPE = (11, 25, 2, 2048)


def build(native=False):
    """Compile the oracle.  native=True builds a -march=native copy (for CPU-baseline timing on the box)."""
    if native:
        out = os.path.join(_ORACLE_DIR, "libomr_oracle_native.so")
        src = os.path.join(_ORACLE_DIR, "omr_oracle_capi.cpp")
        if (not os.path.exists(out)) or os.path.getmtime(out) < max(
                os.path.getmtime(src), os.path.getmtime(os.path.join(_ORACLE_DIR, "omr_oracle.hpp"))):
            subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-pthread", "-shared",
                                   "-o", out, src], cwd=_ORACLE_DIR)
        return out
    subprocess.check_call(["make", "-s", "-C", _ORACLE_DIR])
    return os.path.join(_ORACLE_DIR, "libomr_oracle.so")


_lib_cache = {}


def lib(native=False):
    if native in _lib_cache:
        return _lib_cache[native]
    path = os.path.join(_ORACLE_DIR, "libomr_oracle_native.so" if native else "libomr_oracle.so")
    if not os.path.exists(path):
        path = build(native)
    L = C.CDLL(path)
    vp, u64, u32, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
    L.orc_const.restype = u64; L.orc_const.argtypes = [C.c_char_p]
    L.orc_keygen.restype = vp; L.orc_keygen.argtypes = [u64]
    L.orc_keygen_sender_only.restype = vp; L.orc_keygen_sender_only.argtypes = [u64]
    L.orc_key_from_blobs.restype = vp; L.orc_key_from_blobs.argtypes = [vp, vp, vp, vp]
    L.orc_keygen_cb.restype = vp; L.orc_keygen_cb.argtypes = [vp, C.c_char_p]
    L.orc_free.argtypes = [vp]
    for nm in ("orc_bsk1", "orc_ksk", "orc_bsk2", "orc_trk"):
        getattr(L, nm).restype = vp; getattr(L, nm).argtypes = [vp]
    L.orc_secret.argtypes = [vp, vp, vp, vp, vp]
    L.orc_gen_clues.argtypes = [vp, u64, u64, sz, vp, vp, i32]
    L.orc_gen_clue_msgs.argtypes = [vp, u64, u64, vp, vp, vp]
    L.orc_gen_clues_cb.argtypes = [vp, C.c_char_p, u64, sz, vp, vp, vp, i32]
    L.orc_clue_key.argtypes = [vp, vp, vp]
    L.orc_decrypt_clue.argtypes = [vp, vp, vp, vp]
    L.orc_detect.argtypes = [vp, vp, vp, sz, vp, i32]
    L.orc_l1.argtypes = [vp, vp, vp, sz, vp, i32]
    L.orc_keyswitch.argtypes = [vp, vp, sz, vp, i32]
    L.orc_l2.argtypes = [vp, vp, sz, vp, i32]
    L.orc_trace.argtypes = [vp, vp, sz, i32]
    L.orc_cmux1.argtypes = [vp, vp, C.c_uint, i32]
    L.orc_cmux2.argtypes = [vp, vp, C.c_uint, i32]
    L.orc_cmux1_simple.argtypes = [vp, vp, C.c_uint, i32]
    L.orc_cmux2_simple.argtypes = [vp, vp, C.c_uint, i32]
    L.orc_encode_indices.argtypes = [sz, i32, vp, sz, u64, u64, u32, vp]
    L.orc_encode_payloads.argtypes = [vp, vp, sz, u64, vp, sz, i32, i32, vp, i32]
    L.orc_decrypt_decode.argtypes = [vp, vp, vp]
    L.orc_decrypt_raw.argtypes = [vp, vp, vp]
    L.orc_decode_digest.restype = i32
    L.orc_decode_digest.argtypes = [vp, sz, i32, vp, i32, vp, vp, sz, vp, vp, vp]
    L.orc_phase_l1.restype = u32; L.orc_phase_l1.argtypes = [vp, vp]
    L.orc_phase_lwe2.restype = u32; L.orc_phase_lwe2.argtypes = [vp, vp]
    for nm in ("orc_ntt1_forward", "orc_ntt1_inverse", "orc_ntt2_forward", "orc_ntt2_inverse"):
        getattr(L, nm).argtypes = [vp, sz]
    L.orc_negacyclic1.argtypes = [vp, vp, vp]; L.orc_negacyclic2.argtypes = [vp, vp, vp]
    L.orc_decompose.argtypes = [i32, vp, sz, vp]
    L.orc_reduce128_q2.restype = u64; L.orc_reduce128_q2.argtypes = [u64, u64]
    L.orc_mod128_q2.restype = u64; L.orc_mod128_q2.argtypes = [u64, u64]
    L.orc_chacha_block.argtypes = [vp, u64, u64, i32, vp]
    L.orc_chacha12_weights.argtypes = [vp, vp, sz]
    L.orc_bucket_of.restype = u32; L.orc_bucket_of.argtypes = [u64, u32, u64, u32]
    L.orc_retrieval_params.argtypes = [sz, i32, vp]
    L.orc_lut1.argtypes = [vp]; L.orc_lut2.argtypes = [vp]
    L.orc_twiddles1.argtypes = [vp, vp]; L.orc_twiddles2.argtypes = [vp, vp]
    L.orc_blob_read.restype = i32; L.orc_blob_read.argtypes = [C.c_char_p, vp, vp, u64]
    L.orc_blob_write.restype = i32; L.orc_blob_write.argtypes = [C.c_char_p, u32, u64, u64, u64, u32, vp, u64]
    L.orc_inv_mod_257.restype = C.c_uint16; L.orc_inv_mod_257.argtypes = [i32]
    _lib_cache[native] = L
    return L


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def const(name, native=False):
    return int(lib(native).orc_const(name.encode()))


def _view(addr, shape, dtype):
    n = int(np.prod(shape))
    buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(addr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class KeyPack:
    """SecretKeyPack + ClueKey + DetectionKey of the oracle (key_gen/secret.rs)."""

    def __init__(self, seed=None, sender_only=False, blobs=None, native=False, cb_from=None, cb_seed=None):
        self.L = lib(native)
        if cb_from is not None:                       # same secrets / clue key, detection key from the counter-based generator
            assert len(cb_seed) == 32
            self.h = self.L.orc_keygen_cb(cb_from.h, bytes(cb_seed))
        elif blobs is not None:
            self._blobs = [np.ascontiguousarray(b) for b in blobs]
            self.h = self.L.orc_key_from_blobs(*[ptr(b) for b in self._blobs])
        elif sender_only:
            self.h = self.L.orc_keygen_sender_only(seed)
        else:
            self.h = self.L.orc_keygen(seed)
        self.sender_only = sender_only

    def __del__(self):
        try:
            self.L.orc_free(self.h)
        except Exception:
            pass

    # zero-copy views of the flat detection key (layouts: include/omr_b200.h)
    @property
    def bsk1(self): return _view(self.L.orc_bsk1(self.h), BSK1_SHAPE, np.uint32)
    @property
    def ksk(self): return _view(self.L.orc_ksk(self.h), KSK_SHAPE, np.uint32)
    @property
    def bsk2(self): return _view(self.L.orc_bsk2(self.h), BSK2_SHAPE, np.uint64)
    @property
    def trk(self): return _view(self.L.orc_trk(self.h), TRK_SHAPE, np.uint64)

    def secrets(self):
        s0 = np.zeros(512, np.int32); z1 = np.zeros(1024, np.int32); s2 = np.zeros(670, np.int32); z2 = np.zeros(2048, np.int32)
        self.L.orc_secret(self.h, ptr(s0), ptr(z1), ptr(s2), ptr(z2))
        return s0, z1, s2, z2

    def gen_clues(self, seed, count, index0=0, threads=8):
        a = np.zeros((count, CLUE_N), np.uint16); b = np.zeros((count, CLUE_COUNT), np.uint16)
        self.L.orc_gen_clues(self.h, seed, index0, count, ptr(a), ptr(b), threads)
        return a, b

    def gen_clues_cb(self, seed, count, index0=0, msgs=None, threads=8):
        """counter-based clue generation (bit-exact twin of the CUDA clue_gen_kernel); seed = 32 bytes (an int is widened
        little-endian)"""
        seed = seed.to_bytes(32, "little") if isinstance(seed, int) else bytes(seed)
        assert len(seed) == 32
        a = np.zeros((count, CLUE_N), np.uint16); b = np.zeros((count, CLUE_COUNT), np.uint16)
        m = None if msgs is None else np.ascontiguousarray(msgs, np.uint8).reshape(count, CLUE_COUNT)
        self.L.orc_gen_clues_cb(self.h, seed, index0, count, None if m is None else ptr(m), ptr(a), ptr(b), threads)
        return a, b

    def clue_key(self):
        pa = np.zeros(CLUE_N, np.uint16); pb = np.zeros(CLUE_N, np.uint16)
        self.L.orc_clue_key(self.h, ptr(pa), ptr(pb)); return pa, pb

    def gen_clue_msgs(self, seed, index, msgs):
        m = np.asarray(msgs, np.uint32); a = np.zeros(CLUE_N, np.uint16); b = np.zeros(CLUE_COUNT, np.uint16)
        self.L.orc_gen_clue_msgs(self.h, seed, index, ptr(m), ptr(a), ptr(b))
        return a, b

    def decrypt_clue(self, a, b):
        out = np.zeros(CLUE_COUNT, np.uint32)
        self.L.orc_decrypt_clue(self.h, ptr(np.ascontiguousarray(a)), ptr(np.ascontiguousarray(b)), ptr(out))
        return out

    def detect(self, a, b, threads=8):
        a = np.ascontiguousarray(a, np.uint16).reshape(-1, CLUE_N); b = np.ascontiguousarray(b, np.uint16).reshape(-1, CLUE_COUNT)
        pv = np.zeros((a.shape[0], 2, N2), np.uint64)
        self.L.orc_detect(self.h, ptr(a), ptr(b), a.shape[0], ptr(pv), threads)
        return pv

    def l1(self, a, b, threads=8):
        a = np.ascontiguousarray(a, np.uint16).reshape(-1, CLUE_N); b = np.ascontiguousarray(b, np.uint16).reshape(-1, CLUE_COUNT)
        out = np.zeros((a.shape[0], 2, N1), np.uint32)
        self.L.orc_l1(self.h, ptr(a), ptr(b), a.shape[0], ptr(out), threads)
        return out

    def keyswitch(self, rlwe, threads=8):
        rlwe = np.ascontiguousarray(rlwe, np.uint32).reshape(-1, 2, N1)
        out = np.zeros((rlwe.shape[0], KSK_STRIDE), np.uint32)
        self.L.orc_keyswitch(self.h, ptr(rlwe), rlwe.shape[0], ptr(out), threads)
        return out

    def l2(self, lwe, threads=8):
        lwe = np.ascontiguousarray(lwe, np.uint32).reshape(-1, KSK_STRIDE)
        out = np.zeros((lwe.shape[0], 2, N2), np.uint64)
        self.L.orc_l2(self.h, ptr(lwe), lwe.shape[0], ptr(out), threads)
        return out

    def trace(self, ct, threads=8):
        ct = np.array(ct, np.uint64).reshape(-1, 2, N2)
        self.L.orc_trace(self.h, ptr(ct), ct.shape[0], threads)
        return ct

    def cmux1(self, acc, a, key_index):
        acc = np.array(acc, np.uint32).reshape(2, N1); self.L.orc_cmux1(self.h, ptr(acc), a, key_index); return acc

    def cmux2(self, acc, a, key_index):
        acc = np.array(acc, np.uint64).reshape(2, N2); self.L.orc_cmux2(self.h, ptr(acc), a, key_index); return acc

    def cmux1_simple(self, acc, a, key_index):
        acc = np.array(acc, np.uint32).reshape(2, N1); self.L.orc_cmux1_simple(self.h, ptr(acc), a, key_index); return acc

    def cmux2_simple(self, acc, a, key_index):
        acc = np.array(acc, np.uint64).reshape(2, N2); self.L.orc_cmux2_simple(self.h, ptr(acc), a, key_index); return acc

    def decrypt_decode(self, ct):
        ct = np.ascontiguousarray(ct, np.uint64); out = np.zeros(N2, np.uint64)
        self.L.orc_decrypt_decode(self.h, ptr(ct), ptr(out)); return out

    def decrypt_raw(self, ct):
        ct = np.ascontiguousarray(ct, np.uint64); out = np.zeros(N2, np.uint64)
        self.L.orc_decrypt_raw(self.h, ptr(ct), ptr(out)); return out

    def phase_l1(self, rlwe):
        return int(self.L.orc_phase_l1(self.h, ptr(np.ascontiguousarray(rlwe, np.uint32))))

    def phase_lwe2(self, lwe):
        return int(self.L.orc_phase_lwe2(self.h, ptr(np.ascontiguousarray(lwe, np.uint32))))

    def decode_digest(self, all_payloads, pertinent, index_cts, payload_cts, weights):
        index_cts = np.ascontiguousarray(index_cts, np.uint64); payload_cts = np.ascontiguousarray(payload_cts, np.uint64)
        weights = np.ascontiguousarray(weights, np.uint16)
        idx = np.zeros(max(pertinent, 1) * 4 + 4096, np.uint64); n = C.c_int(0)
        pl = np.zeros((max(pertinent, 1) + 8, PAYLOAD_LEN), np.uint16)
        st = self.L.orc_decode_digest(self.h, all_payloads, pertinent, ptr(index_cts), index_cts.shape[0], ptr(payload_cts),
                                      ptr(weights), weights.shape[1], ptr(idx), C.byref(n), ptr(pl))
        return st, idx[:n.value].astype(np.int64), pl[:n.value]


def retrieval_params(all_payloads, pertinent):
    out = np.zeros(6, np.int32)
    lib().orc_retrieval_params(all_payloads, pertinent, ptr(out))
    keys = ["slots_per_bucket", "slots_per_segment", "segment_per_cipher", "max_encode_indices_cipher_count",
            "combination_count", "payload_cipher_count"]
    return dict(zip(keys, [int(v) for v in out]))


def encode_indices(all_payloads, pertinent, pv, index0, seed, cipher_idx):
    pv = np.ascontiguousarray(pv, np.uint64); out = np.zeros((2, N2), np.uint64)
    lib().orc_encode_indices(all_payloads, pertinent, ptr(pv), pv.shape[0], index0, seed, cipher_idx, ptr(out))
    return out


def encode_payloads(pv, payloads, index0, weights, n_cipher, cmb_per_cipher=2, threads=8):
    pv = np.ascontiguousarray(pv, np.uint64); payloads = np.ascontiguousarray(payloads, np.uint16)
    weights = np.ascontiguousarray(weights, np.uint16)
    out = np.zeros((n_cipher, 2, N2), np.uint64)
    lib().orc_encode_payloads(ptr(pv), ptr(payloads), pv.shape[0], index0, ptr(weights), weights.shape[1], n_cipher,
                              cmb_per_cipher, ptr(out), threads)
    return out


def chacha12_weights(seed32, count):
    seed = np.frombuffer(bytes(seed32), np.uint8).copy(); out = np.zeros(count, np.uint16)
    lib().orc_chacha12_weights(ptr(seed), ptr(out), count); return out


def random_key_blobs(seed):
    """Uniformly random 'detection key' blobs: detect is data-oblivious integer arithmetic, so garbage keys
    exercise exactly the same code path and make libm-independent bit-exact fixtures."""
    rng = np.random.Generator(np.random.PCG64(seed))
    bsk1 = rng.integers(0, Q1, BSK1_SHAPE, dtype=np.uint32)
    ksk = rng.integers(0, Q1, KSK_SHAPE, dtype=np.uint32)
    bsk2 = rng.integers(0, Q2, BSK2_SHAPE, dtype=np.uint64)
    trk = rng.integers(0, Q2, TRK_SHAPE, dtype=np.uint64)
    return bsk1, ksk, bsk2, trk

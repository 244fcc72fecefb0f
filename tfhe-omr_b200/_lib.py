"""ctypes binding of libomr_b200.so (include/omr_b200.h).  Fails loudly when the CUDA library is missing: there is
no CPU path in this package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libomr_b200.so")

OMR_OK, OMR_ERR_INVALID, OMR_ERR_CUDA, OMR_ERR_ALLOC, OMR_ERR_STATE = range(5)
KEYS_NTT_NATIVE, KEYS_COEFF = 0, 1
OUT_NTT_NATIVE, OUT_COEFF = 0, 1

# every symbol include/omr_b200.h declares (tests check the library exports each of them)
EXPORTS = [
    "omr_ctx_create", "omr_ctx_create_device_keys", "omr_ctx_destroy", "omr_last_error", "omr_detect_key_size",
    "omr_retrieval_params_init", "omr_detect_batch", "omr_pv_reset", "omr_encode_indices", "omr_encode_payloads",
    "omr_detect_batch_device", "omr_encode_indices_device", "omr_encode_payloads_device", "omr_digest_reduce_mod",
    "omr_l1_blind_rotate_device", "omr_keyswitch_device", "omr_l2_blind_rotate_device", "omr_trace_device",
    "omr_ntt_forward_device", "omr_ntt_inverse_device", "omr_launch_count", "omr_mulmod_peak", "omr_digest_add_mod", "omr_decrypt_decode_device", "omr_gen_clues_device", "omr_set_latency_shapes", "omr_decode_digest", "omr_weights_from_seed_device", "omr_encode_payloads_seeded", "omr_set_tensor_core_key_switch",
    "omr_blob_field_count", "omr_blob_field_bytes", "omr_blob_write", "omr_blob_read_header", "omr_blob_read", "omr_ctx_create_from_blob",
    "omr_stream_begin", "omr_stream_push", "omr_stream_snapshot", "omr_stream_end",
    "omr_comm_unique_id", "omr_comm_init", "omr_comm_destroy", "omr_digest_allreduce",
    "omr_generate_detector", "omr_pv_load",
    "omr_first_level_lut", "omr_second_level_lut", "omr_set_output_domain", "omr_key_switch_path",
]


class KeyBlobs(C.Structure):
    _fields_ = [("bsk1", C.c_void_p), ("ksk", C.c_void_p), ("bsk2", C.c_void_p), ("trace", C.c_void_p), ("flags", C.c_uint32)]


class SecretKey(C.Structure):
    _fields_ = [("s0", C.c_void_p), ("z1", C.c_void_p), ("s2", C.c_void_p), ("z2", C.c_void_p)]


class StageTimes(C.Structure):
    _fields_ = [("detect_ms", C.c_float), ("first_level_bootstrapping_ms", C.c_float),
                ("second_level_bootstrapping_ms", C.c_float), ("trace_ms", C.c_float)]


class BlobHeader(C.Structure):
    _fields_ = [("version", C.c_uint32), ("kind", C.c_uint32), ("count", C.c_uint64), ("index0", C.c_uint64), ("aux", C.c_uint64),
                ("payload_bytes", C.c_uint64), ("domain", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class RetrievalParamsC(C.Structure):
    _fields_ = [("index_modulus", C.c_uint64)] + [(n, C.c_uint32) for n in (
        "polynomial_size", "bucket_count_per_segment", "slots_per_bucket", "slots_per_segment", "segment_count",
        "segment_per_cipher", "max_encode_indices_cipher_count", "pertinent_count", "combination_count",
        "cmb_count_per_cipher")] + [("all_payloads_count", C.c_uint64)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(tfhe_omr_b200 has no CPU fallback)")
    L = C.CDLL(os.environ.get("OMR_B200_LIB", LIB_PATH))       # override: A/B runs of experimental builds (scripts/)
    vp, u64, u32, i32, sz = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_size_t
    P = C.POINTER
    L.omr_ctx_create.restype = i32; L.omr_ctx_create.argtypes = [i32, P(KeyBlobs), P(vp)]
    L.omr_ctx_create_device_keys.restype = i32; L.omr_ctx_create_device_keys.argtypes = [i32, P(KeyBlobs), P(vp)]
    L.omr_ctx_destroy.restype = None; L.omr_ctx_destroy.argtypes = [vp]
    L.omr_last_error.restype = C.c_char_p; L.omr_last_error.argtypes = [vp]
    L.omr_detect_key_size.restype = sz; L.omr_detect_key_size.argtypes = [vp]
    L.omr_launch_count.restype = u64; L.omr_launch_count.argtypes = [vp]
    L.omr_set_latency_shapes.restype = i32; L.omr_set_latency_shapes.argtypes = [vp, i32]
    L.omr_set_tensor_core_key_switch.restype = i32; L.omr_set_tensor_core_key_switch.argtypes = [vp, i32]
    L.omr_weights_from_seed_device.restype = i32; L.omr_weights_from_seed_device.argtypes = [vp, C.c_char_p, sz, vp, u32, vp]
    L.omr_encode_payloads_seeded.restype = i32; L.omr_encode_payloads_seeded.argtypes = [vp, vp, sz, C.c_char_p, u64, u32, u32, vp]
    L.omr_decode_digest.restype = i32
    L.omr_decode_digest.argtypes = [vp, P(RetrievalParamsC), vp, vp, u32, vp, u32, vp, sz, vp, P(u32), vp]
    L.omr_retrieval_params_init.restype = i32; L.omr_retrieval_params_init.argtypes = [u64, u32, P(RetrievalParamsC)]
    L.omr_detect_batch.restype = i32; L.omr_detect_batch.argtypes = [vp, vp, vp, sz, u64, vp, P(StageTimes)]
    L.omr_pv_reset.restype = i32; L.omr_pv_reset.argtypes = [vp]
    L.omr_pv_load.restype = i32; L.omr_pv_load.argtypes = [vp, vp, sz, u64]
    L.omr_encode_indices.restype = i32; L.omr_encode_indices.argtypes = [vp, P(RetrievalParamsC), u64, u32, u32, vp]
    L.omr_encode_payloads.restype = i32; L.omr_encode_payloads.argtypes = [vp, vp, sz, vp, sz, sz, u32, u32, vp]
    L.omr_detect_batch_device.restype = i32; L.omr_detect_batch_device.argtypes = [vp, vp, vp, sz, vp, vp, P(StageTimes)]
    L.omr_encode_indices_device.restype = i32
    L.omr_encode_indices_device.argtypes = [vp, P(RetrievalParamsC), vp, sz, u64, u64, u32, u32, vp, vp]
    L.omr_encode_payloads_device.restype = i32
    L.omr_encode_payloads_device.argtypes = [vp, vp, vp, sz, u64, vp, sz, u32, u32, vp, vp]
    L.omr_digest_reduce_mod.restype = i32; L.omr_digest_reduce_mod.argtypes = [vp, vp, sz, vp]
    L.omr_l1_blind_rotate_device.restype = i32; L.omr_l1_blind_rotate_device.argtypes = [vp, vp, vp, sz, vp, vp]
    L.omr_keyswitch_device.restype = i32; L.omr_keyswitch_device.argtypes = [vp, vp, sz, vp, vp]
    L.omr_l2_blind_rotate_device.restype = i32; L.omr_l2_blind_rotate_device.argtypes = [vp, vp, sz, vp, vp]
    L.omr_trace_device.restype = i32; L.omr_trace_device.argtypes = [vp, vp, sz, vp]
    L.omr_ntt_forward_device.restype = i32; L.omr_ntt_forward_device.argtypes = [vp, i32, vp, sz, vp]
    L.omr_ntt_inverse_device.restype = i32; L.omr_ntt_inverse_device.argtypes = [vp, i32, vp, sz, vp]
    L.omr_digest_add_mod.restype = i32; L.omr_digest_add_mod.argtypes = [vp, vp, vp, sz, vp]
    L.omr_decrypt_decode_device.restype = i32; L.omr_decrypt_decode_device.argtypes = [vp, vp, vp, sz, vp, vp]
    L.omr_gen_clues_device.restype = i32; L.omr_gen_clues_device.argtypes = [vp, vp, vp, C.c_char_p, u64, sz, vp, vp, vp, vp]
    L.omr_blob_field_count.restype = u32; L.omr_blob_field_count.argtypes = [u32]
    L.omr_blob_field_bytes.restype = sz; L.omr_blob_field_bytes.argtypes = [u32, u32, u64]
    L.omr_blob_write.restype = i32; L.omr_blob_write.argtypes = [C.c_char_p, u32, u64, u64, u64, u32, P(vp), u32]
    L.omr_blob_read_header.restype = i32; L.omr_blob_read_header.argtypes = [C.c_char_p, P(BlobHeader)]
    L.omr_blob_read.restype = i32; L.omr_blob_read.argtypes = [C.c_char_p, P(BlobHeader), P(vp), u32]
    L.omr_ctx_create_from_blob.restype = i32; L.omr_ctx_create_from_blob.argtypes = [i32, C.c_char_p, P(vp)]
    L.omr_stream_begin.restype = i32; L.omr_stream_begin.argtypes = [vp, P(RetrievalParamsC), u64, C.c_char_p, u64]
    L.omr_stream_push.restype = i32; L.omr_stream_push.argtypes = [vp, vp, vp, vp, sz]
    L.omr_stream_snapshot.restype = i32; L.omr_stream_snapshot.argtypes = [vp, vp, P(u64)]
    L.omr_stream_end.restype = i32; L.omr_stream_end.argtypes = [vp]
    L.omr_comm_unique_id.restype = i32; L.omr_comm_unique_id.argtypes = [vp]
    L.omr_comm_init.restype = i32; L.omr_comm_init.argtypes = [vp, i32, i32, vp]
    L.omr_comm_destroy.restype = i32; L.omr_comm_destroy.argtypes = [vp]
    L.omr_digest_allreduce.restype = i32; L.omr_digest_allreduce.argtypes = [vp, vp, vp, sz, vp]
    L.omr_generate_detector.restype = i32; L.omr_generate_detector.argtypes = [i32, P(SecretKey), C.c_char_p, P(KeyBlobs), P(vp)]
    L.omr_first_level_lut.restype = i32; L.omr_first_level_lut.argtypes = [vp, vp]
    L.omr_second_level_lut.restype = i32; L.omr_second_level_lut.argtypes = [vp, vp]
    L.omr_set_output_domain.restype = i32; L.omr_set_output_domain.argtypes = [vp, u32]
    L.omr_key_switch_path.restype = i32; L.omr_key_switch_path.argtypes = [vp]
    L.omr_mulmod_peak.restype = i32; L.omr_mulmod_peak.argtypes = [vp, i32, i32, P(C.c_double)]
    _lib = L
    return L

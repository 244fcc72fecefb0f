/*
 * omr_b200.h — C ABI of libomr_b200.so: the B200 (sm_100a) implementation of InstantOMR's detection hot path.
 *
 * Drop-in boundary.  The reference (xiangxiecrypto/tfhe-omr, crate omr_core) has no FFI; the boundary is the
 * Rust public API of `Detector` (omr_core/src/lib.rs:21-31).  Each entry point below names the reference
 * interface it replaces.  All buffers are flat, little-endian, caller-owned; every call returns an omr_status
 * and never unwinds.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * OMR_ERR_CUDA.
 *
 * Layout conventions (Primus-fhe's own in-memory layouts are not available — "parity
 * unpinned" — so the shim flattens into these):
 *   negacyclic NTT: forward = Cooley-Tukey, natural order in, bit-reversed order out, out[k] = a(psi^(2*brv(k)+1)),
 *     psi1 = 4073518 (mod q1 = 134215681, N1 = 1024), psi2 = 765727830662934 (mod q2 = 1125899906826241, N2 = 2048).
 *   RLWE (a, b): b = a*s + m + e.   RGSW rows: [0,L) = RLWE(-s*m*g_j), [L,2L) = RLWE(m*g_j), g_j = 2^(drop + w*j).
 *
 * Concurrency.  The reference shares one `&Detector` between rayon workers (examples/omr.rs:160-164); here the batch IS the
 * parallelism.  The host-buffer calls (omr_detect_batch, omr_encode_*, omr_stream_*, omr_decode_digest) hold the context's
 * mutex from the first launch until their result has been copied out, and may be called from any thread.  The *_device calls
 * use scratch buffers owned by the context: calls on one context must be ordered on ONE stream (or externally serialised);
 * use one context per stream — or per recipient key — for concurrent work on a GPU.  Every call runs on the context's device
 * and restores the caller's current CUDA device before returning.
 */
#ifndef OMR_B200_H
#define OMR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMR_CLUE_N 512        /* clue LWE dimension            parameters/mod.rs:41 */
#define OMR_CLUE_COUNT 7      /* clues per message             parameters/mod.rs:48 */
#define OMR_N1 1024           /* first-level ring dimension    parameters/mod.rs:51 */
#define OMR_Q1 134215681u     /* FirstLevelField               parameters/mod.rs:18 */
#define OMR_LWE2_N 670        /* intermediate LWE dimension    parameters/mod.rs:69 */
#define OMR_N2 2048           /* second-level ring dimension   parameters/mod.rs:77 */
#define OMR_Q2 1125899906826241ull /* SecondLevelField         parameters/mod.rs:21 */
#define OMR_P 257             /* output plain modulus          parameters/mod.rs:93 */
#define OMR_PAYLOAD_LEN 612   /* payload.rs:8 */
#define OMR_PV_WORDS (2 * OMR_N2) /* one pertinency ciphertext = NttRlweCiphertext<SecondLevelField> */

typedef enum {
    OMR_OK = 0,
    OMR_ERR_INVALID = 1,   /* bad argument (the reference panics/asserts: detector.rs:236,511) */
    OMR_ERR_CUDA = 2,      /* CUDA runtime failure or no device — never falls back to the CPU */
    OMR_ERR_ALLOC = 3,
    OMR_ERR_STATE = 4      /* call order (e.g. encode before any detect) */
} omr_status;

/* key-blob flags */
#define OMR_KEYS_NTT_NATIVE 0u /* ring polynomials already in this library's NTT convention (above) */
#define OMR_KEYS_COEFF 1u      /* ring polynomials in coefficient form; transformed on upload (Rust-shim path) */

/* Domain of the ring polynomials that cross the HOST-buffer boundary after the keys (omr_set_output_domain): the pertinency
 * ciphertexts of omr_detect_batch, the digests of omr_encode_indices / omr_encode_payloads*, the running digest of
 * omr_stream_snapshot, and the ciphertexts / secret handed to omr_decode_digest.  OMR_OUT_COEFF mitigates SURVEY §8b risk 1 on the
 * output side: the Rust shim re-transforms with Primus-fhe's own NTT table (NttRlwe <-> Rlwe), so nothing depends on the two
 * libraries agreeing on the root of unity or on the output ordering.  Device-pointer calls are always NTT-native. */
#define OMR_OUT_NTT_NATIVE 0u
#define OMR_OUT_COEFF 1u

/* Replaces DetectionKey (omr_core/src/key_gen/detection.rs:9-16) flattened:
 *   BlindRotationKey<F1> -> bsk1, NonPowOf2LweKeySwitchingKey -> ksk, BlindRotationKey<F2> -> bsk2,
 *   TraceKey<F2> -> trace.  N2^-1 and both LUTs (detector.rs:457-503) are derived inside. */
typedef struct {
    const uint32_t* bsk1;  /* [512][8][2][1024]   mod q1 */
    const uint32_t* ksk;   /* [1024][27][671]     mod q1, (a[670], b) */
    const uint64_t* bsk2;  /* [670][12][2][2048]  mod q2 */
    const uint64_t* trace; /* [11][25][2][2048]   mod q2, step t <-> automorphism X -> X^(2^(11-t)+1) */
    uint32_t flags;
} omr_key_blobs;

/* Replaces DetectTimeInfoPerMessage / DetectTimeInfo (detector.rs:43-57): device time per stage, milliseconds,
 * summed over the batch of the call. */
typedef struct {
    float detect_ms;
    float first_level_bootstrapping_ms;  /* L1 blind rotations + key switch + modulus switch */
    float second_level_bootstrapping_ms;
    float trace_ms;
} omr_stage_times;

/* Replaces RetrievalParams<F> (parameters/retrieval_params.rs:11-46); fill with omr_retrieval_params_init. */
typedef struct {
    uint64_t index_modulus;
    uint32_t polynomial_size, bucket_count_per_segment, slots_per_bucket, slots_per_segment, segment_count,
        segment_per_cipher, max_encode_indices_cipher_count, pertinent_count, combination_count, cmb_count_per_cipher;
    uint64_t all_payloads_count;
} omr_retrieval_params;

typedef struct omr_ctx omr_ctx;

/* ---- lifetime -------------------------------------------------------------------------------------------- */
/* Replaces Detector::new(DetectionKey) (detector.rs:85-110): uploads and owns device copies of the keys. */
int omr_ctx_create(int device, const omr_key_blobs* keys, omr_ctx** out);
/* Keys already resident on `device` (same layouts); copied once more into the context's internal form.  The call drains the
 * device (cudaDeviceSynchronize) before the first copy, so the key tensors may have been produced on any stream. */
int omr_ctx_create_device_keys(int device, const omr_key_blobs* device_keys, omr_ctx** out);
/* SecretKeyPack::generate_detector (key_gen/secret.rs:118-187: generate_detection_key + Detector::new) with the key material
 * made ON THE GPU: 512 + 670 RGSW encryptions, 27 648 LWE key-switching rows and 275 trace-key RLWE rows from the
 * recipient's secrets (host arrays: s0 binary [512], z1 ternary [1024], s2 binary [670], z2 ternary [2048]) and a 32-byte seed
 * that keys the ChaCha12 stream every mask and error is drawn from (the reference takes `R: Rng + CryptoRng`; csrc/keygen.cuh
 * states the stream layout and the Gaussian tables).  The context is ready for detection; when host_keys_out != NULL its four
 * buffers (layouts of omr_key_blobs, NTT-native) also receive the flat detection key — what the recipient ships to a detector. */
typedef struct { const int32_t* s0; const int32_t* z1; const int32_t* s2; const int32_t* z2; } omr_secret_key;
int omr_generate_detector(int device, const omr_secret_key* sk, const uint8_t* seed32, const omr_key_blobs* host_keys_out, omr_ctx** out);
void omr_ctx_destroy(omr_ctx* ctx);
const char* omr_last_error(const omr_ctx* ctx); /* NULL ctx -> error of the last failed create */
/* Detector::detect_key_size (detector.rs:112-114): bytes of key material resident on the device */
size_t omr_detect_key_size(const omr_ctx* ctx);
/* Detector::first_level_lut / second_level_lut (detector.rs:117-132; built at :457-503 with lut.rs:12-27): the two test vectors
 * in coefficient form, copied to the host. */
int omr_first_level_lut(const omr_ctx* ctx, uint32_t* out /*[1024] mod q1*/);
int omr_second_level_lut(const omr_ctx* ctx, uint64_t* out /*[2048] mod q2*/);
/* OMR_OUT_NTT_NATIVE (default) or OMR_OUT_COEFF for every later host-buffer call on this context. */
int omr_set_output_domain(omr_ctx* ctx, uint32_t domain);
/* RetrievalParams::new(257, 2048, D, k, 130, 25, 2) — secret.rs:189-209 */
int omr_retrieval_params_init(uint64_t all_payloads_count, uint32_t pertinent_count, omr_retrieval_params* out);

/* ---- the hot path, host buffers (what the Rust shim binds) ---------------------------------------------------- */
/* Replaces `clues_list.par_iter().map(|c| detector.detect(c))` (examples/omr.rs:160-164 over detector.rs:135-166)
 * for B messages.  clue_a [B][512], clue_b [B][7] (CmLweCiphertext<u16> flattened).  The B pertinency ciphertexts
 * are kept in the context's device-resident pertinency store at positions [global_index0, global_index0+B)
 * (the store grows on demand) and, when pv_out != NULL, also copied to the host as [B][2][2048].
 * times may be NULL (detect_with_time_info, detector.rs:169-221). */
int omr_detect_batch(omr_ctx* ctx, const uint16_t* clue_a, const uint16_t* clue_b, size_t B, uint64_t global_index0,
                     uint64_t* pv_out, omr_stage_times* times);
/* Forget the pertinency store (start a new bulletin board). */
int omr_pv_reset(omr_ctx* ctx);
/* Replace the store by pertinency ciphertexts the caller holds ([count][2][2048], in the context's output domain, global
 * indices global_index0..) — what lets Detector::encode_pertinent_indices / encode_pertinent_payloads keep the reference's
 * signature, which takes the pertinency vector as an argument (detector.rs:223-227, 341-351). */
int omr_pv_load(omr_ctx* ctx, const uint64_t* pv, size_t count, uint64_t global_index0);
/* Replaces Detector::encode_pertinent_indices (detector.rs:223-339) over the resident pertinency store, for
 * ciphertexts [cipher_idx0, cipher_idx0+n_cipher).  The reference draws buckets from thread_rng (detector.rs:262);
 * here they are a counter hash of (seed, cipher, global message index, segment), see DESIGN.md.
 * out [n_cipher][2][2048] receives this context's PARTIAL digest already reduced mod q2. */
int omr_encode_indices(omr_ctx* ctx, const omr_retrieval_params* rp, uint64_t seed, uint32_t cipher_idx0,
                       uint32_t n_cipher, uint64_t* out);
/* Replaces Detector::encode_pertinent_payloads (detector.rs:341-453).  payloads [count][612] for the messages in
 * the store (count must equal the store size, global indices index0..), weights [weight_rows][weight_stride] row-major
 * u16 in [0,257) with column = GLOBAL message index (the caller draws them: StdRng + Uniform, detector.rs:376-387),
 * 1 <= weight_rows <= n_cipher*cmb_per_cipher; rows beyond weight_rows count as zero (the reference allocates
 * ceil(cc / per) * per rows and fills combination_count of them, detector.rs:370-387).  out [n_cipher][2][2048]. */
int omr_encode_payloads(omr_ctx* ctx, const uint16_t* payloads, size_t count, const uint16_t* weights, size_t weight_rows,
                        size_t weight_stride, uint32_t n_cipher, uint32_t cmb_per_cipher, uint64_t* out);

/* ---- device-pointer forms (inputs already in HBM; `stream` is a cudaStream_t, NULL = the default stream) -------------------------- */
int omr_detect_batch_device(omr_ctx* ctx, const uint16_t* d_clue_a, const uint16_t* d_clue_b, size_t B,
                            uint64_t* d_pv /*[B][2][2048]*/, void* stream, omr_stage_times* times);
int omr_encode_indices_device(omr_ctx* ctx, const omr_retrieval_params* rp, const uint64_t* d_pv, size_t count,
                              uint64_t index0, uint64_t seed, uint32_t cipher_idx0, uint32_t n_cipher,
                              uint64_t* d_out, void* stream);
int omr_encode_payloads_device(omr_ctx* ctx, const uint64_t* d_pv, const uint16_t* d_payloads, size_t count,
                               uint64_t index0, const uint16_t* d_weights, size_t weight_stride, uint32_t n_cipher,
                               uint32_t cmb_per_cipher, uint64_t* d_out, void* stream);
/* After a cross-GPU integer sum of partial digests (rayon reduce + add_element_wise, detector.rs:333-336,445-448):
 * reduce every word mod q2 in place (inputs < 2^63). */
int omr_digest_reduce_mod(omr_ctx* ctx, uint64_t* d_words, size_t n_words, void* stream);

/* Streaming detection (README.md:9 "processes incoming messages on-the-fly"): fold the digest of a newly detected batch
 * into a running digest, acc = (acc + part) mod q2, both canonical.  Packing is a sum over messages, so a running digest
 * built batch by batch is bit-identical to packing the whole board at once. */
int omr_digest_add_mod(omr_ctx* ctx, uint64_t* d_acc, const uint64_t* d_part, size_t n_words, void* stream);

/* The same as an ingest loop with host buffers.  omr_stream_begin fixes the retrieval layout, the bucket seed of
 * the index digest, the 32-byte rng seed of the combination weights (detector.rs:376-387) and the global index of the first
 * message; omr_stream_push(clues, payloads, n) — any n, any number of times — detects the n next messages and folds their
 * contributions into a RESIDENT running digest, stream-ordered behind double-buffered pinned staging (the call returns when the
 * inputs have been staged, not when the GPU is done); omr_stream_snapshot waits for everything pushed so far and copies the
 * digest out, [max_encode_indices_cipher_count + ceil(combination_count / cmb_count_per_cipher)][2][2048], in the context's
 * output domain.  Identical, word for word, to omr_detect_batch + omr_encode_indices + omr_encode_payloads_seeded over the same
 * messages in one shot.  One stream per context; it does not touch the pertinency store of omr_detect_batch. */
int omr_stream_begin(omr_ctx* ctx, const omr_retrieval_params* rp, uint64_t index_seed, const uint8_t* weight_seed32, uint64_t global_index0);
int omr_stream_push(omr_ctx* ctx, const uint16_t* clue_a /*[n][512]*/, const uint16_t* clue_b /*[n][7]*/, const uint16_t* payloads /*[n][612]*/, size_t n);
int omr_stream_snapshot(omr_ctx* ctx, uint64_t* out, uint64_t* n_messages /*nullable: messages folded in so far*/);
int omr_stream_end(omr_ctx* ctx);

/* K7 — the only collective of the path: the sum over GPUs of the partial digests (the rayon
 * reduce(add_element_wise) of detector.rs:333-336, 445-448), in place on d_digests [n_cipher][2][2048] (canonical partial
 * digests in, their sum mod q2 out), ncclAllReduce(ncclUint64, ncclSum) + one mod-q2 kernel on `stream`.  nccl_comm is the
 * caller's ncclComm_t (e.g. torch's), or NULL to use the context's own communicator: rank 0 draws an id with
 * omr_comm_unique_id, hands the 128 bytes to every rank out of band, and every rank calls omr_comm_init on its context.
 * libnccl is resolved with dlopen at first use (an already-loaded copy is preferred; OMR_NCCL_LIB overrides the name), so a
 * single-GPU deployment does not need it; without it these calls return OMR_ERR_STATE. */
#define OMR_COMM_ID_BYTES 128
int omr_comm_unique_id(uint8_t* id128);
int omr_comm_init(omr_ctx* ctx, int n_ranks, int rank, const uint8_t* id128);
int omr_comm_destroy(omr_ctx* ctx);
int omr_digest_allreduce(omr_ctx* ctx, void* nccl_comm, uint64_t* d_digests, size_t n_cipher, void* stream);

/* Recipient side (Retriever::decode_pertinent_indices / decode_combined_payloads, retriever.rs:63-130,
 * 318-362): decrypt n NTT-domain RLWE ciphertexts with the recipient's NTT-domain secret z2 (b - a*z2, inverse NTT) and
 * decode every coefficient c to round_half_up(c * 257 / q2) folded into [0,257) — exact integer rounding instead of the
 * reference's BigDecimal.  d_out [n][2048] u16.  The bucket scan and the mod-257 solver run on the host
 * (tfhe_omr_b200.retriever). */
int omr_decrypt_decode_device(omr_ctx* ctx, const uint64_t* d_z2_ntt /*[2048]*/, const uint64_t* d_ct /*[n][2][2048]*/, size_t n,
                              uint16_t* d_out /*[n][2048]*/, void* stream);

/* Retriever::decode_digest (retriever.rs:188-260), host buffers: decrypt and decode the index and payload ciphertexts on
 * the GPU, scan the buckets (a bucket counts when its flag slot is exactly 1), look the found columns up in `weights`
 * [combination_count][weight_stride] (the matrix the reference regenerates from its 32-byte seed, :215-226) and solve the
 * system mod 257 (solve_matrix_mod_257, matrix.rs:164-247).  indices_out [pertinent_count] (sorted), *n_found, payloads_out
 * [pertinent_count][612].  A singular system returns OMR_ERR_INVALID with "matrix is not invertible" (OmrError::
 * InvertibleMatrix, error.rs:4-8). */
int omr_decode_digest(omr_ctx* ctx, const omr_retrieval_params* rp, const uint64_t* z2 /*[2048]: NTT(z2), or z2 mod q2 in coefficient form under OMR_OUT_COEFF (like the ciphertexts)*/,
                      const uint64_t* index_cts /*[n_index_cts][2][2048]*/, uint32_t n_index_cts,
                      const uint64_t* payload_cts /*[n_payload_cts][2][2048]*/, uint32_t n_payload_cts,
                      const uint16_t* weights, size_t weight_stride,
                      uint64_t* indices_out, uint32_t* n_found, uint16_t* payloads_out);

/* The combination weights exactly as the reference draws them (detector.rs:376-387, regenerated by the recipient at
 * retriever.rs:215-226): StdRng::from_seed(seed32) (rand 0.8: ChaCha12) feeding Uniform::<u16>::new(0, 257), `count` draws in
 * stream order into d_out (row-major [combination_count][all_payloads_count] when count is their product).  With this the
 * 32-byte seed of the reference's API is all that crosses the boundary.  flags bit 0: generate strictly in order on one
 * thread (the path taken when a draw is rejected, probability 2^-32 per draw; exposed for the tests). */
int omr_weights_from_seed_device(omr_ctx* ctx, const uint8_t* seed32, size_t count, uint16_t* d_out, uint32_t flags, void* stream);

/* omr_encode_payloads with the reference's own argument — the seed of the rng it is handed (detector.rs:341, 376-387) —
 * instead of a weight matrix: weights are generated on the GPU (omr_weights_from_seed_device), rows beyond combination_count
 * are zero.  out [ceil(combination_count / cmb_per_cipher)][2][2048]. */
int omr_encode_payloads_seeded(omr_ctx* ctx, const uint16_t* payloads, size_t count, const uint8_t* seed32, uint64_t all_payloads_count,
                               uint32_t combination_count, uint32_t cmb_per_cipher, uint64_t* out);

/* Sender side (Sender::gen_clues -> ClueKey::gen_clues, sender.rs:27-30, key_gen/clue.rs:27-34): `count` clues
 * under the clue public key (pa, pb) [512] u16 each, for global message indices index0.., encrypting d_msgs[i][7] (values
 * mod 8; NULL = seven 0's as the reference does).  The reference takes `R: Rng + CryptoRng`; here every draw (the binary mask
 * r, the errors e1, e2) comes from ChaCha12 keyed by seed32 (host pointer, 32 bytes from the caller's CSPRNG) in counter mode
 * over (message index, domain, block) — see kernels.cuh: clue_gen_kernel.  Never reuse a seed for the same index range.
 * d_a [count][512], d_b [count][7]. */
int omr_gen_clues_device(omr_ctx* ctx, const uint16_t* d_pa, const uint16_t* d_pb, const uint8_t* seed32, uint64_t index0, size_t count,
                         const uint8_t* d_msgs, uint16_t* d_a, uint16_t* d_b, void* stream);

/* ---- stage entry points (device pointers) — the stage list of benches/two_level_bs.rs:47-145 ------------------- */
int omr_l1_blind_rotate_device(omr_ctx* ctx, const uint16_t* d_clue_a, const uint16_t* d_clue_b, size_t B,
                               uint32_t* d_rlwe /*[B][2][1024] sum of the 7 accumulators*/, void* stream);
int omr_keyswitch_device(omr_ctx* ctx, const uint32_t* d_rlwe /*[B][2][1024]*/, size_t B,
                         uint32_t* d_lwe /*[B][671] mod 4096, offset added*/, void* stream);
int omr_l2_blind_rotate_device(omr_ctx* ctx, const uint32_t* d_lwe /*[B][671]*/, size_t B,
                               uint64_t* d_rlwe /*[B][2][2048]*/, void* stream);
int omr_trace_device(omr_ctx* ctx, uint64_t* d_rlwe /*[B][2][2048], in place -> NttRlwe*/, size_t B, void* stream);
/* batched negacyclic NTTs, in place, canonical outputs; level 1: u32 [batch][1024], level 2: u64 [batch][2048] */
int omr_ntt_forward_device(omr_ctx* ctx, int level, void* d_data, size_t batch, void* stream);
int omr_ntt_inverse_device(omr_ctx* ctx, int level, void* d_data, size_t batch, void* stream);

/* ---- wire / on-disk format ------------------------------------------------------------------------------------
 * The reference serialises nothing (only `Size` byte counts: key_gen/detection.rs:81-88, sender.rs:36); these versioned flat
 * little-endian blobs are what the Rust shim (ffi/omr-b200-sys), the CPU oracle, tfhe_omr_b200.blobs (Python) and this library
 * exchange.  File = 64-byte header ("OMRB200\0", then the fields of omr_blob_header in order) + the arrays of the kind, in the
 * order below, C-contiguous.  `domain` says how ring polynomials are stored (OMR_OUT_NTT_NATIVE / OMR_OUT_COEFF).
 *   1 detection_key      bsk1 u32[512][8][2][1024], ksk u32[1024][27][671], bsk2 u64[670][12][2][2048], trace u64[11][25][2][2048]
 *   2 clues              a u16[count][512], b u16[count][7]                    (CmLweCiphertext<u16> x count)
 *   3 pertinency_vector  pv u64[count][2][2048]                                (NttRlweCiphertext<F2> x count, index0 = first global index)
 *   4 digest             ct u64[count][2][2048]                                (aux = number of index ciphertexts, the rest are payload ciphertexts)
 *   5 payloads           payloads u16[count][612]
 *   6 secret_key         s0 i32[512], z1 i32[1024], s2 i32[670], z2 i32[2048]  (test vectors only)
 *   7 rlwe1              ct u32[count][2][1024]   sum of the 7 first-level accumulators (detector.rs:556), coefficient form
 *   8 lwe2               ct u32[count][671]       after key switch, modulus switch and offset (detector.rs:560-596)
 *   9 rlwe2              ct u64[count][2][2048]   after the second-level blind rotation (detector.rs:623), coefficient form
 *  10 clue_key           pa u16[512], pb u16[512]
 * Errors of these context-free calls are reported through omr_last_error(NULL). */
#define OMR_BLOB_VERSION 1u
#define OMR_BLOB_DETECTION_KEY 1u
#define OMR_BLOB_CLUES 2u
#define OMR_BLOB_PERTINENCY_VECTOR 3u
#define OMR_BLOB_DIGEST 4u
#define OMR_BLOB_PAYLOADS 5u
#define OMR_BLOB_SECRET_KEY 6u
#define OMR_BLOB_RLWE1 7u
#define OMR_BLOB_LWE2 8u
#define OMR_BLOB_RLWE2 9u
#define OMR_BLOB_CLUE_KEY 10u
typedef struct {
    uint32_t version, kind;
    uint64_t count, index0, aux, payload_bytes;
    uint32_t domain, reserved[3];
} omr_blob_header;
uint32_t omr_blob_field_count(uint32_t kind);                                    /* arrays of a kind (0 = unknown kind) */
size_t omr_blob_field_bytes(uint32_t kind, uint32_t field, uint64_t count);      /* bytes of array `field` at `count` */
int omr_blob_write(const char* path, uint32_t kind, uint64_t count, uint64_t index0, uint64_t aux, uint32_t domain,
                   const void* const* arrays, uint32_t n_arrays);
int omr_blob_read_header(const char* path, omr_blob_header* hdr);
/* arrays[i] must hold omr_blob_field_bytes(hdr.kind, i, hdr.count) bytes (read the header first) */
int omr_blob_read(const char* path, omr_blob_header* hdr /*nullable*/, void* const* arrays, uint32_t n_arrays);
/* Detector::new (detector.rs:85) from a detection-key blob; its `domain` selects OMR_KEYS_NTT_NATIVE / OMR_KEYS_COEFF */
int omr_ctx_create_from_blob(int device, const char* path, omr_ctx** out);

/* Step-0 peak: sustained rate of the register-only Shoup butterfly loop the NTTs are made of,
 * level 1 = 32-bit integer (q1), level 2 = 64-bit integer (q2), level 3 = q2 on the FP64 pipe (what the level-2 kernel
 * uses); the denominators of the compute roofline in bench.py. */
int omr_mulmod_peak(omr_ctx* ctx, int level, int iters, double* mulmods_per_second);

/* Launch shapes.  The reference detects one message per rayon task (examples/omr.rs:160-164), so a caller may hand over
 * anything from one clue set to a whole board.  Batches with fewer messages than SMs default to latency shapes (one level-1
 * blind rotation per CTA, key-switch rows split across CTAs, 512 threads per level-2 blind rotation); larger batches use
 * the throughput shapes.  Both give bit-identical results; enable = 0 forces the throughput shapes for every batch size. */
int omr_set_latency_shapes(omr_ctx* ctx, int enable);

/* Key switch path.  The LWE key switch (detector.rs:560-563) is a {-1,0,1} x u32 matrix product.  Default: the hand-written
 * CUDA-core kernels — keyswitch_dp4a_kernel (digit bits x byte-packed key limbs with IDP.4A; OMR_KS_DP4A=0 falls back to the plain
 * integer keyswitch_kernel) for large batches, keyswitch_kernel<split rows> for small ones.  Opt-in (enable = 1, or OMR_KS_GEMM=1 in the environment), when the library was built
 * with the CUTLASS headers: an exact int8 tensor-core GEMM (a CUTLASS template instance: int32 accumulation, key split into
 * 8-bit limbs; 74 MB of extra key material built on first use).  Identical results.  Enabling it on a library built without
 * it returns OMR_ERR_STATE.  omr_key_switch_path: 0 = CUDA cores, 1 = tensor-core GEMM. */
int omr_set_tensor_core_key_switch(omr_ctx* ctx, int enable);
int omr_key_switch_path(const omr_ctx* ctx);

/* number of kernels this library has launched on the context since creation (bench.py's gpu_launches) */
uint64_t omr_launch_count(const omr_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* OMR_B200_H */

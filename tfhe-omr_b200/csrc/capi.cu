// capi.cu — the C ABI of include/omr_b200.h over the kernels of kernels.cuh.  No CPU fallback anywhere: every
// compute entry point launches CUDA kernels or returns OMR_ERR_CUDA.
#include "../../include/omr_b200.h"
#include "kernels.cuh"
#include "keygen.cuh"
#include "nccl_dl.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>
#include <mutex>

using namespace omr;

namespace {

thread_local std::string g_create_error;

// ---- host-side constant tables (psi powers etc.) ------------------------------------------------------------------
template <class T> struct HostWide;
template <> struct HostWide<u32> { typedef u64 W; static constexpr int BITS = 32; };
template <> struct HostWide<u64> { typedef u128 W; static constexpr int BITS = 64; };
template <class T> T h_mulmod(T a, T b, T q) { return (T)((typename HostWide<T>::W)a * b % q); }
template <class T> T h_powmod(T b, u64 e, T q) { T r = 1; while (e) { if (e & 1) r = h_mulmod(r, b, q); b = h_mulmod(b, b, q); e >>= 1; } return r; }
template <class T> T h_shoup(T w, T q) { return (T)((((typename HostWide<T>::W)w) << HostWide<T>::BITS) / q); }
unsigned h_bitrev(unsigned x, int bits) { unsigned r = 0; for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; } return r; }

// psi = g^((q-1)/2N) for the smallest g whose power has order exactly 2N; tables in bit-reversed order
template <class T, class TW> void make_twiddles(int n, int logn, T q, std::vector<TW>& fwd, std::vector<TW>& inv) {
    T psi = 0;
    for (T g = 2;; ++g) { T c = h_powmod<T>(g, ((u64)q - 1) / (2 * (u64)n), q); if (h_powmod<T>(c, (u64)n, q) == q - 1) { psi = c; break; } }
    T ipsi = h_powmod<T>(psi, (u64)q - 2, q);
    std::vector<T> pw(n), ipw(n);
    pw[0] = ipw[0] = 1;
    for (int i = 1; i < n; ++i) { pw[i] = h_mulmod(pw[i - 1], psi, q); ipw[i] = h_mulmod(ipw[i - 1], ipsi, q); }
    fwd.resize(n); inv.resize(n);
    for (int i = 0; i < n; ++i) {
        unsigned r = h_bitrev(i, logn);
        fwd[i].x = pw[r]; fwd[i].y = h_shoup<T>(pw[r], q);
        inv[i].x = ipw[r]; inv[i].y = h_shoup<T>(ipw[r], q);
    }
}

#define CK(expr)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess) { ctx_fail(ctx, std::string(#expr) + ": " + cudaGetErrorString(e_)); return OMR_ERR_CUDA; } \
    } while (0)

}  // namespace

struct omr_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    Tables tb{};
    void *d_tw1 = nullptr, *d_itw1 = nullptr, *d_tw2 = nullptr, *d_itw2 = nullptr, *d_lut1 = nullptr, *d_lut2 = nullptr, *d_tw2d = nullptr, *d_itw2d = nullptr;
    u32 *bsk1 = nullptr, *ksk = nullptr; u64 *bsk2 = nullptr, *trk = nullptr;   // internal forms: bsk1/trk Montgomery * N^-1; bsk2 = centred doubles * N^-1
    uint2 n1_inv{}; ulonglong2 n2_inv{};
    size_t key_bytes = 0;
    // scratch for the batched pipeline, sized for `cap` messages
    size_t cap = 0; u32* s_rlwe1 = nullptr; u32* s_lwe2 = nullptr;
    size_t clue_cap = 0; unsigned short *s_ca = nullptr, *s_cb = nullptr;   // staged clues of a host-buffer call
    size_t cap7 = 0; u32* s_rlwe7 = nullptr;      // per-(message, clue) accumulators of the L1 kernel
    int l2c_max_clusters = 0;                     // co-resident 6-CTA clusters (cudaOccupancyMaxActiveClusters)
    // tensor-core key switch (large batches): key limbs [KSG_N][KSG_K] int8, per-chunk digits / products / CUTLASS workspace
    signed char* ksg_bt = nullptr; signed char* ksg_a = nullptr; i32* ksg_c = nullptr; void* ksg_ws = nullptr; size_t ksg_ws_bytes = 0;
    bool ks_gemm = false; size_t ksg_min_b = KSG_MIN_B;
    int* d_flag = nullptr;                        // "a weight draw was rejected" flag of omr_weights_from_seed_device
    double* l2c_scratch = nullptr;                // partial sums exchanged inside a level-2 cluster
    unsigned long long* ks_part = nullptr;        // [KS_SPLIT_MAXB][KSK_PAD] partial sums of the split key switch
    uint4* ksd = nullptr; long long* ksd_colsum = nullptr;   // IDP.4A key switch: byte-packed key limbs (77 MB) and the per-column constant
    // packing scratch
    u64* s_partial = nullptr; size_t partial_words = 0;
    u64* s_digest = nullptr; size_t digest_words = 0;
    unsigned short *s_payloads = nullptr, *s_weights = nullptr; size_t payload_elems = 0, weight_elems = 0;
    // resident pertinency store
    u64* pv = nullptr; size_t pv_cap = 0, pv_count = 0; u64 pv_index0 = 0; bool pv_any = false;
    cudaEvent_t ev[5] = {};
    uint64_t launches = 0;
    int n_sm = 148;
    bool latency_shapes = true;                   // omr_set_latency_shapes / OMR_LATENCY_SHAPES=0: throughput shapes for every batch size
    uint32_t out_domain = OMR_OUT_NTT_NATIVE;     // omr_set_output_domain: domain of host-buffer ciphertexts
    u64* s_coeff = nullptr; size_t coeff_words = 0;   // staging for coefficient-domain copies of pertinency ciphertexts
    std::vector<u32> h_lut1; std::vector<u64> h_lut2; // host copies of the test vectors (omr_first_level_lut / omr_second_level_lut)
    void* comm = nullptr;                             // own NCCL communicator (omr_comm_init), K7
    // streaming ingest (omr_stream_*): resident running digest + double-buffered pinned staging
    struct Stream {
        bool active = false; omr_retrieval_params rp{}; u64 index_seed = 0; u64 index0 = 0, count = 0; uint32_t n_idx = 0, n_pay = 0;
        u64* digest = nullptr; u64* part = nullptr; u64* pv = nullptr; unsigned short* weights = nullptr; size_t weight_elems = 0;
        unsigned char* pinned[2] = {nullptr, nullptr}; unsigned char* dev[2] = {nullptr, nullptr}; cudaEvent_t copied[2] = {nullptr, nullptr};
        int next = 0;
    } st;
};
constexpr size_t STREAM_CHUNK = 16384;                                       // messages per staged chunk (= the detect chunk: whole waves of both big kernels)
constexpr size_t STREAM_MSG_BYTES = (CLUE_N + CLUE_COUNT + PAYLOAD_LEN) * 2; // clue a, clue b, payload

namespace omr {            // ks_gemm.cu: the key switch as an int8 tensor-core GEMM (a CUTLASS template instance; opt-in), if it was compiled in
int ks_gemm_i8(const int8_t* A, const int8_t* B, int32_t* C, int M, int N, int K, void* workspace, size_t workspace_bytes, cudaStream_t s);
size_t ks_gemm_workspace(int M, int N, int K);
bool ks_gemm_available();
}

namespace omr { void set_global_error(const std::string& m) { g_create_error = m; } }   // blob.cu, nccl.cu: context-free errors

namespace {
void ctx_fail(omr_ctx* ctx, const std::string& m) { if (ctx) ctx->err = m; else g_create_error = m; }

// Every entry point runs on the context's device and puts the caller's current device back on return: a process that drives
// several GPUs (torch reads the current device with cudaGetDevice) must not find it changed behind its back.
struct DeviceGuard {
    int prev = -1; cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) err = cudaSetDevice(dev); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete; DeviceGuard& operator=(const DeviceGuard&) = delete;
};
// lock the context, switch to its device
#define ENTER(ctx) std::lock_guard<std::mutex> lk_((ctx)->mu); DeviceGuard dg_((ctx)->device); CK(dg_.err)

template <class T> int dalloc(omr_ctx* ctx, T** p, size_t n) {
    CK(cudaMalloc((void**)p, n * sizeof(T)));
    return OMR_OK;
}

int ensure_scratch(omr_ctx* ctx, size_t B) {
    if (B <= ctx->cap) return OMR_OK;
    cudaFree(ctx->s_rlwe1); cudaFree(ctx->s_lwe2);
    ctx->s_rlwe1 = nullptr; ctx->s_lwe2 = nullptr;      // a failed re-allocation must not leave stale pointers
    ctx->cap = 0;
    int st;
    if ((st = dalloc(ctx, &ctx->s_rlwe1, B * 2 * F1::N))) return st;
    if ((st = dalloc(ctx, &ctx->s_lwe2, B * LWE2_STRIDE_IN))) return st;
    ctx->cap = B;
    return OMR_OK;
}

int ensure_pv(omr_ctx* ctx, size_t need) {
    if (need <= ctx->pv_cap) return OMR_OK;
    size_t ncap = ctx->pv_cap ? ctx->pv_cap : 1024;
    while (ncap < need) ncap *= 2;
    u64* np = nullptr;
    CK(cudaMalloc((void**)&np, ncap * OMR_PV_WORDS * sizeof(u64)));
    if (ctx->pv && ctx->pv_count)
        CK(cudaMemcpyAsync(np, ctx->pv, ctx->pv_count * OMR_PV_WORDS * sizeof(u64), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->pv);
    ctx->pv = np; ctx->pv_cap = ncap;
    return OMR_OK;
}

// ---- launches -------------------------------------------------------------------------------------------------------
int launch_l1_raw(omr_ctx* ctx, const unsigned short* ca, const unsigned short* cb, size_t B, u32* rlwe7, u32* out, cudaStream_t s);
int launch_l1(omr_ctx* ctx, const unsigned short* ca, const unsigned short* cb, size_t B, u32* out, cudaStream_t s) {
    if (!B) return OMR_OK;
    if (B > ctx->cap7) {
        if (ctx->s_rlwe7) { CK(cudaStreamSynchronize(s)); cudaFree(ctx->s_rlwe7); ctx->s_rlwe7 = nullptr; ctx->cap7 = 0; }
        CK(cudaMalloc((void**)&ctx->s_rlwe7, B * CLUE_COUNT * 2 * F1::N * sizeof(u32)));
        ctx->cap7 = B;
    }
    return launch_l1_raw(ctx, ca, cb, B, ctx->s_rlwe7, out, s);
}
constexpr size_t KS_SPLIT_MAXB = 256;
int launch_ks(omr_ctx* ctx, const u32* rlwe, size_t B, u32* out, cudaStream_t s) {
    if (!B) return OMR_OK;
    dim3 grid((unsigned)((B + KS_MB - 1) / KS_MB), (KSK_PAD + KS_THREADS - 1) / KS_THREADS);
    const bool gemm = ctx->ks_gemm && B >= ctx->ksg_min_b;
    if (!gemm && B <= KS_SPLIT_MAXB && ctx->latency_shapes) {
        // small batch: too few CTAs to fill the GPU and each walks 27 648 key rows serially -> split the rows
        if (!ctx->ks_part) CK(cudaMalloc((void**)&ctx->ks_part, KS_SPLIT_MAXB * KSK_PAD * sizeof(unsigned long long)));
        unsigned z = 1;
        while (z < 64 && grid.x * grid.y * z < 2u * (unsigned)ctx->n_sm) z *= 2;
        grid.z = z;
        CK(cudaMemsetAsync(ctx->ks_part, 0, B * KSK_PAD * sizeof(unsigned long long), s));
        keyswitch_kernel<true><<<grid, KS_THREADS, KS_SMEM, s>>>(rlwe, ctx->ksk, out, (int)B, ctx->ks_part);
        keyswitch_finish_kernel<<<(unsigned)((B * KSK_PAD + 255) / 256), 256, 0, s>>>(rlwe, ctx->ks_part, out, (int)B);
        ctx->launches += 2; CK(cudaGetLastError());
        return OMR_OK;
    }
    if (gemm) {
        // large batch: digits x key limbs on the tensor cores, exact in int32; chunks bound the digit matrix (27 648 B per message)
        if (!ctx->ksg_a) {
            CK(cudaMalloc((void**)&ctx->ksg_a, KSG_CHUNK * (size_t)KSG_K));
            CK(cudaMalloc((void**)&ctx->ksg_c, KSG_CHUNK * (size_t)KSG_N * sizeof(i32)));
            ctx->ksg_ws_bytes = ks_gemm_workspace((int)KSG_CHUNK, KSG_N, KSG_K);
            if (ctx->ksg_ws_bytes) CK(cudaMalloc(&ctx->ksg_ws, ctx->ksg_ws_bytes));
        }
        for (size_t off = 0; off < B; off += KSG_CHUNK) {
            const size_t nb = B - off < KSG_CHUNK ? B - off : KSG_CHUNK;
            const u32* r = rlwe + off * 2 * F1::N;
            ks_digits_kernel<<<(unsigned)((nb * (KSG_K / 16) + 255) / 256), 256, 0, s>>>(r, ctx->ksg_a, (int)nb);
            const int rc = ks_gemm_i8((const int8_t*)ctx->ksg_a, (const int8_t*)ctx->ksg_bt, ctx->ksg_c, (int)nb, KSG_N, KSG_K, ctx->ksg_ws, ctx->ksg_ws_bytes, s);
            if (rc) { ctx_fail(ctx, "key switch GEMM failed (status " + std::to_string(rc) + ")"); return OMR_ERR_CUDA; }
            ks_combine_kernel<<<(unsigned)((nb * (LWE2_N + 1) + 255) / 256), 256, 0, s>>>(r, ctx->ksg_c, out + off * LWE2_STRIDE_IN, (int)nb);
            ctx->launches += 3; CK(cudaGetLastError());
        }
        return OMR_OK;
    }
    if (ctx->ksd) {     // throughput shape: IDP.4A over byte-packed key limbs (keyswitch_kernel<false> stays as the cross-check, OMR_KS_DP4A=0)
        dim3 dgrid((unsigned)((B + KSD_MB - 1) / KSD_MB), (KSK_PAD + KSD_THREADS - 1) / KSD_THREADS);
        keyswitch_dp4a_kernel<<<dgrid, KSD_THREADS, KSD_SMEM, s>>>(rlwe, ctx->ksd, ctx->ksd_colsum, out, (int)B);
        ++ctx->launches; CK(cudaGetLastError());
        return OMR_OK;
    }
    keyswitch_kernel<false><<<grid, KS_THREADS, KS_SMEM, s>>>(rlwe, ctx->ksk, out, (int)B, nullptr);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
// The 512-thread shape (one message per SM, a wave of n_sm messages in ~13.8 ms) has 97 % of the per-message throughput of the
// 256-thread shape (two messages per SM, a wave of 2 n_sm messages in ~26.8 ms) at half the wave granularity: whichever has
// the shorter sum of whole waves wins (the 512-thread one for every batch up to a few hundred messages, never for large ones).
bool l2_prefers_wide_ctas(size_t B, size_t n_sm) {
    if (B + 8 <= 2 * n_sm) return true;            // one or two waves of wide CTAs; a partial wave of the narrow shape measures 30-32 ms
    if (B > 8 * n_sm) return false;
    const size_t waves_wide = (B + n_sm - 1) / n_sm, waves_narrow = (B + 2 * n_sm - 1) / (2 * n_sm);
    return 138 * waves_wide < 268 * waves_narrow;
}
int launch_l2(omr_ctx* ctx, const u32* lwe, size_t B, u64* out, cudaStream_t s) {
    if (!B) return OMR_OK;
    // a cluster of 6 SMs per message: one wave of clusters takes ~1/3 of the time of the 512-thread shape, so up to two waves win
    if (B <= 2 * (size_t)ctx->l2c_max_clusters && ctx->latency_shapes) {
        const size_t cap = 2 * (size_t)ctx->l2c_max_clusters;
        if (!ctx->l2c_scratch) CK(cudaMalloc((void**)&ctx->l2c_scratch, cap * L2C_SCRATCH_WORDS * sizeof(double)));
        l2_blind_rotate_cluster_kernel<<<(unsigned)(B * L2C_CLUSTER), GeoL2::NT, L2C_SMEM, s>>>(lwe, reinterpret_cast<const double*>(ctx->bsk2), out,
                                                                                                 ctx->l2c_scratch, ctx->tb);
    } else if (ctx->latency_shapes && l2_prefers_wide_ctas(B, (size_t)ctx->n_sm))
        l2_blind_rotate_lat_kernel<<<(unsigned)B, L2L_THREADS, L2L_SMEM, s>>>(lwe, reinterpret_cast<const double*>(ctx->bsk2), out, ctx->tb);
    else
        l2_blind_rotate_kernel<<<(unsigned)B, L2_THREADS, L2_SMEM, s>>>(lwe, reinterpret_cast<const double*>(ctx->bsk2), out, ctx->tb);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
int launch_trace(omr_ctx* ctx, u64* ct, size_t B, cudaStream_t s) {
    if (!B) return OMR_OK;
    trace_kernel<<<(unsigned)B, TR_THREADS, TR_SMEM, s>>>(ct, reinterpret_cast<const double*>(ctx->trk), ctx->tb);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

// L1 kernel into caller-provided per-clue buffer (no internal allocation); the shape follows the number of blind rotations
int launch_l1_raw(omr_ctx* ctx, const unsigned short* ca, const unsigned short* cb, size_t B, u32* rlwe7, u32* out, cudaStream_t s) {
    const size_t n_clues = B * CLUE_COUNT;
    // Shape = rotations per CTA (1 = the 8-groups-per-rotation latency kernel).  One wave of n_sm CTAs takes about 2.6 / 6.5 /
    // 8.7 ms for 1 / 4 / 6 rotations per CTA; small and mid-size batches take the shape with the shortest sum of whole waves,
    // large ones (where the tail wave is negligible) the 6-rotation shape with the best per-rotation cost: 12 warps at 168
    // registers beat 16 warps at 128 (spills) by 7 % (profiles/r2_ab_l1_slots6_two_digits.txt).
    int shape = 6;
    if (ctx->latency_shapes && n_clues <= 16 * (size_t)ctx->n_sm) {
        const int slots[3] = {1, 4, 6}, wave_us[3] = {2600, 6530, 8670};
        size_t best = ~(size_t)0;
        for (int k = 0; k < 3; ++k) {
            const size_t ctas = (n_clues + slots[k] - 1) / slots[k], waves = (ctas + ctx->n_sm - 1) / ctx->n_sm, cost = waves * wave_us[k];
            if (cost < best) { best = cost; shape = slots[k]; }
        }
    }
    if (shape == 1)
        l1_blind_rotate_lat_kernel<<<(unsigned)n_clues, L1L_THREADS, L1L_SMEM, s>>>(ca, cb, ctx->bsk1, rlwe7, ctx->tb);
    else if (shape == 4)
        l1_blind_rotate_kernel<4><<<(unsigned)((n_clues + 3) / 4), L1Cfg<4>::THREADS, L1Cfg<4>::SMEM, s>>>(ca, cb, ctx->bsk1, rlwe7, (int)n_clues, ctx->tb);
    else
        l1_blind_rotate_kernel<6><<<(unsigned)((n_clues + 5) / 6), L1Cfg<6>::THREADS, L1Cfg<6>::SMEM, s>>>(ca, cb, ctx->bsk1, rlwe7, (int)n_clues, ctx->tb);
    ++ctx->launches; CK(cudaGetLastError());
    sum7_kernel<<<(unsigned)((B * 2 * F1::N + 255) / 256), 256, 0, s>>>(rlwe7, out, B);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

// The whole per-message pipeline for a batch, stream-ordered on the caller's stream; batches beyond MAXB messages are cut
// into chunks so that the scratch stays bounded.  (Running the integer level-1 side of one chunk concurrently with the FP64
// level-2 side of the previous one was tried and is slower — DESIGN.md §4 — so there is exactly one stream.)
int detect_device(omr_ctx* ctx, const unsigned short* d_ca, const unsigned short* d_cb, size_t B, u64* d_pv, cudaStream_t s,
                  omr_stage_times* times) {
    int st;
    if (times) *times = omr_stage_times{};
    if (!B) return OMR_OK;
    const size_t MAXB = 16384;          // (one launch per kernel for a whole 65 536 board was measured: no faster end to end, 4.3 GB of scratch)
    if ((st = ensure_scratch(ctx, B < MAXB ? B : MAXB))) return st;
    for (size_t off = 0; off < B; off += MAXB) {
        const size_t nb = B - off < MAXB ? B - off : MAXB;
        u64* pv = d_pv + off * OMR_PV_WORDS;
        if (times) CK(cudaEventRecord(ctx->ev[0], s));
        if ((st = launch_l1(ctx, d_ca + off * CLUE_N, d_cb + off * CLUE_COUNT, nb, ctx->s_rlwe1, s))) return st;
        if ((st = launch_ks(ctx, ctx->s_rlwe1, nb, ctx->s_lwe2, s))) return st;
        if (times) CK(cudaEventRecord(ctx->ev[1], s));
        if ((st = launch_l2(ctx, ctx->s_lwe2, nb, pv, s))) return st;
        if (times) CK(cudaEventRecord(ctx->ev[2], s));
        if ((st = launch_trace(ctx, pv, nb, s))) return st;
        if (times) {
            CK(cudaEventRecord(ctx->ev[3], s));
            CK(cudaEventSynchronize(ctx->ev[3]));
            float a = 0, b = 0, c = 0;
            CK(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
            CK(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
            CK(cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]));
            times->first_level_bootstrapping_ms += a; times->second_level_bootstrapping_ms += b; times->trace_ms += c;
            times->detect_ms += a + b + c;
        }
    }
    return OMR_OK;
}

int ensure_partial(omr_ctx* ctx, size_t words) {
    if (words <= ctx->partial_words) return OMR_OK;
    cudaFree(ctx->s_partial); ctx->s_partial = nullptr; ctx->partial_words = 0;
    int st; if ((st = dalloc(ctx, &ctx->s_partial, words))) return st;
    ctx->partial_words = words;
    return OMR_OK;
}

int pack_device(omr_ctx* ctx, bool indices, const u64* d_pv, size_t count, u64 index0, PackIndexArgs ia, PackPayloadArgs pa,
                unsigned n_cipher, u64* d_out, cudaStream_t s) {
    if (!n_cipher) return OMR_OK;
    if (count == 0) { CK(cudaMemsetAsync(d_out, 0, (size_t)n_cipher * OMR_PV_WORDS * sizeof(u64), s)); return OMR_OK; }
    const unsigned n_chunks = (unsigned)((count + PACK_CHUNK - 1) / PACK_CHUNK);
    if (n_chunks > 65535u) { ctx_fail(ctx, "pack: too many messages for one call (max 65535*128)"); return OMR_ERR_INVALID; }
    int st;
    if ((st = ensure_partial(ctx, (size_t)n_cipher * n_chunks * OMR_PV_WORDS))) return st;
    dim3 grid(n_cipher, n_chunks);
    if (indices) pack_kernel<true><<<grid, PACK_THREADS, PACK_SMEM, s>>>(d_pv, count, index0, ia, pa, ctx->s_partial, ctx->tb);
    else pack_kernel<false><<<grid, PACK_THREADS, PACK_SMEM, s>>>(d_pv, count, index0, ia, pa, ctx->s_partial, ctx->tb);
    ++ctx->launches; CK(cudaGetLastError());
    dim3 rgrid((OMR_PV_WORDS + 255) / 256, n_cipher);
    reduce_partials_kernel<<<rgrid, 256, 0, s>>>(ctx->s_partial, (int)n_chunks, d_out);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

int ensure_digest(omr_ctx* ctx, size_t words) {
    if (words <= ctx->digest_words) return OMR_OK;
    cudaFree(ctx->s_digest); ctx->s_digest = nullptr; ctx->digest_words = 0;
    int st; if ((st = dalloc(ctx, &ctx->s_digest, words))) return st;
    ctx->digest_words = words;
    return OMR_OK;
}

void stream_release(omr_ctx* ctx) {
    auto& st = ctx->st;
    cudaFree(st.digest); cudaFree(st.part); cudaFree(st.pv); cudaFree(st.weights);
    for (int i = 0; i < 2; ++i) { if (st.pinned[i]) cudaFreeHost(st.pinned[i]); cudaFree(st.dev[i]); if (st.copied[i]) cudaEventDestroy(st.copied[i]); }
    st = omr_ctx::Stream{};
}

// detection-key generation request (omr_generate_detector): the key is made on the device and loaded from there
struct KeygenSpec { const omr_secret_key* sk; const uint8_t* seed32; const omr_key_blobs* host_out; };
struct DevBufs {           // temporaries freed on every return path
    std::vector<void*> p;
    template <class T> cudaError_t alloc(T** q, size_t n) { cudaError_t e = cudaMalloc((void**)q, n * sizeof(T)); if (e == cudaSuccess) p.push_back(*q); return e; }
    ~DevBufs() { for (void* q : p) cudaFree(q); }
};

int create_impl(int device, const omr_key_blobs* keys, bool keys_on_device, omr_ctx** out, const KeygenSpec* gen = nullptr) {
    if (!out || (!gen && (!keys || !keys->bsk1 || !keys->ksk || !keys->bsk2 || !keys->trace))) { g_create_error = "null argument"; return OMR_ERR_INVALID; }
    if (!gen && keys->flags != OMR_KEYS_NTT_NATIVE && keys->flags != OMR_KEYS_COEFF) { g_create_error = "unknown key flags"; return OMR_ERR_INVALID; }
    if (gen && (!gen->sk || !gen->seed32 || !gen->sk->s0 || !gen->sk->z1 || !gen->sk->s2 || !gen->sk->z2)) { g_create_error = "null argument"; return OMR_ERR_INVALID; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e);
        return OMR_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { g_create_error = "bad device index"; return OMR_ERR_INVALID; }
    omr_ctx* ctx = new omr_ctx;
    ctx->device = device;
    auto fail = [&](int st) { g_create_error = ctx->err; omr_ctx_destroy(ctx); return st; };
#define CKC(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e_); return fail(OMR_ERR_CUDA); } } while (0)
    DeviceGuard dg(device); CKC(dg.err);
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (auto& ev : ctx->ev) CKC(cudaEventCreate(&ev));
    // constants
    if ((u32)(Q1 * F1::QINV_NEG) != 0xFFFFFFFFu || (u64)(Q2 * F2::QINV_NEG) != ~0ull) { ctx->err = "bad Montgomery constants"; return fail(OMR_ERR_INVALID); }
    std::vector<uint2> tw1, itw1; std::vector<ulonglong2> tw2, itw2;
    make_twiddles<u32, uint2>(F1::N, F1::LOGN, Q1, tw1, itw1);
    make_twiddles<u64, ulonglong2>(F2::N, F2::LOGN, Q2, tw2, itw2);
    // LUTs: detector.rs:457-503 with lut.rs:12-27 (chunks of N>>log_t coefficients take v0,v1,v1,v2,v2,...)
    std::vector<u32> lut1(F1::N, 0); std::vector<u64> lut2(F2::N, 0);
    {
        const u32 s1 = ((Q1 >> 4) + 1) >> 1, vals[5] = {s1, 0, 0, 0, Q1 - s1};
        const int hd = F1::N >> 3;
        for (int c = 0; c < F1::N / hd; ++c) { int vi = (c + 1) / 2; if (vi < 5) for (int j = 0; j < hd; ++j) lut1[c * hd + j] = vals[vi]; }
        const u64 s2 = (2 * Q2 + OUT_P) / (2ull * OUT_P);        // round_half_up(q2 / 257), detector.rs:489-495
        const int hd2 = F2::N >> 5;
        for (int c = 0; c < F2::N / hd2; ++c) { int vi = (c + 1) / 2; if (vi == 2 * CLUE_COUNT) for (int j = 0; j < hd2; ++j) lut2[c * hd2 + j] = s2; }
    }
    auto upload = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
        cudaError_t r = cudaMalloc(dst, bytes); if (r != cudaSuccess) return r;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    CKC(upload(&ctx->d_tw1, tw1.data(), tw1.size() * sizeof(uint2)));
    CKC(upload(&ctx->d_itw1, itw1.data(), itw1.size() * sizeof(uint2)));
    CKC(upload(&ctx->d_tw2, tw2.data(), tw2.size() * sizeof(ulonglong2)));
    CKC(upload(&ctx->d_itw2, itw2.data(), itw2.size() * sizeof(ulonglong2)));
    {   // FP64-path twiddles: centred w and w/q2 (division in long double, rounded once to double)
        std::vector<double2> t2d(F2::N), it2d(F2::N);
        auto cen = [](u64 v) { return v > (Q2 >> 1) ? -(double)(int64_t)(Q2 - v) : (double)(int64_t)v; };
        for (int i = 0; i < F2::N; ++i) {
            t2d[i].x = cen(tw2[i].x); t2d[i].y = (double)((long double)t2d[i].x / (long double)Q2);
            it2d[i].x = cen(itw2[i].x); it2d[i].y = (double)((long double)it2d[i].x / (long double)Q2);
        }
        CKC(upload(&ctx->d_tw2d, t2d.data(), t2d.size() * sizeof(double2)));
        CKC(upload(&ctx->d_itw2d, it2d.data(), it2d.size() * sizeof(double2)));
    }
    ctx->h_lut1 = lut1; ctx->h_lut2 = lut2;
    CKC(upload(&ctx->d_lut1, lut1.data(), lut1.size() * sizeof(u32)));
    CKC(upload(&ctx->d_lut2, lut2.data(), lut2.size() * sizeof(u64)));
    Tables& tb = ctx->tb;
    tb.tw1 = (const uint2*)ctx->d_tw1; tb.itw1 = (const uint2*)ctx->d_itw1;
    tb.tw2 = (const ulonglong2*)ctx->d_tw2; tb.itw2 = (const ulonglong2*)ctx->d_itw2;
    tb.tw2d = (const double2*)ctx->d_tw2d; tb.itw2d = (const double2*)ctx->d_itw2d;
    tb.lut1 = (const u32*)ctx->d_lut1; tb.lut2 = (const u64*)ctx->d_lut2;
    const u32 n1i = h_powmod<u32>(F1::N, Q1 - 2, Q1); const u64 n2i = h_powmod<u64>(F2::N, Q2 - 2, Q2);
    ctx->n1_inv = make_uint2(n1i, h_shoup<u32>(n1i, Q1)); ctx->n2_inv = make_ulonglong2(n2i, h_shoup<u64>(n2i, Q2));
    tb.n2_inv = ctx->n2_inv;
    const u64 r2 = (u64)(((u128)1 << 64) % Q2); tb.r2 = make_ulonglong2(r2, h_shoup<u64>(r2, Q2));
    for (int t = 0; t < TR_STEPS; ++t) {
        const u32 d = (1u << (TR_STEPS - t)) + 1, M = 2 * F2::N; u32 inv = 1;
        for (u32 x = 1; x < M; x += 2) if ((x * d) % M == 1) { inv = x; break; }
        tb.trace_dinv[t] = inv;
    }
    {   // first-pass twiddles (uniform across threads) into constant memory of this device
        uint2 h1[16] = {}; double2 h2[8] = {};
        for (int i = 1; i < 16; ++i) h1[i] = tw1[i];
        for (int i = 1; i < 8; ++i) {
            const u64 v = tw2[i].x; const double c = v > (Q2 >> 1) ? -(double)(int64_t)(Q2 - v) : (double)(int64_t)v;
            h2[i] = make_double2(c, (double)((long double)c / (long double)Q2));
        }
        CKC(cudaMemcpyToSymbol(c_tw1_head, h1, sizeof h1));
        CKC(cudaMemcpyToSymbol(c_tw2d_head, h2, sizeof h2));
    }
    CKC(cudaFuncSetAttribute(l1_blind_rotate_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L1Cfg<6>::SMEM));
    CKC(cudaFuncSetAttribute(l1_blind_rotate_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L1L_SMEM));
    CKC(cudaFuncSetAttribute(l1_blind_rotate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L1Cfg<4>::SMEM));
    { cudaDeviceProp prop; CKC(cudaGetDeviceProperties(&prop, device)); ctx->n_sm = prop.multiProcessorCount; }
    // always carve out the maximum shared memory for the big kernels: with the driver's default heuristic an occasional
    // launch of l2_blind_rotate_kernel got a smaller carve-out and ran at 1 CTA/SM (278 ms instead of 215 ms for 2 368 messages)
    CKC(cudaFuncSetAttribute(l1_blind_rotate_kernel<6>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(l2_blind_rotate_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(keyswitch_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(pack_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CKC(cudaFuncSetAttribute(pack_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if (const char* e = getenv("OMR_LATENCY_SHAPES")) ctx->latency_shapes = atoi(e) != 0;
    CKC(cudaFuncSetAttribute(keyswitch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM));
    CKC(cudaFuncSetAttribute(keyswitch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS_SMEM));
    CKC(cudaFuncSetAttribute(l2_blind_rotate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L2_SMEM));
    CKC(cudaFuncSetAttribute(l2_blind_rotate_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L2L_SMEM));
    CKC(cudaFuncSetAttribute(l2_blind_rotate_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L2C_SMEM));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(L2C_CLUSTER * ctx->n_sm)); cfg.blockDim = dim3(GeoL2::NT); cfg.dynamicSmemBytes = L2C_SMEM;
        cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = L2C_CLUSTER; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int nc = 0;
        if (cudaOccupancyMaxActiveClusters(&nc, l2_blind_rotate_cluster_kernel, &cfg) != cudaSuccess) { nc = 0; cudaGetLastError(); }
        // one CTA per SM is the point of the shape: never count more clusters than SMs / 6
        ctx->l2c_max_clusters = nc < ctx->n_sm / L2C_CLUSTER ? nc : ctx->n_sm / L2C_CLUSTER;
        if (const char* e = getenv("OMR_L2_CLUSTERS")) ctx->l2c_max_clusters = atoi(e);
    }
    CKC(cudaFuncSetAttribute(trace_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TR_SMEM));
    // keys -> internal form: every ring word * (R * N^-1) mod q; KSK padded to a 672-word row stride
    const size_t n_bsk1 = (size_t)CLUE_N * 2 * G1::LEVELS * 2 * F1::N, n_ksk_rows = (size_t)F1::N * KS_LEVELS,
                 n_bsk2 = (size_t)LWE2_N * 2 * G2::LEVELS * 2 * F2::N, n_trk = (size_t)TR_STEPS * TR_LEVELS * 2 * F2::N;
    CKC(cudaMalloc((void**)&ctx->bsk1, n_bsk1 * 4)); CKC(cudaMalloc((void**)&ctx->ksk, n_ksk_rows * KSK_PAD * 4));
    CKC(cudaMalloc((void**)&ctx->bsk2, n_bsk2 * 8)); CKC(cudaMalloc((void**)&ctx->trk, n_trk * 8));
    ctx->key_bytes = n_bsk1 * 4 + n_ksk_rows * KSK_PAD * 4 + n_bsk2 * 8 + n_trk * 8;
    cudaStream_t s = ctx->stream;
    DevBufs tmp_bufs; omr_key_blobs gen_blobs{};
    if (gen) {
        // SecretKeyPack::generate_detection_key (secret.rs:118-178) on the device, into flat NTT-native blobs
        for (int i = 0; i < CLUE_N; ++i) if (gen->sk->s0[i] != 0 && gen->sk->s0[i] != 1) { ctx->err = "keygen: s0 must be binary"; return fail(OMR_ERR_INVALID); }
        for (int i = 0; i < LWE2_N; ++i) if (gen->sk->s2[i] != 0 && gen->sk->s2[i] != 1) { ctx->err = "keygen: s2 must be binary"; return fail(OMR_ERR_INVALID); }
        for (int i = 0; i < F1::N; ++i) if (gen->sk->z1[i] < -1 || gen->sk->z1[i] > 1) { ctx->err = "keygen: z1 must be ternary"; return fail(OMR_ERR_INVALID); }
        for (int i = 0; i < F2::N; ++i) if (gen->sk->z2[i] < -1 || gen->sk->z2[i] > 1) { ctx->err = "keygen: z2 must be ternary"; return fail(OMR_ERR_INVALID); }
        u32 *g_bsk1 = nullptr, *g_ksk = nullptr, *d_z1n = nullptr; u64 *g_bsk2 = nullptr, *g_trk = nullptr, *d_z2c = nullptr, *d_z2n = nullptr, *d_zs = nullptr;
        i32 *d_s0 = nullptr, *d_z1 = nullptr, *d_s2 = nullptr, *d_z2 = nullptr;
        CKC(tmp_bufs.alloc(&g_bsk1, n_bsk1)); CKC(tmp_bufs.alloc(&g_ksk, n_ksk_rows * LWE2_STRIDE_IN));
        CKC(tmp_bufs.alloc(&g_bsk2, n_bsk2)); CKC(tmp_bufs.alloc(&g_trk, n_trk));
        CKC(tmp_bufs.alloc(&d_s0, (size_t)CLUE_N)); CKC(tmp_bufs.alloc(&d_z1, (size_t)F1::N)); CKC(tmp_bufs.alloc(&d_s2, (size_t)LWE2_N)); CKC(tmp_bufs.alloc(&d_z2, (size_t)F2::N));
        CKC(tmp_bufs.alloc(&d_z1n, (size_t)F1::N)); CKC(tmp_bufs.alloc(&d_z2c, (size_t)F2::N)); CKC(tmp_bufs.alloc(&d_z2n, (size_t)F2::N)); CKC(tmp_bufs.alloc(&d_zs, (size_t)TR_STEPS * F2::N));
        CKC(cudaMemcpyAsync(d_s0, gen->sk->s0, CLUE_N * 4, cudaMemcpyHostToDevice, s)); CKC(cudaMemcpyAsync(d_z1, gen->sk->z1, F1::N * 4, cudaMemcpyHostToDevice, s));
        CKC(cudaMemcpyAsync(d_s2, gen->sk->s2, LWE2_N * 4, cudaMemcpyHostToDevice, s)); CKC(cudaMemcpyAsync(d_z2, gen->sk->z2, F2::N * 4, cudaMemcpyHostToDevice, s));
        const ChaChaKey ck = chacha_key_from_seed(gen->seed32);
        kg_lift_kernel<F1><<<(F1::N + 255) / 256, 256, 0, s>>>(d_z1, d_z1n, F1::N);
        kg_lift_kernel<F2><<<(F2::N + 255) / 256, 256, 0, s>>>(d_z2, d_z2c, F2::N);
        CKC(cudaMemcpyAsync(d_z2n, d_z2c, F2::N * 8, cudaMemcpyDeviceToDevice, s));
        kg_automorph_kernel<<<dim3((F2::N + 255) / 256, TR_STEPS), 256, 0, s>>>(d_z2c, d_zs);
        ntt_kernel<F1, false><<<1, ntt_kernel_threads<F1>(), ntt_kernel_smem<F1>(), s>>>(d_z1n, tb, ctx->n1_inv);
        ntt_kernel<F2, false><<<1, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(d_z2n, tb, ctx->n2_inv);
        ntt_kernel<F2, false><<<TR_STEPS, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(d_zs, tb, ctx->n2_inv);
        kg_rlwe_rows_kernel<F1, G1, 0, 21><<<(unsigned)(CLUE_N * 2 * G1::LEVELS), ntt_kernel_threads<F1>(), ntt_kernel_smem<F1>(), s>>>(ck, KG_BSK1_A, KG_BSK1_E, d_z1n, d_s0, nullptr, g_bsk1, tb);
        kg_rlwe_rows_kernel<F2, G2, 0, 3><<<(unsigned)(LWE2_N * 2 * G2::LEVELS), ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(ck, KG_BSK2_A, KG_BSK2_E, d_z2n, d_s2, nullptr, g_bsk2, tb);
        kg_rlwe_rows_kernel<F2, GT, 1, 3><<<(unsigned)(TR_STEPS * TR_LEVELS), ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(ck, KG_TRK_A, KG_TRK_E, d_z2n, nullptr, d_zs, g_trk, tb);
        kg_ksk_rows_kernel<<<(unsigned)n_ksk_rows, KG_KSK_THREADS, 0, s>>>(ck, d_s2, d_z1, g_ksk);
        ctx->launches += 11; CKC(cudaGetLastError());
        if (gen->host_out) {
            CKC(cudaMemcpyAsync((void*)gen->host_out->bsk1, g_bsk1, n_bsk1 * 4, cudaMemcpyDeviceToHost, s));
            CKC(cudaMemcpyAsync((void*)gen->host_out->ksk, g_ksk, n_ksk_rows * LWE2_STRIDE_IN * 4, cudaMemcpyDeviceToHost, s));
            CKC(cudaMemcpyAsync((void*)gen->host_out->bsk2, g_bsk2, n_bsk2 * 8, cudaMemcpyDeviceToHost, s));
            CKC(cudaMemcpyAsync((void*)gen->host_out->trace, g_trk, n_trk * 8, cudaMemcpyDeviceToHost, s));
        }
        CKC(cudaStreamSynchronize(s));
        gen_blobs = omr_key_blobs{g_bsk1, g_ksk, g_bsk2, g_trk, OMR_KEYS_NTT_NATIVE};
        keys = &gen_blobs; keys_on_device = true;
    }
    const cudaMemcpyKind kind = keys_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    // device-resident keys were produced on some stream of the caller: the context's own (non-blocking) stream has no ordering
    // against it, so drain the device once before the first copy
    if (keys_on_device) CKC(cudaDeviceSynchronize());
    {
        u32* tmp = nullptr; CKC(cudaMalloc((void**)&tmp, n_ksk_rows * LWE2_STRIDE_IN * 4));
        CKC(cudaMemcpyAsync(tmp, keys->ksk, n_ksk_rows * LWE2_STRIDE_IN * 4, kind, s));
        ksk_pad_kernel<<<(unsigned)((n_ksk_rows * KSK_PAD + 255) / 256), 256, 0, s>>>(tmp, ctx->ksk, n_ksk_rows); ++ctx->launches;
        CKC(cudaStreamSynchronize(s)); cudaFree(tmp);
    }
    {   // byte-packed limbs + column constant for the IDP.4A key switch
        const char* e = getenv("OMR_KS_DP4A");
        if (!e || atoi(e) != 0) {
            CKC(cudaMalloc((void**)&ctx->ksd, KSD_WORDS * 4)); CKC(cudaMalloc((void**)&ctx->ksd_colsum, KSK_PAD * sizeof(long long)));
            ksd_build_kernel<<<(unsigned)((KSD_WORDS / 4 + 255) / 256), 256, 0, s>>>(ctx->ksk, ctx->ksd);
            ksd_colsum_kernel<<<(KSK_PAD + 127) / 128, 128, 0, s>>>(ctx->ksk, ctx->ksd_colsum);
            ctx->launches += 2; CKC(cudaGetLastError());
            CKC(cudaStreamSynchronize(s));
            ctx->key_bytes += KSD_WORDS * 4;
        }
    }
    // key switch: the hand-written CUDA-core kernels by default; OMR_KS_GEMM=1 or omr_set_tensor_core_key_switch(ctx, 1) opts into
    // the CUTLASS int8 GEMM (limbs are built then)
    if (const char* e = getenv("OMR_KS_GEMM_MIN")) { long v = atol(e); if (v >= 1) ctx->ksg_min_b = (size_t)v; }
    const bool coeff = keys->flags == OMR_KEYS_COEFF;
    {
        const u32 c1 = h_mulmod<u32>((u32)(((u64)1 << 32) % Q1), n1i, Q1);
        CKC(cudaMemcpyAsync(ctx->bsk1, keys->bsk1, n_bsk1 * 4, kind, s));
        if (coeff) { ntt_kernel<F1, false><<<(unsigned)(n_bsk1 / F1::N), ntt_kernel_threads<F1>(), ntt_kernel_smem<F1>(), s>>>(ctx->bsk1, tb, ctx->n1_inv); ++ctx->launches; }
        scale_kernel<F1><<<(unsigned)((n_bsk1 + 255) / 256), 256, 0, s>>>(ctx->bsk1, ctx->bsk1, n_bsk1, make_uint2(c1, h_shoup<u32>(c1, Q1))); ++ctx->launches;
        CKC(cudaMemcpyAsync(ctx->bsk2, keys->bsk2, n_bsk2 * 8, kind, s));
        if (coeff) { ntt_kernel<F2, false><<<(unsigned)(n_bsk2 / F2::N), ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(ctx->bsk2, tb, ctx->n2_inv); ++ctx->launches; }
        key_to_double_kernel<<<(unsigned)((n_bsk2 + 255) / 256), 256, 0, s>>>(ctx->bsk2, reinterpret_cast<double*>(ctx->bsk2), n_bsk2, ctx->n2_inv); ++ctx->launches;
        CKC(cudaMemcpyAsync(ctx->trk, keys->trace, n_trk * 8, kind, s));
        if (coeff) { ntt_kernel<F2, false><<<(unsigned)(n_trk / F2::N), ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(ctx->trk, tb, ctx->n2_inv); ++ctx->launches; }
        key_to_double_kernel<<<(unsigned)((n_trk + 255) / 256), 256, 0, s>>>(ctx->trk, reinterpret_cast<double*>(ctx->trk), n_trk, ctx->n2_inv); ++ctx->launches;
        CKC(cudaGetLastError());
        CKC(cudaStreamSynchronize(s));
    }
#undef CKC
    if (const char* e = getenv("OMR_KS_GEMM")) {
        if (atoi(e) != 0 && ks_gemm_available()) { const int st = omr_set_tensor_core_key_switch(ctx, 1); if (st) return fail(st); }
    }
    *out = ctx;
    return OMR_OK;
}

}  // namespace

extern "C" {

int omr_generate_detector(int device, const omr_secret_key* sk, const uint8_t* seed32, const omr_key_blobs* host_keys_out, omr_ctx** out) {
    if (host_keys_out && (!host_keys_out->bsk1 || !host_keys_out->ksk || !host_keys_out->bsk2 || !host_keys_out->trace)) { g_create_error = "generate_detector: null key buffer"; return OMR_ERR_INVALID; }
    const KeygenSpec gen{sk, seed32, host_keys_out};
    return create_impl(device, nullptr, false, out, &gen);
}
int omr_ctx_create(int device, const omr_key_blobs* keys, omr_ctx** out) { return create_impl(device, keys, false, out); }
int omr_ctx_create_device_keys(int device, const omr_key_blobs* keys, omr_ctx** out) { return create_impl(device, keys, true, out); }

void omr_ctx_destroy(omr_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard dg(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void* ptrs[] = {ctx->d_tw1, ctx->d_itw1, ctx->d_tw2, ctx->d_itw2, ctx->d_lut1, ctx->d_lut2, ctx->d_tw2d, ctx->d_itw2d, ctx->bsk1, ctx->ksk, ctx->bsk2, ctx->trk,
                    ctx->ks_part, ctx->ksd, ctx->ksd_colsum, ctx->l2c_scratch, ctx->d_flag, ctx->ksg_bt, ctx->ksg_a, ctx->ksg_c, ctx->ksg_ws, ctx->s_rlwe1, ctx->s_rlwe7, ctx->s_lwe2, ctx->s_ca, ctx->s_cb, ctx->s_partial, ctx->s_digest, ctx->s_payloads, ctx->s_weights, ctx->pv, ctx->s_coeff};
    for (void* p : ptrs) if (p) cudaFree(p);
    stream_release(ctx);
    if (ctx->comm) { if (const NcclApi* api = nccl_api(nullptr)) api->comm_destroy(ctx->comm); ctx->comm = nullptr; }
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* omr_last_error(const omr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
size_t omr_detect_key_size(const omr_ctx* ctx) { return ctx ? ctx->key_bytes : 0; }
uint64_t omr_launch_count(const omr_ctx* ctx) { return ctx ? ctx->launches : 0; }
int omr_first_level_lut(const omr_ctx* ctx, uint32_t* out) {
    if (!ctx || !out || ctx->h_lut1.size() != (size_t)OMR_N1) return OMR_ERR_INVALID;
    memcpy(out, ctx->h_lut1.data(), OMR_N1 * sizeof(uint32_t));
    return OMR_OK;
}
int omr_second_level_lut(const omr_ctx* ctx, uint64_t* out) {
    if (!ctx || !out || ctx->h_lut2.size() != (size_t)OMR_N2) return OMR_ERR_INVALID;
    memcpy(out, ctx->h_lut2.data(), OMR_N2 * sizeof(uint64_t));
    return OMR_OK;
}
int omr_set_output_domain(omr_ctx* ctx, uint32_t domain) {
    if (!ctx || (domain != OMR_OUT_NTT_NATIVE && domain != OMR_OUT_COEFF)) { ctx_fail(ctx, "unknown output domain"); return OMR_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->out_domain = domain;
    return OMR_OK;
}
int omr_set_tensor_core_key_switch(omr_ctx* ctx, int enable) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    if (enable && !ks_gemm_available()) { ctx_fail(ctx, "the tensor-core key switch was not built into this library"); return OMR_ERR_STATE; }
    if (enable && !ctx->ksg_bt) {   // key limbs for the tensor-core key switch: [KSG_N][KSG_K] int8, 74 MB, built on first use
        CK(cudaMalloc((void**)&ctx->ksg_bt, (size_t)KSG_N * KSG_K));
        CK(cudaMemsetAsync(ctx->ksg_bt, 0, (size_t)KSG_N * KSG_K, ctx->stream));
        ks_limbs_kernel<<<(unsigned)(((size_t)KSG_K * (KSG_N / KSG_LIMBS) + 255) / 256), 256, 0, ctx->stream>>>(ctx->ksk, ctx->ksg_bt); ++ctx->launches;
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->ks_gemm = enable != 0;
    return OMR_OK;
}
int omr_key_switch_path(const omr_ctx* ctx) { return ctx && ctx->ks_gemm ? 1 : 0; }
int omr_set_latency_shapes(omr_ctx* ctx, int enable) {
    if (!ctx) return OMR_ERR_INVALID;
    ctx->latency_shapes = enable != 0;
    return OMR_OK;
}

int omr_retrieval_params_init(uint64_t all_payloads_count, uint32_t pertinent_count, omr_retrieval_params* rp) {
    if (!rp) return OMR_ERR_INVALID;
    // RetrievalParams::new — parameters/retrieval_params.rs:50-106, with the constants of secret.rs:196-203
    rp->index_modulus = OMR_P; rp->polynomial_size = OMR_N2; rp->bucket_count_per_segment = 130; rp->segment_count = 25;
    rp->cmb_count_per_cipher = 2; rp->all_payloads_count = all_payloads_count; rp->pertinent_count = pertinent_count;
    uint32_t e = 1; uint64_t pw = OMR_P;
    while (pw < all_payloads_count) { pw *= OMR_P; ++e; }
    rp->slots_per_bucket = e + 1;
    rp->slots_per_segment = rp->slots_per_bucket * rp->bucket_count_per_segment;
    rp->segment_per_cipher = rp->polynomial_size / rp->slots_per_segment;
    rp->max_encode_indices_cipher_count = rp->segment_count / rp->segment_per_cipher;
    rp->combination_count = pertinent_count + 5;
    return OMR_OK;
}

}  // extern "C"

// ---- unlocked implementations (the caller holds ctx->mu and has switched to ctx->device) ---------------------------------
namespace {

int check_rp(omr_ctx* ctx, const omr_retrieval_params* rp) {
    // encode_pertinent_indices asserts polynomial_size == ntt dimension (detector.rs:236)
    if (!rp || rp->polynomial_size != (uint32_t)OMR_N2 || rp->index_modulus != OMR_P || rp->slots_per_bucket < 2 ||
        rp->slots_per_segment != rp->slots_per_bucket * rp->bucket_count_per_segment ||
        rp->segment_per_cipher * rp->slots_per_segment > (uint32_t)OMR_N2 || rp->segment_per_cipher == 0) {
        ctx_fail(ctx, "invalid retrieval params"); return OMR_ERR_INVALID;
    }
    return OMR_OK;
}

int encode_indices_impl(omr_ctx* ctx, const omr_retrieval_params* rp, const u64* d_pv, size_t count, u64 index0, u64 seed, uint32_t cipher_idx0,
                        uint32_t n_cipher, u64* d_out, cudaStream_t s) {
    int st; if ((st = check_rp(ctx, rp))) return st;
    PackIndexArgs ia{rp->slots_per_bucket, rp->slots_per_segment, rp->segment_per_cipher, rp->bucket_count_per_segment, seed, cipher_idx0};
    return pack_device(ctx, true, d_pv, count, index0, ia, PackPayloadArgs{}, n_cipher, d_out, s);
}
int encode_payloads_impl(omr_ctx* ctx, const u64* d_pv, const unsigned short* d_payloads, size_t count, u64 index0, const unsigned short* d_weights,
                         size_t weight_stride, uint32_t n_cipher, uint32_t cmb_per_cipher, u64* d_out, cudaStream_t s) {
    if (cmb_per_cipher == 0 || cmb_per_cipher * OMR_PAYLOAD_LEN > OMR_N2 || index0 + count > weight_stride) {
        ctx_fail(ctx, "encode_payloads: bad combination layout"); return OMR_ERR_INVALID;
    }
    PackPayloadArgs pa{d_payloads, d_weights, weight_stride, cmb_per_cipher};
    return pack_device(ctx, false, d_pv, count, index0, PackIndexArgs{}, pa, n_cipher, d_out, s);
}
int weights_from_seed_impl(omr_ctx* ctx, const uint8_t* seed32, size_t count, unsigned short* d_out, uint32_t flags, cudaStream_t s) {
    if (!count) return OMR_OK;
    if (!ctx->d_flag) CK(cudaMalloc((void**)&ctx->d_flag, sizeof(int)));
    const ChaChaKey key = chacha_key_from_seed(seed32);
    CK(cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), s));
    const size_t blocks = (count + 15) / 16;
    weights_kernel<<<(unsigned)((blocks + 127) / 128), 128, 0, s>>>(key, count, d_out, ctx->d_flag);
    weights_serial_kernel<<<1, 1, 0, s>>>(key, count, d_out, ctx->d_flag, (int)(flags & 1u));
    ctx->launches += 2; CK(cudaGetLastError());
    return OMR_OK;
}
int decrypt_decode_impl(omr_ctx* ctx, const u64* d_z2_ntt, const u64* d_ct, size_t n, unsigned short* d_out, cudaStream_t s) {
    if (!n) return OMR_OK;
    int st; if ((st = ensure_partial(ctx, n * F2::N))) return st;
    const size_t total = n * F2::N;
    decrypt_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(d_ct, d_z2_ntt, ctx->s_partial, n);
    ++ctx->launches; CK(cudaGetLastError());
    ntt_kernel<F2, true><<<(unsigned)n, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(ctx->s_partial, ctx->tb, ctx->n2_inv);
    ++ctx->launches; CK(cudaGetLastError());
    decode_round_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ctx->s_partial, d_out, total);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
// n_polys second-level polynomials in place: NTT-native -> coefficient form (scaled by N^-1) or back
int to_coeff_impl(omr_ctx* ctx, u64* d_polys, size_t n_polys, cudaStream_t s) {
    if (!n_polys) return OMR_OK;
    ntt_kernel<F2, true><<<(unsigned)n_polys, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(d_polys, ctx->tb, ctx->n2_inv);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
int from_coeff_impl(omr_ctx* ctx, u64* d_polys, size_t n_polys, cudaStream_t s) {
    if (!n_polys) return OMR_OK;
    ntt_kernel<F2, false><<<(unsigned)n_polys, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>(d_polys, ctx->tb, ctx->n2_inv);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
// device -> host copy of n ciphertexts in the context's output domain (OMR_OUT_COEFF: through a bounded staging buffer so
// that the resident NTT-native data stays untouched)
int copy_out_cts(omr_ctx* ctx, uint64_t* h_out, const u64* d_src, size_t n, cudaStream_t s) {
    if (!n) return OMR_OK;
    if (ctx->out_domain == OMR_OUT_NTT_NATIVE) {
        CK(cudaMemcpyAsync(h_out, d_src, n * OMR_PV_WORDS * sizeof(u64), cudaMemcpyDeviceToHost, s));
        return OMR_OK;
    }
    const size_t STAGE = 1024;                                   // 32 MiB
    if (ctx->coeff_words < STAGE * OMR_PV_WORDS) {
        cudaFree(ctx->s_coeff); ctx->s_coeff = nullptr; ctx->coeff_words = 0;
        int st; if ((st = dalloc(ctx, &ctx->s_coeff, STAGE * OMR_PV_WORDS))) return st;
        ctx->coeff_words = STAGE * OMR_PV_WORDS;
    }
    for (size_t off = 0; off < n; off += STAGE) {
        const size_t nb = n - off < STAGE ? n - off : STAGE;
        CK(cudaMemcpyAsync(ctx->s_coeff, d_src + off * OMR_PV_WORDS, nb * OMR_PV_WORDS * sizeof(u64), cudaMemcpyDeviceToDevice, s));
        int st; if ((st = to_coeff_impl(ctx, ctx->s_coeff, 2 * nb, s))) return st;
        CK(cudaMemcpyAsync(h_out + off * OMR_PV_WORDS, ctx->s_coeff, nb * OMR_PV_WORDS * sizeof(u64), cudaMemcpyDeviceToHost, s));
    }
    return OMR_OK;
}
int ensure_elems(omr_ctx* ctx, unsigned short** p, size_t* have, size_t need) {
    if (need <= *have) return OMR_OK;
    cudaFree(*p); *p = nullptr; *have = 0;
    int st; if ((st = dalloc(ctx, p, need))) return st;
    *have = need;
    return OMR_OK;
}
int ensure_clues(omr_ctx* ctx, size_t B) {
    if (B <= ctx->clue_cap) return OMR_OK;
    cudaFree(ctx->s_ca); cudaFree(ctx->s_cb); ctx->s_ca = nullptr; ctx->s_cb = nullptr; ctx->clue_cap = 0;
    int st;
    if ((st = dalloc(ctx, &ctx->s_ca, B * CLUE_N))) return st;
    if ((st = dalloc(ctx, &ctx->s_cb, B * CLUE_COUNT))) return st;
    ctx->clue_cap = B;
    return OMR_OK;
}

}  // namespace

extern "C" {

// ---- device-pointer forms ----------------------------------------------------------------------------------------------
int omr_detect_batch_device(omr_ctx* ctx, const uint16_t* d_clue_a, const uint16_t* d_clue_b, size_t B, uint64_t* d_pv, void* stream,
                            omr_stage_times* times) {
    if (!ctx || (B && (!d_clue_a || !d_clue_b || !d_pv))) { ctx_fail(ctx, "detect: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    return detect_device(ctx, d_clue_a, d_clue_b, B, (u64*)d_pv, (cudaStream_t)stream, times);
}
int omr_l1_blind_rotate_device(omr_ctx* ctx, const uint16_t* a, const uint16_t* b, size_t B, uint32_t* d_rlwe, void* stream) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    return launch_l1(ctx, a, b, B, d_rlwe, (cudaStream_t)stream);
}
int omr_keyswitch_device(omr_ctx* ctx, const uint32_t* d_rlwe, size_t B, uint32_t* d_lwe, void* stream) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    return launch_ks(ctx, d_rlwe, B, d_lwe, (cudaStream_t)stream);
}
int omr_l2_blind_rotate_device(omr_ctx* ctx, const uint32_t* d_lwe, size_t B, uint64_t* d_rlwe, void* stream) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    return launch_l2(ctx, d_lwe, B, (u64*)d_rlwe, (cudaStream_t)stream);
}
int omr_trace_device(omr_ctx* ctx, uint64_t* d_rlwe, size_t B, void* stream) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    return launch_trace(ctx, (u64*)d_rlwe, B, (cudaStream_t)stream);
}
int omr_ntt_forward_device(omr_ctx* ctx, int level, void* d, size_t batch, void* stream) {
    if (!ctx || (level != 1 && level != 2)) return OMR_ERR_INVALID;
    ENTER(ctx);
    cudaStream_t s = (cudaStream_t)stream;
    if (!batch) return OMR_OK;
    if (level == 1) ntt_kernel<F1, false><<<(unsigned)batch, ntt_kernel_threads<F1>(), ntt_kernel_smem<F1>(), s>>>((u32*)d, ctx->tb, ctx->n1_inv);
    else ntt_kernel<F2, false><<<(unsigned)batch, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>((u64*)d, ctx->tb, ctx->n2_inv);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}
int omr_ntt_inverse_device(omr_ctx* ctx, int level, void* d, size_t batch, void* stream) {
    if (!ctx || (level != 1 && level != 2)) return OMR_ERR_INVALID;
    ENTER(ctx);
    cudaStream_t s = (cudaStream_t)stream;
    if (!batch) return OMR_OK;
    if (level == 1) ntt_kernel<F1, true><<<(unsigned)batch, ntt_kernel_threads<F1>(), ntt_kernel_smem<F1>(), s>>>((u32*)d, ctx->tb, ctx->n1_inv);
    else ntt_kernel<F2, true><<<(unsigned)batch, ntt_kernel_threads<F2>(), ntt_kernel_smem<F2>(), s>>>((u64*)d, ctx->tb, ctx->n2_inv);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

int omr_encode_indices_device(omr_ctx* ctx, const omr_retrieval_params* rp, const uint64_t* d_pv, size_t count, uint64_t index0,
                              uint64_t seed, uint32_t cipher_idx0, uint32_t n_cipher, uint64_t* d_out, void* stream) {
    if (!ctx || !d_out || (count && !d_pv)) { ctx_fail(ctx, "encode_indices: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    return encode_indices_impl(ctx, rp, (const u64*)d_pv, count, index0, seed, cipher_idx0, n_cipher, (u64*)d_out, (cudaStream_t)stream);
}
int omr_encode_payloads_device(omr_ctx* ctx, const uint64_t* d_pv, const uint16_t* d_payloads, size_t count, uint64_t index0,
                               const uint16_t* d_weights, size_t weight_stride, uint32_t n_cipher, uint32_t cmb_per_cipher, uint64_t* d_out,
                               void* stream) {
    if (!ctx || !d_out || (count && (!d_pv || !d_payloads || !d_weights))) { ctx_fail(ctx, "encode_payloads: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    return encode_payloads_impl(ctx, (const u64*)d_pv, d_payloads, count, index0, d_weights, weight_stride, n_cipher, cmb_per_cipher, (u64*)d_out,
                                (cudaStream_t)stream);
}
int omr_digest_reduce_mod(omr_ctx* ctx, uint64_t* d_words, size_t n, void* stream) {
    if (!ctx || (n && !d_words)) return OMR_ERR_INVALID;
    ENTER(ctx);
    if (!n) return OMR_OK;
    digest_mod_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((u64*)d_words, n);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

int omr_digest_add_mod(omr_ctx* ctx, uint64_t* d_acc, const uint64_t* d_part, size_t n, void* stream) {
    if (!ctx || (n && (!d_acc || !d_part))) return OMR_ERR_INVALID;
    ENTER(ctx);
    if (!n) return OMR_OK;
    digest_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((u64*)d_acc, (const u64*)d_part, n);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

int omr_decrypt_decode_device(omr_ctx* ctx, const uint64_t* d_z2_ntt, const uint64_t* d_ct, size_t n, uint16_t* d_out, void* stream) {
    if (!ctx || (n && (!d_z2_ntt || !d_ct || !d_out))) { ctx_fail(ctx, "decrypt_decode: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    return decrypt_decode_impl(ctx, (const u64*)d_z2_ntt, (const u64*)d_ct, n, d_out, (cudaStream_t)stream);
}

int omr_weights_from_seed_device(omr_ctx* ctx, const uint8_t* seed32, size_t count, uint16_t* d_out, uint32_t flags, void* stream) {
    if (!ctx || (count && (!seed32 || !d_out))) { ctx_fail(ctx, "weights_from_seed: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    return weights_from_seed_impl(ctx, seed32, count, d_out, flags, (cudaStream_t)stream);
}

int omr_gen_clues_device(omr_ctx* ctx, const uint16_t* d_pa, const uint16_t* d_pb, const uint8_t* seed32, uint64_t index0, size_t count,
                         const uint8_t* d_msgs, uint16_t* d_a, uint16_t* d_b, void* stream) {
    if (!ctx || !seed32 || (count && (!d_pa || !d_pb || !d_a || !d_b))) { ctx_fail(ctx, "gen_clues: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    if (!count) return OMR_OK;
    clue_gen_kernel<<<(unsigned)count, CLUE_THREADS, 0, (cudaStream_t)stream>>>(d_pa, d_pb, chacha_key_from_seed(seed32), index0, d_msgs, d_a, d_b);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

// ---- host-buffer forms ---------------------------------------------------------------------------------------------------
int omr_pv_reset(omr_ctx* ctx) {
    if (!ctx) return OMR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->pv_count = 0; ctx->pv_any = false; ctx->pv_index0 = 0;
    return OMR_OK;
}

int omr_pv_load(omr_ctx* ctx, const uint64_t* pv, size_t count, uint64_t global_index0) {
    if (!ctx || (count && !pv)) { ctx_fail(ctx, "pv_load: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    ctx->pv_count = 0; ctx->pv_any = false; ctx->pv_index0 = global_index0;
    if (!count) return OMR_OK;
    int st; if ((st = ensure_pv(ctx, count))) return st;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->pv, pv, count * OMR_PV_WORDS * sizeof(u64), cudaMemcpyHostToDevice, s));
    if (ctx->out_domain == OMR_OUT_COEFF && (st = from_coeff_impl(ctx, ctx->pv, 2 * count, s))) return st;
    CK(cudaStreamSynchronize(s));
    ctx->pv_count = count; ctx->pv_any = true;
    return OMR_OK;
}

int omr_detect_batch(omr_ctx* ctx, const uint16_t* clue_a, const uint16_t* clue_b, size_t B, uint64_t global_index0, uint64_t* pv_out,
                     omr_stage_times* times) {
    if (!ctx || (B && (!clue_a || !clue_b))) { ctx_fail(ctx, "detect: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    if (times) *times = omr_stage_times{};
    if (!B) return OMR_OK;
    // the store holds one contiguous run of global indices
    if (!ctx->pv_any) { ctx->pv_index0 = global_index0; ctx->pv_count = 0; }
    if (global_index0 != ctx->pv_index0 + ctx->pv_count) { ctx_fail(ctx, "detect: global_index0 must continue the pertinency store"); return OMR_ERR_STATE; }
    int st;
    if ((st = ensure_pv(ctx, ctx->pv_count + B))) return st;
    if ((st = ensure_clues(ctx, B))) return st;
    cudaStream_t s = ctx->stream;
    // all clues of the call are staged at once (1 038 bytes per message), so the chunks of detect_device run back to back
    // with no host synchronisation in between
    CK(cudaMemcpyAsync(ctx->s_ca, clue_a, B * CLUE_N * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->s_cb, clue_b, B * CLUE_COUNT * 2, cudaMemcpyHostToDevice, s));
    u64* dst = ctx->pv + ctx->pv_count * OMR_PV_WORDS;
    if ((st = detect_device(ctx, ctx->s_ca, ctx->s_cb, B, dst, s, times))) return st;
    if (pv_out && (st = copy_out_cts(ctx, pv_out, dst, B, s))) return st;
    CK(cudaStreamSynchronize(s));
    ctx->pv_count += B; ctx->pv_any = true;
    return OMR_OK;
}

int omr_encode_indices(omr_ctx* ctx, const omr_retrieval_params* rp, uint64_t seed, uint32_t cipher_idx0, uint32_t n_cipher, uint64_t* out) {
    if (!ctx || !out) { ctx_fail(ctx, "encode_indices: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);                                       // held across the launches AND the copy out of the shared digest buffer
    if (!ctx->pv_any) { ctx_fail(ctx, "encode_indices: empty pertinency store"); return OMR_ERR_STATE; }
    int st;
    if ((st = ensure_digest(ctx, (size_t)n_cipher * OMR_PV_WORDS))) return st;
    if ((st = encode_indices_impl(ctx, rp, ctx->pv, ctx->pv_count, ctx->pv_index0, seed, cipher_idx0, n_cipher, ctx->s_digest, ctx->stream))) return st;
    if ((st = copy_out_cts(ctx, out, ctx->s_digest, n_cipher, ctx->stream))) return st;
    CK(cudaStreamSynchronize(ctx->stream));
    return OMR_OK;
}

int omr_encode_payloads(omr_ctx* ctx, const uint16_t* payloads, size_t count, const uint16_t* weights, size_t weight_rows, size_t weight_stride,
                        uint32_t n_cipher, uint32_t cmb_per_cipher, uint64_t* out) {
    if (!ctx || !out || !payloads || !weights) { ctx_fail(ctx, "encode_payloads: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    if (!ctx->pv_any) { ctx_fail(ctx, "encode_payloads: empty pertinency store"); return OMR_ERR_STATE; }
    if (count != ctx->pv_count) { ctx_fail(ctx, "encode_payloads: payload count != pertinency store size"); return OMR_ERR_INVALID; }
    const size_t rows = (size_t)n_cipher * cmb_per_cipher;
    // the reference allocates ceil(cc / per) * per rows and fills the first combination_count (detector.rs:370-387): the caller
    // passes the rows it has, the missing tail rows are zero
    if (weight_rows == 0 || weight_rows > rows) { ctx_fail(ctx, "encode_payloads: weight_rows must be in [1, n_cipher * cmb_per_cipher]"); return OMR_ERR_INVALID; }
    int st;
    if ((st = ensure_digest(ctx, (size_t)n_cipher * OMR_PV_WORDS))) return st;
    const size_t pe = count * OMR_PAYLOAD_LEN, we = rows * weight_stride;
    if ((st = ensure_elems(ctx, &ctx->s_payloads, &ctx->payload_elems, pe))) return st;
    if ((st = ensure_elems(ctx, &ctx->s_weights, &ctx->weight_elems, we))) return st;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->s_payloads, payloads, pe * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->s_weights, weights, weight_rows * weight_stride * 2, cudaMemcpyHostToDevice, s));
    if (weight_rows < rows) CK(cudaMemsetAsync(ctx->s_weights + weight_rows * weight_stride, 0, (rows - weight_rows) * weight_stride * 2, s));
    if ((st = encode_payloads_impl(ctx, ctx->pv, ctx->s_payloads, count, ctx->pv_index0, ctx->s_weights, weight_stride, n_cipher, cmb_per_cipher,
                                   ctx->s_digest, s))) return st;
    if ((st = copy_out_cts(ctx, out, ctx->s_digest, n_cipher, s))) return st;
    CK(cudaStreamSynchronize(s));
    return OMR_OK;
}

// encode_pertinent_payloads with the reference's own argument: the rng seed instead of a weight matrix (detector.rs:341-453)
int omr_encode_payloads_seeded(omr_ctx* ctx, const uint16_t* payloads, size_t count, const uint8_t* seed32, uint64_t all_payloads_count,
                               uint32_t combination_count, uint32_t cmb_per_cipher, uint64_t* out) {
    if (!ctx || !out || !payloads || !seed32 || !cmb_per_cipher || !combination_count) { ctx_fail(ctx, "encode_payloads_seeded: bad argument"); return OMR_ERR_INVALID; }
    const uint32_t n_cipher = (combination_count + cmb_per_cipher - 1) / cmb_per_cipher;
    const size_t rows = (size_t)n_cipher * cmb_per_cipher, we = rows * all_payloads_count, pe = count * OMR_PAYLOAD_LEN;
    ENTER(ctx);
    if (!ctx->pv_any) { ctx_fail(ctx, "encode_payloads: empty pertinency store"); return OMR_ERR_STATE; }
    if (count != ctx->pv_count) { ctx_fail(ctx, "encode_payloads: payload count != pertinency store size"); return OMR_ERR_INVALID; }
    if (ctx->pv_index0 + count > all_payloads_count) { ctx_fail(ctx, "encode_payloads: store exceeds all_payloads_count"); return OMR_ERR_INVALID; }
    int st;
    if ((st = ensure_digest(ctx, (size_t)n_cipher * OMR_PV_WORDS))) return st;
    if ((st = ensure_elems(ctx, &ctx->s_payloads, &ctx->payload_elems, pe))) return st;
    if ((st = ensure_elems(ctx, &ctx->s_weights, &ctx->weight_elems, we))) return st;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->s_payloads, payloads, pe * 2, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->s_weights, 0, we * 2, s));                      // unused tail rows stay zero (detector.rs:370-371)
    if ((st = weights_from_seed_impl(ctx, seed32, (size_t)combination_count * all_payloads_count, ctx->s_weights, 0, s))) return st;
    if ((st = encode_payloads_impl(ctx, ctx->pv, ctx->s_payloads, count, ctx->pv_index0, ctx->s_weights, all_payloads_count, n_cipher, cmb_per_cipher,
                                   ctx->s_digest, s))) return st;
    if ((st = copy_out_cts(ctx, out, ctx->s_digest, n_cipher, s))) return st;
    CK(cudaStreamSynchronize(s));
    return OMR_OK;
}

// ---- recipient side, host-buffer form (Retriever::decode_digest, retriever.rs:188-260) -------------------------------------
}  // extern "C"
namespace {
// solve_matrix_mod_257 (matrix.rs:164-247): Gaussian elimination over Z_257 on (m [rows][cols], pl [rows][612]); first
// non-zero pivot, row swap, normalise, eliminate below, back-substitute.  false = singular (OmrError::InvertibleMatrix).
bool solve_mod_257(std::vector<u32>& m, std::vector<u32>& pl, size_t rows, size_t cols) {
    constexpr u32 P = OUT_P; constexpr size_t W = PAYLOAD_LEN;
    if (rows < cols) return false;
    u32 inv[P]; inv[0] = 0; inv[1] = 1;
    for (u32 v = 2; v < P; ++v) inv[v] = (P - (P / v) * inv[P % v] % P) % P;          // INV_MOD_257 (matrix.rs:28-41)
    for (size_t i = 0; i < cols; ++i) {
        size_t pick = i;
        while (pick < rows && m[pick * cols + i] == 0) ++pick;
        if (pick == rows) return false;                                               // matrix.rs:181-183
        if (pick != i) {
            for (size_t c = 0; c < cols; ++c) std::swap(m[i * cols + c], m[pick * cols + c]);
            for (size_t c = 0; c < W; ++c) std::swap(pl[i * W + c], pl[pick * W + c]);
        }
        const u32 iv = inv[m[i * cols + i]];
        if (iv != 1) {
            for (size_t c = i; c < cols; ++c) m[i * cols + c] = m[i * cols + c] * iv % P;
            for (size_t c = 0; c < W; ++c) pl[i * W + c] = pl[i * W + c] * iv % P;
        }
        if (i == cols - 1) break;
        for (size_t r = i + 1; r < rows; ++r) {
            const u32 f = m[r * cols + i];
            if (!f) continue;
            for (size_t c = i; c < cols; ++c) m[r * cols + c] = (m[r * cols + c] + (P - f) * m[i * cols + c]) % P;
            for (size_t c = 0; c < W; ++c) pl[r * W + c] = (pl[r * W + c] + (P - f) * pl[i * W + c]) % P;
        }
    }
    for (size_t ic = cols; ic-- > 1;)
        for (size_t r = 0; r < ic; ++r) {
            const u32 f = m[r * cols + ic];
            if (!f) continue;
            for (size_t c = 0; c < W; ++c) pl[r * W + c] = (pl[r * W + c] + (P - f) * pl[ic * W + c]) % P;
            m[r * cols + ic] = 0;
        }
    return true;
}
}  // namespace
extern "C" {

int omr_decode_digest(omr_ctx* ctx, const omr_retrieval_params* rp, const uint64_t* z2, const uint64_t* index_cts, uint32_t n_index_cts,
                      const uint64_t* payload_cts, uint32_t n_payload_cts, const uint16_t* weights, size_t weight_stride,
                      uint64_t* indices_out, uint32_t* n_found, uint16_t* payloads_out) {
    if (!ctx || !rp || !z2 || !index_cts || !payload_cts || !weights || !indices_out || !n_found || !payloads_out) {
        ctx_fail(ctx, "decode_digest: null argument"); return OMR_ERR_INVALID;
    }
    *n_found = 0;
    const size_t w = rp->slots_per_bucket, S = rp->slots_per_segment, cc = rp->combination_count, per = rp->cmb_count_per_cipher;
    if (rp->polynomial_size != OMR_N2 || rp->index_modulus != OMR_P || w < 2 || S == 0 || S % w || S > OMR_N2 || per == 0 ||
        per * OMR_PAYLOAD_LEN > OMR_N2 || (size_t)n_payload_cts * per < cc || weight_stride < rp->all_payloads_count) {
        ctx_fail(ctx, "decode_digest: inconsistent retrieval parameters"); return OMR_ERR_INVALID;
    }
    const size_t n = (size_t)n_index_cts + n_payload_cts;
    std::vector<uint16_t> slots(n * OMR_N2);
    {   // decrypt + inverse NTT + exact-integer decode of every slot on the GPU
        ENTER(ctx);
        u64 *d_key = nullptr, *d_ct = nullptr; uint16_t* d_out = nullptr;
        auto release = [&]() { cudaFree(d_key); cudaFree(d_ct); cudaFree(d_out); };
        cudaStream_t s = ctx->stream;
        cudaError_t e = cudaMalloc((void**)&d_key, OMR_N2 * sizeof(u64));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_ct, n * OMR_PV_WORDS * sizeof(u64));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_out, n * OMR_N2 * sizeof(uint16_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_key, z2, OMR_N2 * sizeof(u64), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ct, index_cts, (size_t)n_index_cts * OMR_PV_WORDS * sizeof(u64), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_ct + (size_t)n_index_cts * OMR_PV_WORDS, payload_cts, (size_t)n_payload_cts * OMR_PV_WORDS * sizeof(u64), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) { release(); ctx_fail(ctx, std::string("decode_digest: ") + cudaGetErrorString(e)); return OMR_ERR_CUDA; }
        int st = OMR_OK;
        if (ctx->out_domain == OMR_OUT_COEFF) {                      // coefficient-form secret and ciphertexts: transform on the way in
            st = from_coeff_impl(ctx, d_key, 1, s);
            if (st == OMR_OK) st = from_coeff_impl(ctx, d_ct, 2 * n, s);
        }
        if (st == OMR_OK) st = decrypt_decode_impl(ctx, d_key, d_ct, n, d_out, s);
        if (st == OMR_OK) {
            e = cudaMemcpyAsync(slots.data(), d_out, slots.size() * sizeof(uint16_t), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) { ctx_fail(ctx, std::string("decode_digest: ") + cudaGetErrorString(e)); st = OMR_ERR_CUDA; }
        } else cudaStreamSynchronize(s);
        release();
        if (st) return st;
    }
    // decode_pertinent_indices (retriever.rs:63-130): a bucket counts when its flag slot decodes to exactly 1; the index is the
    // base-257 number in the other slots, most significant last; ciphertexts are consumed until the set is full (:200-204)
    std::vector<uint64_t> found;
    for (uint32_t c = 0; c < n_index_cts && found.size() != rp->pertinent_count; ++c) {
        const uint16_t* dec = slots.data() + (size_t)c * OMR_N2;
        for (size_t seg = 0; seg + S <= (size_t)OMR_N2; seg += S)
            for (size_t bk = 0; bk + w <= S; bk += w) {
                const uint16_t* bucket = dec + seg + bk;
                if (bucket[w - 1] != 1) continue;
                uint64_t idx = 0;
                for (size_t k = w - 1; k-- > 0;) idx = idx * OMR_P + bucket[k];
                if (std::find(found.begin(), found.end(), idx) == found.end()) found.push_back(idx);
            }
    }
    std::sort(found.begin(), found.end());
    if (found.size() > rp->pertinent_count) { ctx_fail(ctx, "decode_digest: more indices than pertinent_count"); return OMR_ERR_INVALID; }
    *n_found = (uint32_t)found.size();
    for (size_t i = 0; i < found.size(); ++i) indices_out[i] = found[i];
    if (found.empty()) return OMR_OK;
    for (uint64_t i : found)
        if (i >= rp->all_payloads_count) { ctx_fail(ctx, "decode_digest: decoded index outside the board"); return OMR_ERR_INVALID; }
    // combination matrix from the weights (retriever.rs:215-240) and the combined payloads (:318-362)
    const size_t cols = found.size();
    std::vector<u32> m(cc * cols), pl(cc * OMR_PAYLOAD_LEN);
    for (size_t r = 0; r < cc; ++r) {
        for (size_t k = 0; k < cols; ++k) m[r * cols + k] = weights[r * weight_stride + found[k]] % OMR_P;
        const uint16_t* dec = slots.data() + ((size_t)n_index_cts + r / per) * OMR_N2 + (r % per) * OMR_PAYLOAD_LEN;
        for (size_t k = 0; k < OMR_PAYLOAD_LEN; ++k) pl[r * OMR_PAYLOAD_LEN + k] = dec[k];
    }
    if (!solve_mod_257(m, pl, cc, cols)) { ctx_fail(ctx, "matrix is not invertible"); return OMR_ERR_INVALID; }   // error.rs:4-8
    for (size_t i = 0; i < cols * OMR_PAYLOAD_LEN; ++i) payloads_out[i] = (uint16_t)pl[i];
    return OMR_OK;
}

// ---- streaming detection (README.md:9 "the detector processes incoming messages on-the-fly") -----------------------
// A resident running digest: every push detects its messages and folds their index- and payload-digest contributions in,
// stream-ordered on the context's stream.  Inputs go through two pinned staging buffers, so the host fills one while the copy
// engine drains the other and there is no host synchronisation per chunk; omr_stream_snapshot is the only call that waits.
// Packing is a sum over messages, so the running digest is bit-identical to packing the whole board at once.
int omr_stream_begin(omr_ctx* ctx, const omr_retrieval_params* rp, uint64_t index_seed, const uint8_t* weight_seed32, uint64_t global_index0) {
    if (!ctx || !rp || !weight_seed32) { ctx_fail(ctx, "stream_begin: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    int st; if ((st = check_rp(ctx, rp))) return st;
    if (rp->cmb_count_per_cipher == 0 || rp->cmb_count_per_cipher * OMR_PAYLOAD_LEN > OMR_N2 || rp->combination_count == 0 ||
        global_index0 > rp->all_payloads_count) { ctx_fail(ctx, "stream_begin: bad combination layout"); return OMR_ERR_INVALID; }
    cudaStream_t s = ctx->stream;
    CK(cudaStreamSynchronize(s));
    auto& S = ctx->st;
    const uint32_t n_idx = rp->max_encode_indices_cipher_count, n_pay = (rp->combination_count + rp->cmb_count_per_cipher - 1) / rp->cmb_count_per_cipher;
    const size_t words = (size_t)(n_idx + n_pay) * OMR_PV_WORDS;
    const size_t weight_elems = (size_t)n_pay * rp->cmb_count_per_cipher * rp->all_payloads_count;
    // a new board with the same layout (the common case: one stream per bulletin-board epoch) keeps its buffers
    const bool reuse = S.digest && S.n_idx == n_idx && S.n_pay == n_pay && S.weight_elems == weight_elems;
    if (!reuse) {
        stream_release(ctx);
        S.n_idx = n_idx; S.n_pay = n_pay; S.weight_elems = weight_elems;
        if ((st = dalloc(ctx, &S.digest, words)) || (st = dalloc(ctx, &S.part, words)) || (st = dalloc(ctx, &S.pv, STREAM_CHUNK * OMR_PV_WORDS))) { stream_release(ctx); return st; }
        if ((st = dalloc(ctx, &S.weights, S.weight_elems ? S.weight_elems : 1))) { stream_release(ctx); return st; }
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaMallocHost((void**)&S.pinned[i], STREAM_CHUNK * STREAM_MSG_BYTES);
            if (e == cudaSuccess) e = cudaMalloc((void**)&S.dev[i], STREAM_CHUNK * STREAM_MSG_BYTES);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&S.copied[i], cudaEventDisableTiming);
            if (e != cudaSuccess) { stream_release(ctx); ctx_fail(ctx, std::string("stream_begin: ") + cudaGetErrorString(e)); return OMR_ERR_ALLOC; }
        }
    }
    S.active = false; S.rp = *rp; S.index_seed = index_seed; S.index0 = global_index0; S.count = 0; S.next = 0;
    CK(cudaMemsetAsync(S.digest, 0, words * sizeof(u64), s));
    CK(cudaMemsetAsync(S.weights, 0, S.weight_elems * 2, s));                // rows beyond combination_count stay zero (detector.rs:370-371)
    if ((st = weights_from_seed_impl(ctx, weight_seed32, (size_t)rp->combination_count * rp->all_payloads_count, S.weights, 0, s))) { stream_release(ctx); return st; }
    CK(cudaStreamSynchronize(s));
    S.active = true;
    return OMR_OK;
}

int omr_stream_push(omr_ctx* ctx, const uint16_t* clue_a, const uint16_t* clue_b, const uint16_t* payloads, size_t n) {
    if (!ctx || (n && (!clue_a || !clue_b || !payloads))) { ctx_fail(ctx, "stream_push: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    auto& S = ctx->st;
    if (!S.active) { ctx_fail(ctx, "stream_push: no stream (call omr_stream_begin)"); return OMR_ERR_STATE; }
    if (S.index0 + S.count + n > S.rp.all_payloads_count) { ctx_fail(ctx, "stream_push: more messages than all_payloads_count"); return OMR_ERR_INVALID; }
    cudaStream_t s = ctx->stream;
    int st;
    if ((st = ensure_scratch(ctx, n < STREAM_CHUNK ? n : STREAM_CHUNK))) return st;
    for (size_t off = 0; off < n; off += STREAM_CHUNK) {
        const size_t nb = n - off < STREAM_CHUNK ? n - off : STREAM_CHUNK;
        const int b = S.next; S.next ^= 1;
        CK(cudaEventSynchronize(S.copied[b]));                               // the copy out of this pinned buffer two chunks ago is done
        unsigned char* h = S.pinned[b];
        const size_t na = nb * CLUE_N * 2, nbb = nb * CLUE_COUNT * 2, np = nb * PAYLOAD_LEN * 2;
        memcpy(h, clue_a + off * CLUE_N, na); memcpy(h + na, clue_b + off * CLUE_COUNT, nbb); memcpy(h + na + nbb, payloads + off * PAYLOAD_LEN, np);
        CK(cudaMemcpyAsync(S.dev[b], h, na + nbb + np, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(S.copied[b], s));
        const unsigned short* d_a = reinterpret_cast<const unsigned short*>(S.dev[b]);
        const unsigned short* d_b = reinterpret_cast<const unsigned short*>(S.dev[b] + na);
        const unsigned short* d_p = reinterpret_cast<const unsigned short*>(S.dev[b] + na + nbb);
        const u64 gi = S.index0 + S.count;
        if ((st = detect_device(ctx, d_a, d_b, nb, S.pv, s, nullptr))) return st;
        if ((st = encode_indices_impl(ctx, &S.rp, S.pv, nb, gi, S.index_seed, 0, S.n_idx, S.part, s))) return st;
        if ((st = encode_payloads_impl(ctx, S.pv, d_p, nb, gi, S.weights, S.rp.all_payloads_count, S.n_pay, S.rp.cmb_count_per_cipher,
                                       S.part + (size_t)S.n_idx * OMR_PV_WORDS, s))) return st;
        const size_t words = (size_t)(S.n_idx + S.n_pay) * OMR_PV_WORDS;
        digest_add_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(S.digest, S.part, words);
        ++ctx->launches; CK(cudaGetLastError());
        S.count += nb;
    }
    return OMR_OK;
}

int omr_stream_snapshot(omr_ctx* ctx, uint64_t* out, uint64_t* n_messages) {
    if (!ctx || !out) { ctx_fail(ctx, "stream_snapshot: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    auto& S = ctx->st;
    if (!S.active) { ctx_fail(ctx, "stream_snapshot: no stream (call omr_stream_begin)"); return OMR_ERR_STATE; }
    int st; if ((st = copy_out_cts(ctx, out, S.digest, S.n_idx + S.n_pay, ctx->stream))) return st;
    CK(cudaStreamSynchronize(ctx->stream));
    if (n_messages) *n_messages = S.count;
    return OMR_OK;
}

int omr_stream_end(omr_ctx* ctx) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    CK(cudaStreamSynchronize(ctx->stream));
    stream_release(ctx);
    return OMR_OK;
}

// ---- K7: cross-GPU sum of the partial digests (the rayon reduce + add_element_wise of detector.rs:333-336, 445-448) ---------------
int omr_comm_unique_id(uint8_t* id128) {
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!id128 || !api) { omr::set_global_error(api ? "comm_unique_id: null argument" : why); return api ? OMR_ERR_INVALID : OMR_ERR_STATE; }
    NcclUniqueId id;
    const int rc = api->get_unique_id(&id);
    if (rc) { omr::set_global_error(std::string("ncclGetUniqueId: ") + api->error_string(rc)); return OMR_ERR_CUDA; }
    memcpy(id128, id.internal, sizeof id.internal);
    return OMR_OK;
}
int omr_comm_init(omr_ctx* ctx, int n_ranks, int rank, const uint8_t* id128) {
    if (!ctx || !id128 || n_ranks < 1 || rank < 0 || rank >= n_ranks) { ctx_fail(ctx, "comm_init: bad argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) { ctx_fail(ctx, why); return OMR_ERR_STATE; }
    if (ctx->comm) { api->comm_destroy(ctx->comm); ctx->comm = nullptr; }
    NcclUniqueId id; memcpy(id.internal, id128, sizeof id.internal);
    const int rc = api->comm_init_rank(&ctx->comm, n_ranks, id, rank);
    if (rc) { ctx->comm = nullptr; ctx_fail(ctx, std::string("ncclCommInitRank: ") + api->error_string(rc)); return OMR_ERR_CUDA; }
    return OMR_OK;
}
int omr_comm_destroy(omr_ctx* ctx) {
    if (!ctx) return OMR_ERR_INVALID;
    ENTER(ctx);
    if (ctx->comm) { if (const NcclApi* api = nccl_api(nullptr)) api->comm_destroy(ctx->comm); ctx->comm = nullptr; }
    return OMR_OK;
}
// in place: d_digests [n_cipher][2][2048] canonical partial digests -> their sum over all ranks mod q2.  Values < q2 < 2^50, so a
// raw u64 sum over up to 2^13 ranks cannot overflow; one small kernel reduces mod q2 right behind the collective on the stream.
int omr_digest_allreduce(omr_ctx* ctx, void* nccl_comm, uint64_t* d_digests, size_t n_cipher, void* stream) {
    if (!ctx || (n_cipher && !d_digests)) { ctx_fail(ctx, "digest_allreduce: null argument"); return OMR_ERR_INVALID; }
    ENTER(ctx);
    void* comm = nccl_comm ? nccl_comm : ctx->comm;
    if (!comm) { ctx_fail(ctx, "digest_allreduce: no communicator (pass an ncclComm_t or call omr_comm_init)"); return OMR_ERR_STATE; }
    std::string why;
    const NcclApi* api = nccl_api(&why);
    if (!api) { ctx_fail(ctx, why); return OMR_ERR_STATE; }
    if (!n_cipher) return OMR_OK;
    const size_t words = n_cipher * OMR_PV_WORDS;
    const int rc = api->all_reduce(d_digests, d_digests, words, NcclApi::UINT64, NcclApi::SUM, comm, (cudaStream_t)stream);
    if (rc) { ctx_fail(ctx, std::string("ncclAllReduce: ") + api->error_string(rc)); return OMR_ERR_CUDA; }
    digest_mod_kernel<<<(unsigned)((words + 255) / 256), 256, 0, (cudaStream_t)stream>>>((u64*)d_digests, words);
    ++ctx->launches; CK(cudaGetLastError());
    return OMR_OK;
}

int omr_mulmod_peak(omr_ctx* ctx, int level, int iters, double* mulmods_per_second) {
    if (!ctx || !mulmods_per_second || level < 1 || level > 3 || iters <= 0) return OMR_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu); CK(cudaSetDevice(ctx->device));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    int st; if ((st = ensure_digest(ctx, (size_t)blocks * threads))) return st;
    cudaStream_t s = ctx->stream;
    for (int rep = 0; rep < 2; ++rep) {          // first run warms up
        CK(cudaEventRecord(ctx->ev[0], s));
        if (level == 1) mulmod_peak_kernel<F1><<<blocks, threads, 0, s>>>((u32*)ctx->s_digest, ctx->n1_inv, iters);
        else if (level == 2) mulmod_peak_kernel<F2><<<blocks, threads, 0, s>>>((u64*)ctx->s_digest, ctx->n2_inv, iters);
        else mulmod_peak_f64_kernel<<<blocks, threads, 0, s>>>((double*)ctx->s_digest, make_double2(123456789012345.0, 123456789012345.0 / 1125899906826241.0), iters);
        ++ctx->launches; CK(cudaGetLastError());
        CK(cudaEventRecord(ctx->ev[1], s)); CK(cudaEventSynchronize(ctx->ev[1]));
    }
    float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    *mulmods_per_second = (double)blocks * threads * 8.0 * iters / (ms * 1e-3);
    return OMR_OK;
}

}  // extern "C"

// kernels.cuh — sm_100a kernels of the InstantOMR detection hot path.
//   K1 l1_blind_rotate_kernel    detector.rs:505-531, 553-557   7 blind rotations over R_q1 + sum
//   K2 keyswitch_kernel          detector.rs:560-596            sample extract, LWE key switch, mod switch, offset
//   K3 l2_blind_rotate_kernel    detector.rs:599-624            blind rotation over R_q2 (FP64 pipe, exact)
//   K4 trace_kernel              detector.rs:626-639            *N^-1, homomorphic trace, to NTT form
//   K5 pack_kernel<INDICES>      detector.rs:223-339            index digest
//   K6 pack_kernel<PAYLOADS>     detector.rs:341-453            payload digest
// plus partial-sum reduction, batched standalone NTTs and the key pre-transforms.
// Everything here runs on the CUDA cores (integer and FP64 pipes): the contraction is an exact modular NTT, not a floating-point
// GEMM.  The one GEMM-shaped stage, the LWE key switch, has an opt-in tensor-core variant (ks_gemm.cu).
#pragma once
#include "ntt.cuh"

namespace omr {

constexpr int CLUE_N = 512, CLUE_COUNT = 7, CLUE_Q = 2048;
constexpr int LWE2_N = 670, LWE2_Q = 4096, LWE2_STRIDE_IN = 671;
constexpr int KSK_PAD = 672;                   // device row stride of the key-switching key (671 padded)
constexpr int KS_LEVELS = 27;
constexpr int TR_STEPS = 11, TR_LEVELS = 25;
constexpr int PAYLOAD_LEN = 612;
constexpr u32 OUT_P = 257;

// gadget bases (parameters/mod.rs:55,81,89); SURVEY A.4: keep the LEVELS most significant base-2^LOGB digits
struct G1 { static constexpr int LOGB = 5, LEVELS = 4, DROP = 27 - 20; };   // NonPowOf2ApproxSignedBasis(q1,5,Some(4))
struct G2 { static constexpr int LOGB = 7, LEVELS = 6, DROP = 50 - 42; };   // (q2,7,Some(6))
struct GT { static constexpr int LOGB = 2, LEVELS = 25, DROP = 0; };        // (q2,2,None)

struct Tables {
    const uint2* tw1; const uint2* itw1;            // [1024] bit-reversed psi powers with Shoup companions
    const ulonglong2* tw2; const ulonglong2* itw2;  // [2048]
    const double2* tw2d; const double2* itw2d;      // [2048] the same powers, centred, as (w, w/q2) doubles (FP64 path)
    const u32* lut1; const u64* lut2;               // test vectors (detector.rs:457-503)
    ulonglong2 n2_inv;                              // N2^-1 mod q2 (secret.rs:167-176), Shoup pair
    ulonglong2 r2;                                  // 2^64 mod q2, Shoup pair (undo REDC in the packing kernels)
    u32 trace_dinv[TR_STEPS];                       // (2^k+1)^-1 mod 2*N2, k = 11..1
};

// ---- signed gadget decomposition --------------------------------------------------------------------
// offset word: u = round(v / 2^DROP) + SUM_{j<L-1} (B/2) B^j ; digit j<L-1 = ((u >> wj) & (B-1)) - B/2 ; top = u >> w(L-1)
template <class F, class G> __device__ __forceinline__ typename F::S gadget_word(typename F::S v) {
    typedef typename F::S S;
    S c = 0;
#pragma unroll
    for (int j = 0; j < G::LEVELS - 1; ++j) c += (S)(1 << (G::LOGB - 1)) << (G::LOGB * j);
    if (G::DROP > 0) v = (v + ((S)1 << (G::DROP > 0 ? G::DROP - 1 : 0))) >> G::DROP;
    return v + c;
}
// digit r of offset word u as a small signed integer
template <class F, class G> __device__ __forceinline__ int gadget_digit_signed(typename F::S u, int r) {
    typedef typename F::S S;
    constexpr S B = (S)1 << G::LOGB;
    return (int)((r < G::LEVELS - 1) ? (((u >> (G::LOGB * r)) & (B - 1)) - (B >> 1)) : (u >> (G::LOGB * (G::LEVELS - 1))));
}
// ... and as a lazy field element in (0, 2q)
template <class F, class G> __device__ __forceinline__ typename F::T gadget_digit(typename F::S u, int r) {
    return (typename F::T)((typename F::S)F::Q + (typename F::S)gadget_digit_signed<F, G>(u, r));
}
// centre x in (-2q, q) (a signed difference of two canonical values) to [-(q-1)/2, (q-1)/2]
template <class F> __device__ __forceinline__ typename F::S centre_diff(typename F::S w) {
    typedef typename F::S S;
    constexpr S Q = (S)F::Q, H = (S)(F::Q >> 1);
    if (w < -H) w += Q;
    if (w < -H) w += Q;
    if (w > H) w -= Q;
    return w;
}
// value of (X^a * p)[pos] for a in [0, 2N): signed (negated when the rotation wraps an odd number of times)
template <class F> __device__ __forceinline__ typename F::S rotated(const typename F::T* p, int pos, int a) {
    typedef typename F::S S;
    int s = pos - a; bool neg = false;
    if (s < 0) { s += F::N; neg = !neg; }
    if (s < 0) { s += F::N; neg = !neg; }
    S r = (S)p[s];
    return neg ? -r : r;
}
// offset words of ((X^a - 1) * p) at the pass-0 positions of thread t
template <class F, class G, class GEO> __device__ __forceinline__ void decompose_words(typename F::S (&u)[GEO::E], const typename F::T* p, int a, int t) {
    typedef typename F::S S;
#pragma unroll
    for (int k = 0; k < GEO::E; ++k) {
        const int pos = t + GEO::NT * k;
        u[k] = gadget_word<F, G>(centre_diff<F>(rotated<F>(p, pos, a) - (S)p[pos]));
    }
}
// acc = (0, LUT * X^(2N - b))  — start of BlindRotationKey::blind_rotate (detector.rs:555,623)
template <class F, class GEO> __device__ __forceinline__ void init_acc(typename F::T* acc, const typename F::T* __restrict__ lut, int b, int t) {
    typedef typename F::S S;
    constexpr int N = F::N;
    const int rot = (2 * N - b) % (2 * N);
#pragma unroll
    for (int k = 0; k < GEO::E; ++k) {
        const int pos = t + GEO::NT * k;
        S v = rotated<F>(lut, pos, rot);
        acc[pos] = 0;
        acc[N + pos] = (typename F::T)(v < 0 ? v + (S)F::Q : v);
    }
}

// ---- TMA (cp.async.bulk) + mbarrier helpers ---------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n .reg .pred p;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra LAB_DONE;\n bra LAB_WAIT;\n LAB_DONE:\n }\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one elected thread: bulk copy `bytes` (multiple of 16) global -> shared, completion signalled on `bar`
__device__ __forceinline__ void tma_load(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// ---- K1: first-level blind rotations ---------------------------------------------------------------------------------
// A CTA runs 6 blind rotations at once — 6 consecutive (message, clue) pairs of the batch — one per group of 64 threads
// (2 warps, 16 coefficients per thread): 12 warps = 3 per SM partition at 168 registers, which holds the 32 u64 MAC
// accumulators AND two digit transforms in flight without the spills of the 16-warp / 128-register shape (measured: 8 groups
// 1 206 ms, 6 groups 1 155 ms, 6 groups with two digits per pass 1 120 ms per 16 384 messages).  Every blind rotation walks the
// same key sequence BSK1[0..511], so per CMux step the 64 KiB RGSW tile is staged ONCE into shared memory by a TMA bulk copy and
// reused by all groups; the copy of tile i+1 is in flight while the groups run their inverse transforms and the first
// forward transforms of the next step.  Twiddles live in shared memory.  Clue extraction (CmLwe::extract_all,
// detector.rs:514) is index arithmetic on the fly.  Output: one RLWE accumulator per (message, clue); sum7_kernel adds
// the 7 accumulators of a message (add_element_wise, detector.rs:556).
constexpr int L1_GROUP = GeoL1::NT;                        // 64
constexpr int L1_TILE_WORDS = 2 * G1::LEVELS * 2 * F1::N;  // 16384 u32 = 64 KiB
constexpr int L1_GROUP_WORDS = 2 * F1::N + 2 * GeoL1::BUF;
// SLOTS = blind rotations per CTA: 6 is the throughput shape (384 threads); 4 serves mid-size batches that could not give every SM a
// 6-rotation CTA (capi.cu: launch_l1_raw).
template <int SLOTS> struct L1Cfg {
    static constexpr int THREADS = SLOTS * L1_GROUP;
    // one CTA per SM: each of the 4 SM partitions holds ceil(warps / 4) warps and 16 384 registers (allocated 8 per thread at a time)
    static constexpr int WARPS_PER_PARTITION = (THREADS / 32 + 3) / 4;
    static constexpr int MAXREG = (16384 / (32 * WARPS_PER_PARTITION)) / 8 * 8 > 248 ? 248 : (16384 / (32 * WARPS_PER_PARTITION)) / 8 * 8;
    static constexpr int TILE_WORDS = L1_TILE_WORDS;
    static constexpr size_t SMEM = (size_t)TILE_WORDS * 4 + 2 * F1::N * sizeof(uint2) + (size_t)SLOTS * L1_GROUP_WORDS * 4 +
                                   (size_t)SLOTS * CLUE_N * sizeof(unsigned short) + 16;
};

static_assert(L1Cfg<6>::MAXREG == 168 && L1Cfg<4>::MAXREG == 248 && L1Cfg<8>::MAXREG == 128, "registers per thread follow the warps per SM partition");
static_assert(L1Cfg<6>::SMEM <= 227 * 1024, "one CTA per SM");
template <int SLOTS>
__global__ void __maxnreg__(L1Cfg<SLOTS>::MAXREG)
l1_blind_rotate_kernel(const unsigned short* __restrict__ clue_a, const unsigned short* __restrict__ clue_b,
                       const u32* __restrict__ bsk1, u32* __restrict__ out /*[n_clues][2][N]*/, int n_clues, Tables tb) {
    typedef F1 F; typedef G1 G; typedef GeoL1 GEO; typedef ArInt<F1> AR; typedef L1Cfg<SLOTS> CFG;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS, THREADS = CFG::THREADS;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    u32* ktile = reinterpret_cast<u32*>(smem_dyn);
    uint2* s_tw = reinterpret_cast<uint2*>(ktile + CFG::TILE_WORDS);
    uint2* s_itw_smem = s_tw + N;
    u32* groups = reinterpret_cast<u32*>(s_tw + 2 * N);
    unsigned short* ca_all = reinterpret_cast<unsigned short*>(groups + (size_t)SLOTS * L1_GROUP_WORDS);
    u64* mbar = reinterpret_cast<u64*>(ca_all + SLOTS * CLUE_N);

    const int slot = threadIdx.x / L1_GROUP, t = threadIdx.x % L1_GROUP;
    const int cid_raw = blockIdx.x * SLOTS + slot;
    const int cid = cid_raw < n_clues ? cid_raw : n_clues - 1;        // tail groups redo the last clue (no store)
    const int msg = cid / CLUE_COUNT, c = cid % CLUE_COUNT;
    u32* acc = groups + (size_t)slot * L1_GROUP_WORDS;       // [2][N]
    unsigned short* ca = ca_all + slot * CLUE_N;
    ExBuf<u32> eb{acc + 2 * N, acc + 2 * N + GEO::BUF};
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    fill_twiddles_deint<GEO>(s_tw, tb.tw1, threadIdx.x, THREADS); fill_twiddles_deint<GEO>(s_itw_smem, tb.itw1, threadIdx.x, THREADS);
    for (int i = t; i < CLUE_N; i += L1_GROUP) ca[i] = clue_a[(size_t)msg * CLUE_N + i] & (CLUE_Q - 1);   // canonical mod 2048
    init_acc<F, GEO>(acc, tb.lut1, clue_b[(size_t)msg * CLUE_COUNT + c] & (CLUE_Q - 1), t);
    __syncthreads();
    if (threadIdx.x == 0) tma_load(ktile, bsk1, CFG::TILE_WORDS * 4, mbar);
    const int bar = 1 + slot;
    u32 phase = 0;
#pragma unroll 1
    for (int i = 0; i < CLUE_N; ++i) {
        // a^(c)_i = a_{c-i} (i <= c), -a_{512+c-i} (i > c)   SURVEY A.3.   a == 0 is not skipped: (X^0 - 1) acc = 0
        // decomposes to all-zero digits and adds nothing, bit-identical to skipping (keeps the groups in lock step).
        const int a = i <= c ? ca[c - i] : ((CLUE_Q - ca[CLUE_N + c - i]) & (CLUE_Q - 1));
        u64 ma[E], mb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = 0; mb[k] = 0; }
        // Both digit loops are fully unrolled and the butterfly additions are three-input adds (F1::add_alu): together they take the
        // register moves of the u64 accumulators and the IMAD.IADD / IMAD.MOV forms ptxas likes off the heavy FMA pipe (per CMux step and
        // thread: 7 894 -> 6 904 instructions, 791 -> 44 adds / moves on that pipe; either change alone is re-balanced away by ptxas).
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            i32 u[E];
            decompose_words<F, G, GEO>(u, acc + p * N, a, t);
#pragma unroll
            for (int r = 0; r < L; r += 2) {
                u32 x[E], y[E];
#pragma unroll
                for (int k = 0; k < E; ++k) { x[k] = gadget_digit<F, G>(u[k], r); y[k] = gadget_digit<F, G>(u[k], r + 1); }
                ntt_forward2s<AR, GEO, LdSharedD>(x, y, eb.a, eb.b, s_tw, t, bar);      // two digits in flight: twice the ILP
                if (r == 0 && p == 0) { mbar_wait(mbar, phase & 1); ++phase; }             // tile i has landed
                const u32* ka = ktile + (size_t)(p * L + r) * 2 * N + out_idx<GEO>(t, 0);
                const u32* kb = ka + N;
#pragma unroll
                for (int k = 0; k < E; k += 4) {
                    const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
                    const uint4 va = *reinterpret_cast<const uint4*>(ka + o), vb = *reinterpret_cast<const uint4*>(kb + o);
                    F::mac(ma[k], x[k], va.x); F::mac(ma[k + 1], x[k + 1], va.y); F::mac(ma[k + 2], x[k + 2], va.z); F::mac(ma[k + 3], x[k + 3], va.w);
                    F::mac(mb[k], x[k], vb.x); F::mac(mb[k + 1], x[k + 1], vb.y); F::mac(mb[k + 2], x[k + 2], vb.z); F::mac(mb[k + 3], x[k + 3], vb.w);
                    const uint4 wa = *reinterpret_cast<const uint4*>(ka + 2 * N + o), wb = *reinterpret_cast<const uint4*>(kb + 2 * N + o);
                    F::mac(ma[k], y[k], wa.x); F::mac(ma[k + 1], y[k + 1], wa.y); F::mac(ma[k + 2], y[k + 2], wa.z); F::mac(ma[k + 3], y[k + 3], wa.w);
                    F::mac(mb[k], y[k], wb.x); F::mac(mb[k + 1], y[k + 1], wb.y); F::mac(mb[k + 2], y[k + 2], wb.z); F::mac(mb[k + 3], y[k + 3], wb.w);
                }
            }
        }
        u32 ya[E], yb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { ya[k] = F::inv_prepare(F::redc(ma[k])); yb[k] = F::inv_prepare(F::redc(mb[k])); }
        __syncthreads();                                                         // every group is done with tile i
        if (threadIdx.x == 0 && i + 1 < CLUE_N) tma_load(ktile, bsk1 + (size_t)(i + 1) * L1_TILE_WORDS, CFG::TILE_WORDS * 4, mbar);
        // both inverse transforms in flight (two buffers, staggered exchanges): twice the ILP of running them back to back
        ntt_inverse2s<AR, GEO, LdSharedDI>(ya, yb, eb.a, eb.b, s_itw_smem, t, bar);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = t + GEO::NT * k;
            acc[pos] = F::add_canon(acc[pos], ya[k]);
            acc[N + pos] = F::add_canon(acc[N + pos], yb[k]);
        }
        group_sync<GEO::NT>(bar);
    }
    if (cid_raw < n_clues) {
        u32* o = out + (size_t)cid * 2 * N;
#pragma unroll
        for (int k = 0; k < 2 * E; ++k) o[t + GEO::NT * k] = acc[t + GEO::NT * k];
    }
}
// K1, latency shape (fewer blind rotations than SMs): one CTA per (message, clue), 8 groups of 64 threads.  Group g
// transforms digit g & 3 of polynomial g >> 2 and multiplies it with its two key rows; the 8 partial products (REDC'd to
// < 2q) are summed slice-wise by all 512 threads, then groups 0 and 1 run the two inverse transforms.  The critical path of
// a CMux step is one forward and one inverse transform instead of eight and two; sums are exact mod q, so the accumulator
// is bit-identical to the throughput shape.
constexpr int L1L_GROUPS = 2 * G1::LEVELS, L1L_THREADS = L1L_GROUPS * L1_GROUP;
constexpr size_t L1L_SMEM = (size_t)L1_TILE_WORDS * 4 + 2 * F1::N * sizeof(uint2) + (size_t)4 * F1::N * 4 +
                            (size_t)L1L_GROUPS * 2 * GeoL1::BUF * 4 + CLUE_N * sizeof(unsigned short) + 16;
static_assert(2 * GeoL1::BUF >= 2 * F1::N, "a group's exchange buffers hold its two partial products");

__global__ void __launch_bounds__(L1L_THREADS, 1)
l1_blind_rotate_lat_kernel(const unsigned short* __restrict__ clue_a, const unsigned short* __restrict__ clue_b,
                           const u32* __restrict__ bsk1, u32* __restrict__ out /*[n_clues][2][N]*/, Tables tb) {
    typedef F1 F; typedef G1 G; typedef GeoL1 GEO; typedef ArInt<F1> AR;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    u32* ktile = reinterpret_cast<u32*>(smem_dyn);
    uint2* s_tw = reinterpret_cast<uint2*>(ktile + L1_TILE_WORDS);
    uint2* s_itw = s_tw + N;
    u32* acc = reinterpret_cast<u32*>(s_itw + N);            // [2][N]
    u32* tot = acc + 2 * N;                                  // [2][N] summed products of the step
    u32* bufs = tot + 2 * N;                                 // [8][2 * BUF]
    unsigned short* ca = reinterpret_cast<unsigned short*>(bufs + (size_t)L1L_GROUPS * 2 * GEO::BUF);
    u64* mbar = reinterpret_cast<u64*>(ca + CLUE_N);

    const int g = threadIdx.x / L1_GROUP, t = threadIdx.x % L1_GROUP, bar = 1 + g;
    const int cid = blockIdx.x, msg = cid / CLUE_COUNT, c = cid % CLUE_COUNT;
    const int p = g / L, r = g % L;
    u32* mine = bufs + (size_t)g * 2 * GEO::BUF;
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    fill_twiddles_deint<GEO>(s_tw, tb.tw1, threadIdx.x, L1L_THREADS); fill_twiddles_deint<GEO>(s_itw, tb.itw1, threadIdx.x, L1L_THREADS);
    for (int i = threadIdx.x; i < CLUE_N; i += L1L_THREADS) ca[i] = clue_a[(size_t)msg * CLUE_N + i] & (CLUE_Q - 1);
    if (g == 0) init_acc<F, GEO>(acc, tb.lut1, clue_b[(size_t)msg * CLUE_COUNT + c] & (CLUE_Q - 1), t);
    __syncthreads();
    if (threadIdx.x == 0) tma_load(ktile, bsk1, L1_TILE_WORDS * 4, mbar);
#pragma unroll 1
    for (int i = 0; i < CLUE_N; ++i) {
        const int a = i <= c ? ca[c - i] : ((CLUE_Q - ca[CLUE_N + c - i]) & (CLUE_Q - 1));
        i32 u[E];
        decompose_words<F, G, GEO>(u, acc + p * N, a, t);
        u32 x[E];
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = gadget_digit<F, G>(u[k], r);
        ExBuf<u32> eb{mine, mine + GEO::BUF};
        ntt_forward<AR, GEO, LdSharedD>(x, eb, s_tw, t, bar);
        mbar_wait(mbar, i & 1);                                                  // tile i has landed
        const u32* ka = ktile + (size_t)g * 2 * N + out_idx<GEO>(t, 0);
        const u32* kb = ka + N;
        group_sync<GEO::NT>(bar);                                                // the group's last exchange has been read
#pragma unroll
        for (int k = 0; k < E; k += 4) {
            const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
            const uint4 va = *reinterpret_cast<const uint4*>(ka + o), vb = *reinterpret_cast<const uint4*>(kb + o);
            mine[t + GEO::NT * k] = F::redc((u64)x[k] * va.x); mine[t + GEO::NT * (k + 1)] = F::redc((u64)x[k + 1] * va.y);
            mine[t + GEO::NT * (k + 2)] = F::redc((u64)x[k + 2] * va.z); mine[t + GEO::NT * (k + 3)] = F::redc((u64)x[k + 3] * va.w);
            mine[N + t + GEO::NT * k] = F::redc((u64)x[k] * vb.x); mine[N + t + GEO::NT * (k + 1)] = F::redc((u64)x[k + 1] * vb.y);
            mine[N + t + GEO::NT * (k + 2)] = F::redc((u64)x[k + 2] * vb.z); mine[N + t + GEO::NT * (k + 3)] = F::redc((u64)x[k + 3] * vb.w);
        }
        __syncthreads();                                                         // products visible; tile i is free
        if (threadIdx.x == 0 && i + 1 < CLUE_N) tma_load(ktile, bsk1 + (size_t)(i + 1) * L1_TILE_WORDS, L1_TILE_WORDS * 4, mbar);
#pragma unroll
        for (int j = 0; j < 2 * N / L1L_THREADS; ++j) {
            const int e = threadIdx.x + L1L_THREADS * j;
            u32 sum = 0;                                                         // 8 terms < 2q each
#pragma unroll
            for (int gg = 0; gg < L1L_GROUPS; ++gg) sum += bufs[(size_t)gg * 2 * GEO::BUF + e];
            tot[e] = F::inv_prepare(sum);
        }
        __syncthreads();
        if (g < 2) {
            u32 y[E];
#pragma unroll
            for (int k = 0; k < E; ++k) y[k] = tot[g * N + t + GEO::NT * k];
            ExBuf<u32> ei{mine, mine + GEO::BUF};
            ntt_inverse<AR, GEO, LdSharedDI>(y, ei, s_itw, t, bar);
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const int pos = g * N + t + GEO::NT * k;
                acc[pos] = F::add_canon(acc[pos], y[k]);
            }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < 2 * N; e += L1L_THREADS) out[(size_t)cid * 2 * N + e] = acc[e];
}

// sum of the 7 accumulators of each message (detector.rs:556)
__global__ void sum7_kernel(const u32* __restrict__ in /*[B*7][2][N]*/, u32* __restrict__ out /*[B][2][N]*/, size_t B) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= B * 2 * F1::N) return;
    const size_t m = e / (2 * F1::N), w = e % (2 * F1::N);
    u32 s = 0;
#pragma unroll
    for (int c = 0; c < CLUE_COUNT; ++c) s += in[(m * CLUE_COUNT + c) * 2 * F1::N + w];    // 7q < 2^30
    out[e] = s % Q1;
}

// ---- K3: second-level blind rotation, transforms and MAC on the FP64 pipe (D2 in field.cuh) -----------------------------
// One CTA (256 threads, 8 coefficients each) per message, two CTAs per SM.  acc stays canonical u64 in shared memory (decomposition is
// integer bit work); digits enter the NTT as doubles, two digits per pass (shared barriers, twice the ILP); key words are
// centred doubles already multiplied by N^-1, streamed from L2 with L1::no_allocate so the twiddle tables stay in L1;
// the MAC accumulators are 16 doubles.
constexpr int L2_THREADS = GeoL2::NT;
constexpr size_t L2_SMEM = (size_t)2 * F2::N * 8 + (size_t)2 * GeoL2::BUF * 8 + F2::N * sizeof(double2) + 672 * sizeof(unsigned short);

__device__ __forceinline__ double2 ld_stream_f64x2(const double* p) {
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(L2_THREADS, 2)
l2_blind_rotate_kernel(const u32* __restrict__ lwe, const double* __restrict__ bsk2, u64* __restrict__ out, Tables tb) {
    typedef F2 F; typedef G2 G; typedef GeoL2 GEO; typedef ArD2 AR;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    double* bx = reinterpret_cast<double*>(acc + 2 * N); double* by = bx + GEO::BUF;
    double2* s_tw = reinterpret_cast<double2*>(by + GEO::BUF);          // forward twiddles: the L1 cache left beside 2 x 100 KiB
    unsigned short* la = reinterpret_cast<unsigned short*>(s_tw + N);   // of shared memory is too small to hold them
    const int msg = blockIdx.x, t = threadIdx.x;
    fill_twiddles_deint<GEO>(s_tw, tb.tw2d, t, L2_THREADS);
    for (int i = t; i < LWE2_STRIDE_IN; i += L2_THREADS) la[i] = (unsigned short)(lwe[(size_t)msg * LWE2_STRIDE_IN + i] & (LWE2_Q - 1));
    __syncthreads();
    init_acc<F, GEO>(acc, tb.lut2, la[LWE2_N], t);
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < LWE2_N; ++i) {
        const int a = la[i];
        if (a == 0) continue;                                  // CTA-uniform: (X^0 - 1) acc = 0 adds nothing
        const double* key = bsk2 + (size_t)i * 2 * L * 2 * N + out_idx<GEO>(t, 0);
        double ma[E], mb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = 0.0; mb[k] = 0.0; }
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            i64 u[E];
            decompose_words<F, G, GEO>(u, acc + p * N, a, t);
#pragma unroll 1
            for (int r = 0; r < L; r += 2) {
                double x[E], y[E];
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    x[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r));
                    y[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r + 1));
                }
                ntt_forward2s<AR, GEO, LdSharedD>(x, y, bx, by, s_tw, t, 0);
                const double* kx = key + (size_t)(p * L + r) * 2 * N;         // rows r and r+1: [a | b] each
#pragma unroll
                for (int k = 0; k < E; k += 2) {
                    const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
                    const double2 xa = ld_stream_f64x2(kx + o), xb = ld_stream_f64x2(kx + N + o);
                    const double2 ya = ld_stream_f64x2(kx + 2 * N + o), yb = ld_stream_f64x2(kx + 3 * N + o);
                    // 12 terms x 0.66q < 2^53: the running sums stay exact
                    ma[k] = __dadd_rn(ma[k], __dadd_rn(D2::mulmod_key(x[k], xa.x), D2::mulmod_key(y[k], ya.x)));
                    ma[k + 1] = __dadd_rn(ma[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xa.y), D2::mulmod_key(y[k + 1], ya.y)));
                    mb[k] = __dadd_rn(mb[k], __dadd_rn(D2::mulmod_key(x[k], xb.x), D2::mulmod_key(y[k], yb.x)));
                    mb[k + 1] = __dadd_rn(mb[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xb.y), D2::mulmod_key(y[k + 1], yb.y)));
                }
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = D2::renorm(ma[k]); mb[k] = D2::renorm(mb[k]); }
        ntt_inverse2s<AR, GEO, LdGlobal>(ma, mb, bx, by, tb.itw2d, t, 0);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = t + GEO::NT * k;
            i64 va = (i64)acc[pos] + D2::to_i64(ma[k]), vb = (i64)acc[N + pos] + D2::to_i64(mb[k]);
            va += va < 0 ? (i64)F::Q : 0; va -= va >= (i64)F::Q ? (i64)F::Q : 0;
            vb += vb < 0 ? (i64)F::Q : 0; vb -= vb >= (i64)F::Q ? (i64)F::Q : 0;
            acc[pos] = (u64)va; acc[N + pos] = (u64)vb;
        }
        __syncthreads();
    }
    for (int e = t; e < 2 * N; e += L2_THREADS) out[(size_t)msg * 2 * N + e] = acc[e];
}

// K3, latency shape (batches of at most one message per SM): 512 threads per message.  Half h (256 threads) decomposes
// polynomial h of acc and runs its 6 digit transforms and their MACs; the halves swap one partial sum each through shared
// memory, so half 0 finishes the a-polynomial and half 1 the b-polynomial (one inverse transform each).  All sums are
// exact integers below 2^53 in both shapes, so the result is bit-identical to l2_blind_rotate_kernel.
constexpr int L2L_THREADS = 2 * GeoL2::NT;
constexpr size_t L2L_SMEM = (size_t)2 * F2::N * 8 + (size_t)4 * GeoL2::BUF * 8 + (size_t)2 * F2::N * 8 + F2::N * sizeof(double2) + 672 * sizeof(unsigned short);

__global__ void __launch_bounds__(L2L_THREADS, 1)
l2_blind_rotate_lat_kernel(const u32* __restrict__ lwe, const double* __restrict__ bsk2, u64* __restrict__ out, Tables tb) {
    typedef F2 F; typedef G2 G; typedef GeoL2 GEO; typedef ArD2 AR;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    double* bufs = reinterpret_cast<double*>(acc + 2 * N);
    double* swp = bufs + 4 * GEO::BUF;                                   // [2][N] partial sums handed to the other half
    double2* s_tw = reinterpret_cast<double2*>(swp + 2 * N);
    unsigned short* la = reinterpret_cast<unsigned short*>(s_tw + N);
    const int msg = blockIdx.x, h = threadIdx.x / GEO::NT, t = threadIdx.x % GEO::NT, bar = 1 + h;
    double* bx = bufs + (size_t)h * 2 * GEO::BUF; double* by = bx + GEO::BUF;
    fill_twiddles_deint<GEO>(s_tw, tb.tw2d, threadIdx.x, L2L_THREADS);
    for (int i = threadIdx.x; i < LWE2_STRIDE_IN; i += L2L_THREADS) la[i] = (unsigned short)(lwe[(size_t)msg * LWE2_STRIDE_IN + i] & (LWE2_Q - 1));
    __syncthreads();
    if (h == 0) init_acc<F, GEO>(acc, tb.lut2, la[LWE2_N], t);
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < LWE2_N; ++i) {
        const int a = la[i];
        if (a == 0) continue;
        const double* key = bsk2 + ((size_t)i * 2 * L + (size_t)h * L) * 2 * N + out_idx<GEO>(t, 0);
        double ma[E], mb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = 0.0; mb[k] = 0.0; }
        i64 u[E];
        decompose_words<F, G, GEO>(u, acc + h * N, a, t);
#pragma unroll 1
        for (int r = 0; r < L; r += 2) {
            double x[E], y[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                x[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r));
                y[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r + 1));
            }
            ntt_forward2s<AR, GEO, LdSharedD>(x, y, bx, by, s_tw, t, bar);
            const double* kx = key + (size_t)r * 2 * N;
#pragma unroll
            for (int k = 0; k < E; k += 2) {
                const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
                const double2 xa = ld_stream_f64x2(kx + o), xb = ld_stream_f64x2(kx + N + o);
                const double2 ya = ld_stream_f64x2(kx + 2 * N + o), yb = ld_stream_f64x2(kx + 3 * N + o);
                ma[k] = __dadd_rn(ma[k], __dadd_rn(D2::mulmod_key(x[k], xa.x), D2::mulmod_key(y[k], ya.x)));
                ma[k + 1] = __dadd_rn(ma[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xa.y), D2::mulmod_key(y[k + 1], ya.y)));
                mb[k] = __dadd_rn(mb[k], __dadd_rn(D2::mulmod_key(x[k], xb.x), D2::mulmod_key(y[k], yb.x)));
                mb[k + 1] = __dadd_rn(mb[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xb.y), D2::mulmod_key(y[k + 1], yb.y)));
            }
        }
        // half 0 keeps the a-sums and hands over its b-sums, half 1 the other way round
        double m[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { m[k] = h ? mb[k] : ma[k]; swp[(size_t)h * N + t + GEO::NT * k] = h ? ma[k] : mb[k]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < E; ++k) m[k] = D2::renorm(__dadd_rn(m[k], swp[(size_t)(1 - h) * N + t + GEO::NT * k]));
        ExBuf<double> eb{bx, by};
        ntt_inverse<AR, GEO, LdGlobal>(m, eb, tb.itw2d, t, bar);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = h * N + t + GEO::NT * k;
            i64 v = (i64)acc[pos] + D2::to_i64(m[k]);
            v += v < 0 ? (i64)F::Q : 0; v -= v >= (i64)F::Q ? (i64)F::Q : 0;
            acc[pos] = (u64)v;
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < 2 * N; e += L2L_THREADS) out[(size_t)msg * 2 * N + e] = acc[e];
}

// K3, cluster shape (at most n_sm / 6 messages): one thread-block cluster of 6 CTAs (6 SMs) per message.  CTA c holds
// polynomial c / 3 of the accumulator, transforms digits 2(c % 3) and 2(c % 3) + 1 of it and multiplies them with their
// key rows; the 6 partial sums are exchanged through a double-buffered global scratch (L2-resident) around ONE cluster
// barrier per CMux step, and the three CTAs of a polynomial each run its inverse transform (redundantly, so no broadcast
// is needed).  Critical path per step: one transform pair and one inverse instead of six pairs and one inverse pair.
constexpr int L2C_CLUSTER = 6;
constexpr size_t L2C_SMEM = (size_t)F2::N * 8 + (size_t)2 * GeoL2::BUF * 8 + F2::N * sizeof(double2) + 672 * sizeof(unsigned short);
constexpr size_t L2C_SCRATCH_WORDS = (size_t)2 * L2C_CLUSTER * 2 * F2::N;          // doubles per message

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(L2C_CLUSTER, 1, 1) __launch_bounds__(GeoL2::NT, 1)
l2_blind_rotate_cluster_kernel(const u32* __restrict__ lwe, const double* __restrict__ bsk2, u64* __restrict__ out,
                               double* __restrict__ scratch, Tables tb) {
    typedef F2 F; typedef G2 G; typedef GeoL2 GEO; typedef ArD2 AR;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);                         // polynomial p only
    double* bx = reinterpret_cast<double*>(acc + N); double* by = bx + GEO::BUF;
    double2* s_tw = reinterpret_cast<double2*>(by + GEO::BUF);
    unsigned short* la = reinterpret_cast<unsigned short*>(s_tw + N);
    const int msg = blockIdx.x / L2C_CLUSTER, c = blockIdx.x % L2C_CLUSTER, t = threadIdx.x;
    const int p = c / 3, r0 = 2 * (c % 3);
    double* scr = scratch + (size_t)msg * L2C_SCRATCH_WORDS;
    fill_twiddles_deint<GEO>(s_tw, tb.tw2d, t, GEO::NT);
    for (int i = t; i < LWE2_STRIDE_IN; i += GEO::NT) la[i] = (unsigned short)(lwe[(size_t)msg * LWE2_STRIDE_IN + i] & (LWE2_Q - 1));
    __syncthreads();
    {
        const int rot = (2 * N - la[LWE2_N]) % (2 * N);                  // acc = (0, LUT * X^(2N - b))
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = t + GEO::NT * k;
            const i64 v = rotated<F>(tb.lut2, pos, rot);
            acc[pos] = p ? (u64)(v < 0 ? v + (i64)F::Q : v) : 0;
        }
    }
    __syncthreads();
    int par = 0;
#pragma unroll 1
    for (int i = 0; i < LWE2_N; ++i) {
        const int a = la[i];
        if (a == 0) continue;                                  // uniform over the cluster
        const double* kx = bsk2 + ((size_t)i * 2 * L + (size_t)p * L + r0) * 2 * N + out_idx<GEO>(t, 0);
        i64 u[E];
        decompose_words<F, G, GEO>(u, acc, a, t);
        double x[E], y[E], ma[E], mb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            x[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r0));
            y[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r0 + 1));
        }
        ntt_forward2s<AR, GEO, LdSharedD>(x, y, bx, by, s_tw, t, 0);
#pragma unroll
        for (int k = 0; k < E; k += 2) {
            const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
            const double2 xa = ld_stream_f64x2(kx + o), xb = ld_stream_f64x2(kx + N + o);
            const double2 ya = ld_stream_f64x2(kx + 2 * N + o), yb = ld_stream_f64x2(kx + 3 * N + o);
            ma[k] = __dadd_rn(D2::mulmod_key(x[k], xa.x), D2::mulmod_key(y[k], ya.x));
            ma[k + 1] = __dadd_rn(D2::mulmod_key(x[k + 1], xa.y), D2::mulmod_key(y[k + 1], ya.y));
            mb[k] = __dadd_rn(D2::mulmod_key(x[k], xb.x), D2::mulmod_key(y[k], yb.x));
            mb[k + 1] = __dadd_rn(D2::mulmod_key(x[k + 1], xb.y), D2::mulmod_key(y[k + 1], yb.y));
        }
        double* w = scr + ((size_t)par * L2C_CLUSTER + c) * 2 * N;
#pragma unroll
        for (int k = 0; k < E; ++k) { __stcg(w + t + GEO::NT * k, ma[k]); __stcg(w + N + t + GEO::NT * k, mb[k]); }
        cluster_barrier();
        double m[E];
#pragma unroll
        for (int k = 0; k < E; ++k) m[k] = p ? mb[k] : ma[k];
#pragma unroll
        for (int cc = 1; cc < L2C_CLUSTER; ++cc) {
            const int src = (c + cc) % L2C_CLUSTER;
            const double* rd = scr + ((size_t)par * L2C_CLUSTER + src) * 2 * N + (size_t)p * N;
#pragma unroll
            for (int k = 0; k < E; ++k) m[k] = __dadd_rn(m[k], __ldcg(rd + t + GEO::NT * k));   // 12 terms x 0.66q < 2^53: exact
        }
#pragma unroll
        for (int k = 0; k < E; ++k) m[k] = D2::renorm(m[k]);
        ExBuf<double> eb{bx, by};
        ntt_inverse<AR, GEO, LdGlobal>(m, eb, tb.itw2d, t, 0);
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = t + GEO::NT * k;
            i64 v = (i64)acc[pos] + D2::to_i64(m[k]);
            v += v < 0 ? (i64)F::Q : 0; v -= v >= (i64)F::Q ? (i64)F::Q : 0;
            acc[pos] = (u64)v;
        }
        __syncthreads();
        par ^= 1;
    }
    if (c % 3 == 0)
        for (int e = t; e < N; e += GEO::NT) out[(size_t)msg * 2 * N + (size_t)p * N + e] = acc[e];
}

// ---- K2: sample extraction + LWE key switch + modulus switch + offset ---------------------------------------------
// out[col] = (0,..,0,b0) - SUM_{i<1024, j<27} d_ij * KSK[i][j][col]  (d = balanced base-2 digits of a'_i), then
// x -> round(x * 4096 / q1) mod 4096, b += 7 * 128.   A {-1,0,1} x u32 integer product: thread = column,
// KS_MB messages per CTA share each key row load.
constexpr int KS_MB = 16, KS_THREADS = 128;
constexpr size_t KS_SMEM = (size_t)KS_MB * F1::N * sizeof(i32);

// Final step of K2 for one (message, column): reduce, subtract from (0,..,0,b0), modulus switch, offset.
__device__ __forceinline__ u32 ks_finish(i64 sum, u32 b0, int col) {
    i64 s = sum % (i64)Q1; if (s < 0) s += Q1;
    const u32 base = col == LWE2_N ? b0 : 0u;
    const u32 x = base >= (u32)s ? base - (u32)s : base + Q1 - (u32)s;
    u32 y = (u32)(((u64)2 * LWE2_Q * x + Q1) / (2ull * Q1)) & (LWE2_Q - 1);
    if (col == LWE2_N) y = (y + CLUE_COUNT * (LWE2_Q >> 5)) & (LWE2_Q - 1);
    return y;
}

// SPLIT = false: one CTA walks all 1024 key rows (throughput shape).  SPLIT = true (small batches): gridDim.z CTAs each
// walk 1024 / gridDim.z rows and add their exact integer partial sums into `part` [B][KSK_PAD] (zeroed by the caller);
// keyswitch_finish_kernel completes the step.  Integer addition commutes, so both shapes give identical words.
template <bool SPLIT> __global__ void __launch_bounds__(KS_THREADS)
keyswitch_kernel(const u32* __restrict__ rlwe, const u32* __restrict__ ksk, u32* __restrict__ out, int B,
                 unsigned long long* __restrict__ part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    i32 (*uw)[F1::N] = reinterpret_cast<i32 (*)[F1::N]>(smem_raw);        // [KS_MB][N] offset words
    const int m0 = blockIdx.x * KS_MB, col = blockIdx.y * KS_THREADS + threadIdx.x;
    for (int e = threadIdx.x; e < KS_MB * F1::N; e += KS_THREADS) {
        const int m = e / F1::N, i = e % F1::N;
        i32 w = 0;
        if (m0 + m < B) {
            const u32* a = rlwe + (size_t)(m0 + m) * 2 * F1::N;
            // a' = (a0, -a_{N-1}, ..., -a_1): constant-term sample extraction (extract_lwe_locally, detector.rs:561)
            u32 ai = i == 0 ? a[0] : (a[F1::N - i] ? Q1 - a[F1::N - i] : 0);
            i32 v = ai > (Q1 >> 1) ? (i32)ai - (i32)Q1 : (i32)ai;
            w = v + ((1 << 26) - 1);                    // offset word, base 2, 27 levels, no drop
        }
        uw[m][i] = w;
    }
    __syncthreads();
    if (col >= KSK_PAD) return;
    i64 acc[KS_MB];
#pragma unroll
    for (int m = 0; m < KS_MB; ++m) acc[m] = 0;
    const int rows = SPLIT ? F1::N / (int)gridDim.z : F1::N, i_lo = SPLIT ? (int)blockIdx.z * rows : 0;
#pragma unroll 1
    for (int i = i_lo; i < i_lo + rows; ++i) {
        i32 w[KS_MB];
#pragma unroll
        for (int m = 0; m < KS_MB; ++m) w[m] = uw[m][i];
        const u32* row = ksk + (size_t)i * KS_LEVELS * KSK_PAD + col;
        // 32-bit partial sums over halves of the 27 levels (14 x 2^27 < 2^31), widened once per half: IMAD instead of
        // IMAD.WIDE (64 vs 25 lanes/clk/SM) in the innermost loop
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            i32 part[KS_MB];
#pragma unroll
            for (int m = 0; m < KS_MB; ++m) part[m] = 0;
            const int j0 = half ? 14 : 0, j1 = half ? KS_LEVELS : 14;
#pragma unroll
            for (int j = j0; j < j1; ++j) {
                const i32 k = (i32)__ldg(row + (size_t)j * KSK_PAD);
#pragma unroll
                for (int m = 0; m < KS_MB; ++m) {
                    const i32 d = j < KS_LEVELS - 1 ? ((w[m] >> j) & 1) - 1 : (w[m] >> (KS_LEVELS - 1));
                    part[m] += d * k;
                }
            }
#pragma unroll
            for (int m = 0; m < KS_MB; ++m) acc[m] += (i64)part[m];
        }
    }
    if (col > LWE2_N) return;
#pragma unroll
    for (int m = 0; m < KS_MB; ++m) {
        if (m0 + m >= B) break;
        if (SPLIT) atomicAdd(part + (size_t)(m0 + m) * KSK_PAD + col, (unsigned long long)acc[m]);
        else out[(size_t)(m0 + m) * LWE2_STRIDE_IN + col] = ks_finish(acc[m], rlwe[(size_t)(m0 + m) * 2 * F1::N + F1::N], col);
    }
}

__global__ void keyswitch_finish_kernel(const u32* __restrict__ rlwe, const unsigned long long* __restrict__ part,
                                        u32* __restrict__ out, int B) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, m = e / KSK_PAD, col = e % KSK_PAD;
    if (m >= B || col > LWE2_N) return;
    out[(size_t)m * LWE2_STRIDE_IN + col] = ks_finish((i64)part[e], rlwe[(size_t)m * 2 * F1::N + F1::N], col);
}

// offset word of the extracted mask coefficient a'_i: every digit of the base-2 decomposition is a bit of it (minus one below the top)
__device__ __forceinline__ i32 ks_offset_word(const u32* __restrict__ a, int i) {
    const u32 ai = i == 0 ? a[0] : (a[F1::N - i] ? Q1 - a[F1::N - i] : 0);      // a' = (a0, -a_{N-1}, ..., -a_1)
    const i32 v = ai > (Q1 >> 1) ? (i32)ai - (i32)Q1 : (i32)ai;
    return v + ((1 << 26) - 1);                                                  // offset word, base 2, 27 levels
}

// K2, throughput shape: the same {0,1}-digit x key product on the CUDA cores with IDP.4A (4 multiply-adds per lane-op, full rate on
// B200: profiles/r2_pipe_microbench.txt).  Balanced base-2 digits of a 27-bit offset word w are d_j = bit_j(w) - 1 (j < 26) and
// d_26 = bit_26(w), so  SUM_j d_j K_j = SUM_j bit_j(w) K_j - SUM_{j<26} K_j : the second term is a per-column constant of the key
// (ksd_colsum), the first is a dot product of BITS with the key.  The key is stored as four balanced base-256 limbs, byte-packed over
// groups of four levels (levels padded 27 -> 28): ksd[(i*7 + g)][col] = uint4 {limb0..limb3}, each u32 = the limb bytes of levels
// 4g..4g+3 — one coalesced 16-byte load per (coefficient, group, column).  A CTA = 128 columns x 16 messages; the digit bytes of its
// messages are expanded once per chunk of 64 coefficients into shared memory ([i][g][message] words, read as broadcast 16-byte
// vectors), so the inner loop is 1 LDG.128 + 4 LDS.128 + 64 IDP.4A per (coefficient, group).  int32 limb sums cannot overflow
// (28 672 x 128 < 2^31); they recombine to the exact integer keyswitch_kernel accumulates.
constexpr int KSD_G = 7, KSD_MB = 16, KSD_THREADS = 128, KSD_IC = 64;
constexpr size_t KSD_WORDS = (size_t)F1::N * KSD_G * KSK_PAD * 4;            // u32 words of the packed key (77 MB)
constexpr size_t KSD_SMEM = (size_t)KSD_IC * KSD_G * KSD_MB * 4;              // 28 KiB of digit words

__global__ void ksd_build_kernel(const u32* __restrict__ ksk /*[1024*27][KSK_PAD]*/, uint4* __restrict__ ksd) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)F1::N * KSD_G * KSK_PAD) return;
    const int col = (int)(e % KSK_PAD); const size_t ig = e / KSK_PAD; const int g = (int)(ig % KSD_G), i = (int)(ig / KSD_G);
    u32 limb[4] = {0, 0, 0, 0};
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
        const int j = 4 * g + jj;
        u32 w = (j < KS_LEVELS && col <= LWE2_N) ? ksk[((size_t)i * KS_LEVELS + j) * KSK_PAD + col] : 0u;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            i32 t = (i32)(w & 255u); w >>= 8;
            if (t >= 128) { t -= 256; w += 1; }
            limb[l] |= (u32)(t & 0xFF) << (8 * jj);
        }
    }
    ksd[e] = make_uint4(limb[0], limb[1], limb[2], limb[3]);
}
// colsum[col] = SUM_i SUM_{j<26} KSK[i][j][col]
__global__ void ksd_colsum_kernel(const u32* __restrict__ ksk, long long* __restrict__ colsum) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= KSK_PAD) return;
    long long s = 0;
    for (int i = 0; i < F1::N; ++i)
        for (int j = 0; j < KS_LEVELS - 1; ++j) s += (long long)ksk[((size_t)i * KS_LEVELS + j) * KSK_PAD + col];
    colsum[col] = s;
}
__global__ void __launch_bounds__(KSD_THREADS)
keyswitch_dp4a_kernel(const u32* __restrict__ rlwe, const uint4* __restrict__ ksd, const long long* __restrict__ colsum, u32* __restrict__ out, int B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* dw = reinterpret_cast<u32*>(smem_raw);                               // [KSD_IC][KSD_G][KSD_MB]
    const int m0 = blockIdx.x * KSD_MB, col = blockIdx.y * KSD_THREADS + threadIdx.x;
    const bool live = col < KSK_PAD;
    i32 acc[KSD_MB][4];
#pragma unroll
    for (int m = 0; m < KSD_MB; ++m) { acc[m][0] = 0; acc[m][1] = 0; acc[m][2] = 0; acc[m][3] = 0; }
#pragma unroll 1
    for (int i0 = 0; i0 < F1::N; i0 += KSD_IC) {
        __syncthreads();                                                      // the previous chunk's digits have been consumed
        for (int e = threadIdx.x; e < KSD_IC * KSD_MB; e += KSD_THREADS) {
            const int m = e % KSD_MB, il = e / KSD_MB;
            u32 w = 0;                                                        // a message beyond the batch contributes nothing
            if (m0 + m < B) w = (u32)ks_offset_word(rlwe + (size_t)(m0 + m) * 2 * F1::N, i0 + il);
#pragma unroll
            for (int g = 0; g < KSD_G; ++g) dw[(il * KSD_G + g) * KSD_MB + m] = (((w >> (4 * g)) & 0xFu) * 0x00204081u) & 0x01010101u;
        }
        __syncthreads();
        if (live) {
            const uint4* kp = ksd + (size_t)i0 * KSD_G * KSK_PAD + col;
#pragma unroll 1
            for (int il = 0; il < KSD_IC; ++il) {
#pragma unroll
                for (int g = 0; g < KSD_G; ++g) {
                    const uint4 kk = __ldg(kp + (size_t)(il * KSD_G + g) * KSK_PAD);
                    const uint4* d4 = reinterpret_cast<const uint4*>(dw + (il * KSD_G + g) * KSD_MB);
#pragma unroll
                    for (int q = 0; q < KSD_MB / 4; ++q) {
                        const uint4 dv = d4[q];
                        const u32 dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            acc[4 * q + r][0] = __dp4a((int)dd[r], (int)kk.x, acc[4 * q + r][0]);
                            acc[4 * q + r][1] = __dp4a((int)dd[r], (int)kk.y, acc[4 * q + r][1]);
                            acc[4 * q + r][2] = __dp4a((int)dd[r], (int)kk.z, acc[4 * q + r][2]);
                            acc[4 * q + r][3] = __dp4a((int)dd[r], (int)kk.w, acc[4 * q + r][3]);
                        }
                    }
                }
            }
        }
    }
    if (!live || col > LWE2_N) return;
    const i64 cs = colsum[col];
#pragma unroll
    for (int m = 0; m < KSD_MB; ++m) {
        if (m0 + m >= B) break;
        const i64 sum = (i64)acc[m][0] + (i64)acc[m][1] * 256 + (i64)acc[m][2] * 65536 + (i64)acc[m][3] * 16777216 - cs;
        out[(size_t)(m0 + m) * LWE2_STRIDE_IN + col] = ks_finish(sum, rlwe[(size_t)(m0 + m) * 2 * F1::N + F1::N], col);
    }
}

// K2 as a tensor-core GEMM (ks_gemm.cu): operand expansion and the epilogue.  K index = i * 27 + j, N index = col * 4 + limb.
constexpr int KSG_K = F1::N * KS_LEVELS;                     // 27 648
constexpr size_t KSG_MIN_B = 1, KSG_CHUNK = 2048;            // faster than the CUDA-core kernels at every batch size; 2 048 messages per GEMM (79 MB of scratch)
constexpr int KSG_LIMBS = 4, KSG_N = ((LWE2_N + 1) * KSG_LIMBS + 15) / 16 * 16;   // 2 688
// A[m][i*27 + j] = balanced base-2 digit j of the extracted mask coefficient a'_i (the same digits keyswitch_kernel uses).
// One thread writes 16 consecutive bytes of a row (coalesced 16-byte stores); they span at most two coefficients.
__global__ void ks_digits_kernel(const u32* __restrict__ rlwe, signed char* __restrict__ A, int B) {
    constexpr int CH = KSG_K / 16;                                               // 1 728 chunks of 16 bytes per message
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * CH) return;
    const int m = (int)(e / CH), k0 = (int)(e % CH) * 16;
    const u32* a = rlwe + (size_t)m * 2 * F1::N;
    int i = k0 / KS_LEVELS, j = k0 % KS_LEVELS;
    i32 w = ks_offset_word(a, i);
    u32 pack[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        const i32 d = j < KS_LEVELS - 1 ? ((w >> j) & 1) - 1 : (w >> (KS_LEVELS - 1));
        pack[b >> 2] |= (u32)(d & 0xFF) << (8 * (b & 3));
        if (++j == KS_LEVELS) { j = 0; ++i; if (i < F1::N) w = ks_offset_word(a, i); }
    }
    *reinterpret_cast<uint4*>(A + (size_t)m * KSG_K + k0) = make_uint4(pack[0], pack[1], pack[2], pack[3]);
}
// Bt[n = col*4 + limb][k = i*27 + j] = balanced base-256 limb of KSK[i][j][col] (device key layout: rows padded to KSK_PAD)
__global__ void ks_limbs_kernel(const u32* __restrict__ ksk, signed char* __restrict__ Bt) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)KSG_K * (KSG_N / KSG_LIMBS)) return;
    const int k = (int)(e % KSG_K), col = (int)(e / KSG_K);
    u32 w = col <= LWE2_N ? ksk[(size_t)k * KSK_PAD + col] : 0u;
#pragma unroll
    for (int l = 0; l < KSG_LIMBS; ++l) {
        i32 t = (i32)(w & 255u); w >>= 8;
        if (t >= 128) { t -= 256; w += 1; }
        Bt[((size_t)col * KSG_LIMBS + l) * KSG_K + k] = (signed char)t;
    }
}
// sum = SUM_l C[m][col*4 + l] * 256^l, then the same final step as keyswitch_kernel
__global__ void ks_combine_kernel(const u32* __restrict__ rlwe, const i32* __restrict__ C, u32* __restrict__ out, int B) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)B * (LWE2_N + 1)) return;
    const int m = (int)(e / (LWE2_N + 1)), col = (int)(e % (LWE2_N + 1));
    const int4 c = *reinterpret_cast<const int4*>(C + (size_t)m * KSG_N + (size_t)col * KSG_LIMBS);
    const i64 sum = (i64)c.x + (i64)c.y * 256 + (i64)c.z * 65536 + (i64)c.w * 16777216;
    out[(size_t)m * LWE2_STRIDE_IN + col] = ks_finish(sum, rlwe[(size_t)m * 2 * F1::N + F1::N], col);
}

// ---- K4: scale by N^-1, homomorphic trace, forward NTT ----------------------------------------------------------------
// The 11 x 25 digit transforms, the MAC against the trace key and the inverse transforms run on the FP64 pipe like K3
// (two digits per pass, key words = centred doubles x N^-1); the two final forward transforms (to_ntt_rlwe) stay integer.
constexpr int TR_THREADS = GeoL2::NT;
constexpr size_t TR_SMEM = (size_t)2 * F2::N * 8 + (size_t)2 * GeoL2::BUF * 8 + F2::N * sizeof(double2);   // acc, two exchange buffers, forward twiddles

// sigma_d(p)[pos] as a signed value: source index i0 = pos * d^-1 mod 2N
__device__ __forceinline__ i64 automorphed(const u64* p, int pos, u32 dinv) {
    const u32 i0 = ((u32)pos * dinv) & (2 * F2::N - 1);
    return i0 < (u32)F2::N ? (i64)p[i0] : -(i64)p[i0 - F2::N];
}

__global__ void __launch_bounds__(TR_THREADS, 2)
trace_kernel(u64* __restrict__ ct, const double* __restrict__ trk, Tables tb) {
    typedef F2 F; typedef GeoL2 GEO; typedef ArD2 AR;
    constexpr int N = F::N, E = GEO::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);        // [2][N]: a, b (canonical)
    double* bx = reinterpret_cast<double*>(acc + 2 * N); double* by = bx + GEO::BUF;
    double2* s_tw = reinterpret_cast<double2*>(by + GEO::BUF);          // de-interleaved forward twiddles, as in K3
    const int t = threadIdx.x;
    fill_twiddles_deint<GEO>(s_tw, tb.tw2d, t, TR_THREADS);
    u64* g = ct + (size_t)blockIdx.x * 2 * N;
    for (int e = t; e < 2 * N; e += TR_THREADS) acc[e] = F::csub(F::mul_shoup(g[e], tb.n2_inv), F::Q);   // detector.rs:635-636
    __syncthreads();
#pragma unroll 1
    for (int step = 0; step < TR_STEPS; ++step) {
        const u32 dinv = tb.trace_dinv[step];
        i64 u[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            i64 v = automorphed(acc, t + GEO::NT * k, dinv);              // in (-q, q)
            constexpr i64 H = (i64)(F::Q >> 1);
            if (v > H) v -= (i64)F::Q;
            if (v < -H) v += (i64)F::Q;
            u[k] = gadget_word<F, GT>(v);
        }
        double ma[E], mb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = 0.0; mb[k] = 0.0; }
        const double* key = trk + (size_t)step * TR_LEVELS * 2 * N + out_idx<GEO>(t, 0);
#pragma unroll 1
        for (int r = 0; r < TR_LEVELS; r += 2) {
            const bool pair = r + 1 < TR_LEVELS;                          // 25 digits: 12 pairs and a single
            double x[E], y[E];
#pragma unroll
            for (int k = 0; k < E; ++k) {
                x[k] = D2::from_small(gadget_digit_signed<F, GT>(u[k], r));
                y[k] = pair ? D2::from_small(gadget_digit_signed<F, GT>(u[k], r + 1)) : 0.0;
            }
            ntt_forward2s<AR, GEO, LdSharedD>(x, y, bx, by, s_tw, t, 0);
            const double* kx = key + (size_t)r * 2 * N;
            const double* ky = kx + (pair ? 2 * N : 0);                   // (single: y = NTT(0) = 0, any key row)
#pragma unroll
            for (int k = 0; k < E; k += 2) {
                const int o = out_idx<GEO>(0, k) - out_idx<GEO>(0, 0);
                const double2 xa = ld_stream_f64x2(kx + o), xb = ld_stream_f64x2(kx + N + o);
                const double2 ya = ld_stream_f64x2(ky + o), yb = ld_stream_f64x2(ky + N + o);
                ma[k] = __dadd_rn(ma[k], __dadd_rn(D2::mulmod_key(x[k], xa.x), D2::mulmod_key(y[k], ya.x)));
                ma[k + 1] = __dadd_rn(ma[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xa.y), D2::mulmod_key(y[k + 1], ya.y)));
                mb[k] = __dadd_rn(mb[k], __dadd_rn(D2::mulmod_key(x[k], xb.x), D2::mulmod_key(y[k], yb.x)));
                mb[k + 1] = __dadd_rn(mb[k + 1], __dadd_rn(D2::mulmod_key(x[k + 1], xb.y), D2::mulmod_key(y[k + 1], yb.y)));
            }
            if (r == 8 || r == 18) {                                      // every 10 terms: 0.5q + 10 x 0.66q < 2^53
#pragma unroll
                for (int k = 0; k < E; ++k) { ma[k] = D2::renorm(ma[k]); mb[k] = D2::renorm(mb[k]); }
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k) { ma[k] = D2::renorm(ma[k]); mb[k] = D2::renorm(mb[k]); }
        ntt_inverse2s<AR, GEO, LdGlobal>(ma, mb, bx, by, tb.itw2d, t, 0);
        // b' = b + ks.b + sigma_d(b): read the permuted b before anyone overwrites it
        u64 sb[E];
#pragma unroll
        for (int k = 0; k < E; ++k) { i64 v = automorphed(acc + N, t + GEO::NT * k, dinv); sb[k] = (u64)(v < 0 ? v + (i64)F::Q : v); }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int pos = t + GEO::NT * k;
            i64 va = (i64)acc[pos] + D2::to_i64(ma[k]), vb = (i64)F::csub(acc[N + pos] + sb[k], F::Q) + D2::to_i64(mb[k]);
            va += va < 0 ? (i64)F::Q : 0; va -= va >= (i64)F::Q ? (i64)F::Q : 0;
            vb += vb < 0 ? (i64)F::Q : 0; vb -= vb >= (i64)F::Q ? (i64)F::Q : 0;
            acc[pos] = (u64)va; acc[N + pos] = (u64)vb;
        }
        __syncthreads();
    }
    // to_ntt_rlwe (detector.rs:638): forward NTT of a and b, canonical output (integer arithmetic, same two buffers)
    u64 x[E], y[E];
#pragma unroll
    for (int k = 0; k < E; ++k) { x[k] = acc[t + GEO::NT * k]; y[k] = acc[N + t + GEO::NT * k]; }
    ntt_forward2s<ArInt<F2>, GEO, LdGlobal>(x, y, reinterpret_cast<u64*>(bx), reinterpret_cast<u64*>(by), tb.tw2, t, 0);
#pragma unroll
    for (int k = 0; k < E; ++k) { g[out_idx<GEO>(t, k)] = F::canon_lazy(x[k]); g[N + out_idx<GEO>(t, k)] = F::canon_lazy(y[k]); }
}

// ---- K5 / K6: digest packing ------------------------------------------------------------------------------------------
// partial[cipher][chunk] = SUM_{m in chunk} PV_m (.) NTT(plaintext_{m,cipher}); plaintexts are generated directly in
// the registers of the first NTT pass.  grid = (n_cipher, n_chunks): ciphers of one chunk are adjacent so the chunk's
// pertinency ciphertexts are shared through L2.
struct PackIndexArgs {            // detector.rs:223-339 + RetrievalParams
    u32 slots_per_bucket, slots_per_segment, segment_per_cipher, bucket_count; u64 seed; u32 cipher_idx0;
};
struct PackPayloadArgs {          // detector.rs:341-453
    const unsigned short* payloads; const unsigned short* weights; size_t weight_stride; u32 cmb_per_cipher;
};
constexpr int PACK_THREADS = GeoL2::NT, PACK_CHUNK = 128;
constexpr size_t PACK_SMEM = (size_t)2 * GeoL2::BUF * 8;

__device__ __forceinline__ u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
// counter-based bucket choice (replaces thread_rng at detector.rs:262; identical in the oracle)
__device__ __forceinline__ u32 bucket_of(u64 seed, u32 cipher_idx, u64 msg, u32 seg, u32 buckets) {
    u64 h = mix64(seed + 0x9E3779B97F4A7C15ull * (msg + 1));
    h = mix64(h ^ ((((u64)cipher_idx << 32) | seg) * 0xD1342543DE82EF95ull));
    return (u32)(((h >> 32) * buckets) >> 32);
}
__device__ __forceinline__ u64 centred_p(u32 v) { return v < ((OUT_P + 1) >> 1) ? (u64)v : Q2 - OUT_P + v; }

template <bool INDICES>
__global__ void __launch_bounds__(PACK_THREADS, 2)
pack_kernel(const u64* __restrict__ pv, size_t count, u64 index0, PackIndexArgs ia, PackPayloadArgs pa,
            u64* __restrict__ partial /*[n_cipher][n_chunks][2][N]*/, Tables tb) {
    typedef F2 F; typedef F::Acc Acc; typedef GeoL2 GEO; typedef ArInt<F2> AR;
    constexpr int N = F::N, E = GEO::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* bufs = reinterpret_cast<u64*>(smem_raw);
    ExBuf<u64> eb{bufs, bufs + GEO::BUF};
    const int t = threadIdx.x, cipher = blockIdx.x, chunk = blockIdx.y;
    const size_t m_begin = (size_t)chunk * PACK_CHUNK, m_end = min(count, m_begin + PACK_CHUNK);
    Acc ma[E], mb[E];
#pragma unroll
    for (int k = 0; k < E; ++k) { ma[k] = Acc(); mb[k] = Acc(); }
#pragma unroll 1
    for (size_t m = m_begin; m < m_end; ++m) {
        const u64 gi = index0 + m;
        u64 x[E];
        if (INDICES) {
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const u32 pos = t + GEO::NT * k;
                const u32 seg = pos / ia.slots_per_segment, off = pos % ia.slots_per_segment;
                u64 v = 0;
                if (seg < ia.segment_per_cipher) {
                    const u32 bucket = bucket_of(ia.seed, ia.cipher_idx0 + cipher, gi, seg, ia.bucket_count);
                    if (off / ia.slots_per_bucket == bucket) {
                        const u32 slot = off % ia.slots_per_bucket;
                        if (slot == ia.slots_per_bucket - 1) v = 1;
                        else {                                      // digit `slot` of gi in base 257, centred
                            u64 q = gi;
                            for (u32 s = 0; s < slot; ++s) q /= OUT_P;
                            v = centred_p((u32)(q % OUT_P));        // zero digits write 0 = "not written" (detector.rs:300-313)
                        }
                    }
                }
                x[k] = v;
            }
        } else {
            const unsigned short* pl = pa.payloads + m * PAYLOAD_LEN;
#pragma unroll
            for (int k = 0; k < E; ++k) {
                const u32 pos = t + GEO::NT * k;
                const u32 j = pos / PAYLOAD_LEN, kk = pos % PAYLOAD_LEN;
                u64 v = 0;
                if (j < pa.cmb_per_cipher) {
                    const u32 w = pa.weights[(size_t)(cipher * pa.cmb_per_cipher + j) * pa.weight_stride + gi];
                    v = centred_p(((u32)pl[kk] * w) % OUT_P);
                }
                x[k] = v;
            }
        }
        ntt_forward<AR, GEO, LdGlobal>(x, eb, tb.tw2, t, 0);
        const u64* pa_ = pv + m * 2 * N; const u64* pb_ = pa_ + N;
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const int idx = out_idx<GEO>(t, k);
            F::mac(ma[k], x[k], __ldg(pa_ + idx));
            F::mac(mb[k], x[k], __ldg(pb_ + idx));
        }
    }
    u64* o = partial + ((size_t)cipher * gridDim.y + chunk) * 2 * N;
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const int idx = out_idx<GEO>(t, k);
        o[idx] = F::csub(F::mul_shoup(F::redc(ma[k]), tb.r2), F::Q);
        o[N + idx] = F::csub(F::mul_shoup(F::redc(mb[k]), tb.r2), F::Q);
    }
}

// out[cipher][e] = SUM_chunk partial[cipher][chunk][e] mod q2
__global__ void reduce_partials_kernel(const u64* __restrict__ partial, int n_chunks, u64* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, cipher = blockIdx.y;
    if (e >= 2 * F2::N) return;
    const u64* p = partial + (size_t)cipher * n_chunks * 2 * F2::N + e;
    u64 s = 0;
    for (int c = 0; c < n_chunks; ++c) { s += p[(size_t)c * 2 * F2::N]; if ((c & 4095) == 4095) s = F2::canon_lazy(s); }
    out[(size_t)cipher * 2 * F2::N + e] = F2::canon_lazy(s);
}
// streaming: running digest += digest of the newly detected messages (mod q2); both canonical
__global__ void digest_add_kernel(u64* acc, const u64* __restrict__ part, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acc[i] = F2::csub(acc[i] + part[i], F2::Q);
}
__global__ void digest_mod_kernel(u64* words, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) words[i] = F2::canon_lazy(words[i]);
}

// ---- recipient side: decrypt + decode a batch of NTT-domain RLWE ciphertexts -----------------------------
// d = b - a (.) z2 (retriever.rs:79,339: sub_mul), then after the inverse NTT every coefficient c is decoded as
// t = round_half_up(c * p / q2), t >= p -> t - p (retriever.rs:84-89, 349-355) in exact integer arithmetic.
__global__ void decrypt_kernel(const u64* __restrict__ ct /*[n][2][N]*/, const u64* __restrict__ z_ntt /*[N]*/, u64* __restrict__ out /*[n][N]*/, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * F2::N) return;
    const size_t m = i / F2::N, k = i % F2::N;
    const u64 a = ct[m * 2 * F2::N + k], b = ct[m * 2 * F2::N + F2::N + k];
    const u64 az = (u64)(((u128)a * z_ntt[k]) % Q2);
    out[i] = b >= az ? b - az : b + Q2 - az;
}
__global__ void decode_round_kernel(const u64* __restrict__ in /*[n][N] canonical*/, unsigned short* __restrict__ out, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    u64 t = (2ull * OUT_P * in[i] + Q2) / (2ull * Q2);        // 514 * c < 2^60: no overflow
    if (t >= OUT_P) t -= OUT_P;
    out[i] = (unsigned short)t;
}

// ---- ChaCha12 (the core of rand 0.8's StdRng) -------------------------------------------------------------------------------
// Used for (i) the combination weights exactly as the reference draws them (below) and (ii) every random draw of the sender
// side, which the reference requires to come from a CryptoRng (key_gen/clue.rs:27-30, sender.rs:27-30).
struct ChaChaKey { u32 k[8]; };
inline ChaChaKey chacha_key_from_seed(const uint8_t* seed32) {
    ChaChaKey key;
    for (int i = 0; i < 8; ++i) key.k[i] = (u32)seed32[4 * i] | ((u32)seed32[4 * i + 1] << 8) | ((u32)seed32[4 * i + 2] << 16) | ((u32)seed32[4 * i + 3] << 24);
    return key;
}
// state words 12,13 = 64-bit counter, 14,15 = nonce (n0, n1); StdRng is counter = block index from 0, nonce = 0
__device__ __forceinline__ void chacha12_block(const ChaChaKey& key, u64 counter, u32 n0, u32 n1, u32 (&out)[16]) {
    const u32 st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                        key.k[4], key.k[5], key.k[6], key.k[7], (u32)counter, (u32)(counter >> 32), n0, n1};
    u32 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = st[i];
#define OMR_QR(a, b, c, d)                                                                                                        \
    x[a] += x[b]; x[d] = __funnelshift_l(x[d] ^ x[a], x[d] ^ x[a], 16); x[c] += x[d]; x[b] = __funnelshift_l(x[b] ^ x[c], x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = __funnelshift_l(x[d] ^ x[a], x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = __funnelshift_l(x[b] ^ x[c], x[b] ^ x[c], 7);
#pragma unroll
    for (int r = 0; r < 12; r += 2) {
        OMR_QR(0, 4, 8, 12) OMR_QR(1, 5, 9, 13) OMR_QR(2, 6, 10, 14) OMR_QR(3, 7, 11, 15)
        OMR_QR(0, 5, 10, 15) OMR_QR(1, 6, 11, 12) OMR_QR(2, 7, 8, 13) OMR_QR(3, 4, 9, 14)
    }
#undef OMR_QR
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = x[i] + st[i];
}

// ---- sender side: batched clue generation -------------------------------------------------------------------
// ClueKey::gen_clues (key_gen/clue.rs:27-34 -> [UPSTREAM] LwePublicKeyRlweMode::encrypt_multi_messages): with the
// public key (pa, pb = pa*s0 + e) over Z_2048[X]/(X^512+1): clue = (pa*r + e1, first 7 coefficients of pb*r + e2 + 256*m),
// r binary.  One CTA per clue.  Randomness: ChaCha12 keyed by the caller's 32-byte seed in counter mode — block
// (counter = global message index, nonce = (domain, block)) with domain 1 = the 512 bits of r (one block), domain 2 = the 512
// errors e1 (64 bits per draw, 8 draws per block), domain 3 = the 7 errors e2 — so clues are independent of batching and GPU
// count and the mask r is unpredictable without the seed.  The rounded Gaussian comes from an integer cumulative table.
// Bit-identical to the oracle's gen_clue_cb.
__constant__ u32 CLUE_CDT[5] = {1947496405u, 3992218608u, 4283915214u, 4294862567u, 4294967049u};   // P(|e| <= k) 2^32, sigma 0.8293
__device__ __forceinline__ int clue_gauss(u64 h) {
    const u32 u = (u32)h; int m = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) m += u >= CLUE_CDT[k];
    return (h >> 63) ? -m : m;
}
constexpr int CLUE_THREADS = 256;
__global__ void __launch_bounds__(CLUE_THREADS)
clue_gen_kernel(const unsigned short* __restrict__ pa, const unsigned short* __restrict__ pb, ChaChaKey key, u64 index0,
                const unsigned char* __restrict__ msgs /*nullable [count][7]*/, unsigned short* __restrict__ out_a, unsigned short* __restrict__ out_b) {
    __shared__ unsigned short s_pa[CLUE_N], s_pb[CLUE_N];
    __shared__ unsigned char s_r[CLUE_N];
    __shared__ u32 s_rw[16], s_e2[16];
    const u64 index = index0 + blockIdx.x;
    if (threadIdx.x == 0) { u32 w[16]; chacha12_block(key, index, 1u, 0u, w); for (int k = 0; k < 16; ++k) s_rw[k] = w[k]; }
    if (threadIdx.x == 32) { u32 w[16]; chacha12_block(key, index, 3u, 0u, w); for (int k = 0; k < 16; ++k) s_e2[k] = w[k]; }
    for (int j = threadIdx.x; j < CLUE_N; j += CLUE_THREADS) { s_pa[j] = pa[j]; s_pb[j] = pb[j]; }
    __syncthreads();
    for (int j = threadIdx.x; j < CLUE_N; j += CLUE_THREADS) s_r[j] = (unsigned char)((s_rw[j >> 5] >> (j & 31)) & 1u);
    __syncthreads();
    // (p * r)[i] = SUM_{j<=i} p[i-j] r[j] - SUM_{j>i} p[512+i-j] r[j]   (negacyclic, mod 2048 by wrap-around of int32)
    for (int i = threadIdx.x; i < CLUE_N; i += CLUE_THREADS) {
        int acc = 0;
        for (int j = 0; j <= i; ++j) acc += s_r[j] ? (int)s_pa[i - j] : 0;
        for (int j = i + 1; j < CLUE_N; ++j) acc -= s_r[j] ? (int)s_pa[CLUE_N + i - j] : 0;
        u32 w[16]; chacha12_block(key, index, 2u, (u32)(i >> 3), w);
        u64 h = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) if ((i & 7) == k) h = (u64)w[2 * k] | ((u64)w[2 * k + 1] << 32);
        out_a[(size_t)blockIdx.x * CLUE_N + i] = (unsigned short)((acc + clue_gauss(h)) & (CLUE_Q - 1));
    }
    if (threadIdx.x < CLUE_COUNT) {
        const int c = threadIdx.x;
        int acc = 0;
        for (int j = 0; j <= c; ++j) acc += s_r[j] ? (int)s_pb[c - j] : 0;
        for (int j = c + 1; j < CLUE_N; ++j) acc -= s_r[j] ? (int)s_pb[CLUE_N + c - j] : 0;
        const int m = msgs ? (int)(msgs[(size_t)blockIdx.x * CLUE_COUNT + c] & 7) * (CLUE_Q / 8) : 0;
        const u64 h = (u64)s_e2[2 * c] | ((u64)s_e2[2 * c + 1] << 32);
        out_b[(size_t)blockIdx.x * CLUE_COUNT + c] = (unsigned short)((acc + clue_gauss(h) + m) & (CLUE_Q - 1));
    }
}

// ---- combination weights from the reference's 32-byte seed ------------------------------------------------------------------
// detector.rs:376-387 / retriever.rs:215-226: StdRng::from_seed(seed) (rand 0.8: ChaCha12, 64-bit block counter from 0, stream
// 0) drives Uniform::<u16>::new(0, 257).sample_iter: one u32 per draw, v * 257 = hi:lo, accept when lo <= zone (zone = 2^32 - 2:
// one u32 value in 2^32 is rejected), weight = hi.  One thread per 64-byte ChaCha block = 16 draws.  A rejection shifts every
// later draw, so the parallel kernel only flags it and weights_serial_kernel then regenerates the whole stream in order.
constexpr u32 WEIGHT_ZONE = 0xFFFFFFFFu - (u32)((0xFFFFFFFFull - OUT_P + 1) % OUT_P);

__global__ void weights_kernel(ChaChaKey key, size_t count, unsigned short* __restrict__ out, int* __restrict__ rejected) {
    const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blk * 16 >= count) return;
    u32 w[16];
    chacha12_block(key, blk, 0u, 0u, w);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const u64 m = (u64)w[j] * OUT_P;
        if ((u32)m > WEIGHT_ZONE) atomicExch(rejected, 1);
        if (blk * 16 + j < count) out[blk * 16 + j] = (unsigned short)(m >> 32);
    }
}
// exact in-order regeneration; runs only when weights_kernel saw a rejected draw (or on request, for the tests)
__global__ void weights_serial_kernel(ChaChaKey key, size_t count, unsigned short* __restrict__ out, const int* __restrict__ rejected, int force) {
    if (!force && !*rejected) return;
    size_t n = 0;
    for (u64 blk = 0; n < count; ++blk) {
        u32 w[16];
        chacha12_block(key, blk, 0u, 0u, w);
        for (int j = 0; j < 16 && n < count; ++j) {
            const u64 m = (u64)w[j] * OUT_P;
            if ((u32)m <= WEIGHT_ZONE) out[n++] = (unsigned short)(m >> 32);
        }
    }
}

// ---- standalone batched NTTs (key upload in coefficient form, tests, API completeness) -----------------------------
template <class F> struct GeoOf;
template <> struct GeoOf<F1> { typedef GeoL1 G; };
template <> struct GeoOf<F2> { typedef GeoL2 G; };
template <class F> __device__ __forceinline__ const typename F::TW* fwd_tw(const Tables& tb);
template <> __device__ __forceinline__ const uint2* fwd_tw<F1>(const Tables& tb) { return tb.tw1; }
template <> __device__ __forceinline__ const ulonglong2* fwd_tw<F2>(const Tables& tb) { return tb.tw2; }
template <class F> __device__ __forceinline__ const typename F::TW* inv_tw(const Tables& tb);
template <> __device__ __forceinline__ const uint2* inv_tw<F1>(const Tables& tb) { return tb.itw1; }
template <> __device__ __forceinline__ const ulonglong2* inv_tw<F2>(const Tables& tb) { return tb.itw2; }

template <class F, bool INVERSE>
__global__ void __launch_bounds__(256) ntt_kernel(typename F::T* data, Tables tb, typename F::TW n_inv) {
    typedef typename F::T T; typedef typename GeoOf<F>::G GEO; typedef ArInt<F> AR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* bufs = reinterpret_cast<T*>(smem_raw);
    ExBuf<T> eb{bufs, bufs + GEO::BUF};
    const int t = threadIdx.x;
    T* g = data + (size_t)blockIdx.x * F::N;
    T x[GEO::E];
    if (!INVERSE) {
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) x[k] = g[t + GEO::NT * k];
        ntt_forward<AR, GEO, LdGlobal>(x, eb, fwd_tw<F>(tb), t, 0);
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) g[out_idx<GEO>(t, k)] = F::canon_lazy(x[k]);
    } else {
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) x[k] = g[out_idx<GEO>(t, k)];
        ntt_inverse<AR, GEO, LdGlobal>(x, eb, inv_tw<F>(tb), t, 0);
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) g[t + GEO::NT * k] = F::csub(F::mul_shoup(F::canon_lazy(x[k]), n_inv), F::Q);
    }
}
template <class F> constexpr size_t ntt_kernel_smem() { return (size_t)2 * GeoOf<F>::G::BUF * sizeof(typename F::T); }
template <class F> constexpr int ntt_kernel_threads() { return GeoOf<F>::G::NT; }

// key pre-transforms: word -> word * c mod q (c = R * N^-1, Shoup pair); KSK rows re-strided 671 -> 672
template <class F>
__global__ void scale_kernel(const typename F::T* __restrict__ in, typename F::T* __restrict__ out, size_t n, typename F::TW c) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = F::csub(F::mul_shoup(in[i], c), F::Q);
}
// FP64-path key form: word -> centred(word * c mod q2) as a double (c = N2^-1; exact, |value| <= q/2 < 2^53)
__global__ void key_to_double_kernel(const u64* __restrict__ in, double* __restrict__ out, size_t n, ulonglong2 c) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64 v = F2::csub(F2::mul_shoup(in[i], c), F2::Q);
    out[i] = v > (Q2 >> 1) ? -(double)(i64)(Q2 - v) : (double)(i64)v;
}
__global__ void ksk_pad_kernel(const u32* __restrict__ in, u32* __restrict__ out, size_t rows) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * KSK_PAD) return;
    const size_t r = i / KSK_PAD, c = i % KSK_PAD;
    out[i] = c < (size_t)LWE2_STRIDE_IN ? in[r * LWE2_STRIDE_IN + c] : 0u;
}

// ---- step-0 peaks: register-only loops of exactly the forward butterfly the transforms are made of (no memory traffic,
// 8 independent butterflies per thread per iteration) — the denominators of the compute roofline in bench.py.
template <class F>
__global__ void __launch_bounds__(256) mulmod_peak_kernel(typename F::T* sink, typename F::TW w0, int iters) {
    typedef typename F::T T;
    T x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (T)(threadIdx.x * 16 + k + 1);
    typename F::TW w = w0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {                              // Harvey/Shoup: IMAD.HI + 2 IMAD + 2 adds, never reduced
            T u = x[k], v = F::mul_shoup(x[k + 8], w);
            x[k] = u + v; x[k + 8] = u - v + 2 * F::Q;
        }
        if ((it & 7) == 7) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = F::canon_lazy(x[k] & (T)(((T)1 << (F::QBITS + 4)) - 1));
        }
    }
    T acc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += x[k];
    if (acc == (T)0x12345) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
// the FP64 butterfly of the level-2 kernel: 6-op exact mulmod + add/sub = 8 DP instructions
__global__ void __launch_bounds__(256) mulmod_peak_f64_kernel(double* sink, double2 w, int iters) {
    double x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = (double)(threadIdx.x * 16 + k + 1);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const double u = x[k], v = D2::mulmod(x[k + 8], w.x, w.y);
            x[k] = __dadd_rn(u, v); x[k + 8] = __dadd_rn(u, -v);
        }
        if ((it & 3) == 3) {
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = D2::renorm(x[k]);
        }
    }
    double acc = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc += x[k];
    if (acc == 12345.0) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace omr

// kernels.cuh — sm_100a kernels of the InstantOMR detection hot path.
//   K1 l1_blind_rotate_kernel    detector.rs:505-531, 553-557   7 blind rotations over R_q1 + sum
//   K2 keyswitch_kernel          detector.rs:560-596            sample extract, LWE key switch, mod switch, offset
//   K3 l2_blind_rotate_kernel    detector.rs:599-624            blind rotation over R_q2
//   K4 trace_kernel              detector.rs:626-639            *N^-1, homomorphic trace, to NTT form
//   K5 pack_kernel<INDICES>      detector.rs:223-339            index digest
//   K6 pack_kernel<PAYLOADS>     detector.rs:341-453            payload digest
// plus partial-sum reduction, batched standalone NTTs and the key pre-transform.
// Tensor cores are deliberately unused: the contraction here is an exact modular NTT, not a floating-point GEMM.
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

template <class F> __device__ __forceinline__ const typename F::TW* fwd_tw(const Tables& tb);
template <> __device__ __forceinline__ const uint2* fwd_tw<F1>(const Tables& tb) { return tb.tw1; }
template <> __device__ __forceinline__ const ulonglong2* fwd_tw<F2>(const Tables& tb) { return tb.tw2; }
template <class F> __device__ __forceinline__ const typename F::TW* inv_tw(const Tables& tb);
template <> __device__ __forceinline__ const uint2* inv_tw<F1>(const Tables& tb) { return tb.itw1; }
template <> __device__ __forceinline__ const ulonglong2* inv_tw<F2>(const Tables& tb) { return tb.itw2; }

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
// digit r of offset word u as a lazy field element in (0, 2q)
template <class F, class G> __device__ __forceinline__ typename F::T gadget_digit(typename F::S u, int r) {
    typedef typename F::S S; typedef typename F::T T;
    constexpr S B = (S)1 << G::LOGB;
    S d = (r < G::LEVELS - 1) ? (((u >> (G::LOGB * r)) & (B - 1)) - (B >> 1)) : (u >> (G::LOGB * (G::LEVELS - 1)));
    return (T)((S)F::Q + d);
}
// digit r of offset word u as a small signed integer
template <class F, class G> __device__ __forceinline__ int gadget_digit_signed(typename F::S u, int r) {
    typedef typename F::S S;
    constexpr S B = (S)1 << G::LOGB;
    return (int)((r < G::LEVELS - 1) ? (((u >> (G::LOGB * r)) & (B - 1)) - (B >> 1)) : (u >> (G::LOGB * (G::LEVELS - 1))));
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

// ---- one CMux step by a group of NT threads:  acc += ((X^a - 1) acc) [x] RGSW    -------------
// acc: [2][N] canonical, shared memory.  wa, wb: exchange buffers.  key: [2L][2][N] in global memory, each word
// pre-multiplied by R * N^-1 (R = 2^32 / 2^64) so that REDC of the MAC and the unscaled INTT give the exact product.
template <class F, class G>
__device__ __forceinline__ void cmux_group(typename F::T* acc, typename F::T* wa, typename F::T* wb, int a,
                                           const typename F::T* __restrict__ key, const Tables& tb, int t, int bar) {
    typedef typename F::T T; typedef typename F::S S; typedef typename F::Acc Acc;
    constexpr int N = F::N, NT = N / 8, L = G::LEVELS;
    const typename F::TW* tw = fwd_tw<F>(tb);
    Acc ma[8], mb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ma[k] = Acc(); mb[k] = Acc(); }
    int digit_count = 0;
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        const T* ap = acc + p * N;
        S u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = t + NT * k;
            S w = rotated<F>(ap, pos, a) - (S)ap[pos];
            u[k] = gadget_word<F, G>(centre_diff<F>(w));
        }
#pragma unroll 1
        for (int r = 0; r < L; ++r, ++digit_count) {
            T x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = gadget_digit<F, G>(u[k], r);
            T* w = (digit_count & 1) ? wb : wa;
            ntt_forward_regs<F>(x, w, tw, t, bar);
            const T* ka = key + (size_t)(p * L + r) * 2 * N;
            const T* kb = ka + N;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = PassGeom<F, 3>::idx(t, k);
                F::mac(ma[k], x[k], __ldg(ka + idx));
                F::mac(mb[k], x[k], __ldg(kb + idx));
            }
        }
    }
    T ya[8], yb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ya[k] = F::inv_prepare(F::redc(ma[k])); yb[k] = F::inv_prepare(F::redc(mb[k])); }
    group_sync<NT>(bar);                              // last forward pass-3 loads done before wa/wb are reused
    ntt_inverse_regs2<F>(ya, yb, wa, wb, inv_tw<F>(tb), t, bar);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int pos = t + NT * k;
        acc[pos] = F::add_canon(acc[pos], ya[k]);
        acc[N + pos] = F::add_canon(acc[N + pos], yb[k]);
    }
    group_sync<NT>(bar);
}

// ---- the same CMux step for level 2 with the transforms and the MAC on the FP64 pipe (see D2 in field.cuh) ------------
// acc stays canonical u64 in shared memory (the decomposition is integer bit work); digits enter the NTT as doubles,
// key words are centred doubles already multiplied by N^-1, the accumulators are 16 doubles instead of 16 x 128 bits.
__device__ __forceinline__ void cmux_group_f64(u64* acc, double* wa, double* wb, int a, const double* __restrict__ key,
                                               const Tables& tb, int t) {
    typedef F2 F; typedef G2 G;
    constexpr int N = F::N, NT = N / 8, L = G::LEVELS;
    double ma[8], mb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ma[k] = 0.0; mb[k] = 0.0; }
    int digit_count = 0;
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        const u64* ap = acc + p * N;
        i64 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = t + NT * k;
            i64 w = rotated<F>(ap, pos, a) - (i64)ap[pos];
            u[k] = gadget_word<F, G>(centre_diff<F>(w));
        }
#pragma unroll 1
        for (int r = 0; r < L; ++r, ++digit_count) {
            double x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = D2::from_small(gadget_digit_signed<F, G>(u[k], r));
            ntt_forward_regs_d(x, (digit_count & 1) ? wb : wa, tb.tw2d, t);
            const double* ka = key + (size_t)(p * L + r) * 2 * N;
            const double* kb = ka + N;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = PassGeom<F, 3>::idx(t, k);
                ma[k] = __dadd_rn(ma[k], D2::mulmod_key(x[k], __ldg(ka + idx)));      // 12 terms x 0.66q < 2^53
                mb[k] = __dadd_rn(mb[k], D2::mulmod_key(x[k], __ldg(kb + idx)));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { ma[k] = D2::renorm(ma[k]); mb[k] = D2::renorm(mb[k]); }
    __syncthreads();
    ntt_inverse_regs2_d(ma, mb, wa, wb, tb.itw2d, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int pos = t + NT * k;
        i64 va = (i64)acc[pos] + D2::to_i64(ma[k]), vb = (i64)acc[N + pos] + D2::to_i64(mb[k]);
        va += va < 0 ? (i64)F::Q : 0; va -= va >= (i64)F::Q ? (i64)F::Q : 0;
        vb += vb < 0 ? (i64)F::Q : 0; vb -= vb >= (i64)F::Q ? (i64)F::Q : 0;
        acc[pos] = (u64)va; acc[N + pos] = (u64)vb;
    }
    __syncthreads();
}

// acc = (0, LUT * X^(2N - b))  — start of BlindRotationKey::blind_rotate (detector.rs:555,623)
template <class F> __device__ __forceinline__ void init_acc(typename F::T* acc, const typename F::T* __restrict__ lut, int b, int t) {
    typedef typename F::S S;
    constexpr int N = F::N, NT = N / 8;
    const int rot = (2 * N - b) % (2 * N);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int pos = t + NT * k;
        S v = rotated<F>(lut, pos, rot);
        acc[pos] = 0;
        acc[N + pos] = (typename F::T)(v < 0 ? v + (S)F::Q : v);
    }
}

// ---- K1: first-level blind rotations + sum -----------------------------------------------------------------------
// one CTA per message, 7 groups of 128 threads (one per clue), each with its own accumulator; groups synchronise
// with named barriers only.  clue extraction (CmLwe::extract_all, detector.rs:514) is index arithmetic on the fly.
constexpr int L1_THREADS = CLUE_COUNT * 128;
constexpr size_t L1_SMEM = (size_t)CLUE_COUNT * 4 * F1::N * sizeof(u32) + CLUE_N * sizeof(unsigned short) + 16;

__global__ void __launch_bounds__(L1_THREADS, 1)
l1_blind_rotate_kernel(const unsigned short* __restrict__ clue_a, const unsigned short* __restrict__ clue_b,
                       const u32* __restrict__ bsk1, u32* __restrict__ out, Tables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* smem = reinterpret_cast<u32*>(smem_raw);
    const int msg = blockIdx.x, c = threadIdx.x >> 7, t = threadIdx.x & 127;
    u32* acc = smem + (size_t)c * 4 * F1::N;      // [2][N]
    u32* wa = acc + 2 * F1::N; u32* wb = wa + F1::N;
    unsigned short* ca = reinterpret_cast<unsigned short*>(smem + (size_t)CLUE_COUNT * 4 * F1::N);
    for (int i = threadIdx.x; i < CLUE_N; i += L1_THREADS) ca[i] = clue_a[(size_t)msg * CLUE_N + i];
    const int b = clue_b[(size_t)msg * CLUE_COUNT + c];
    init_acc<F1>(acc, tb.lut1, b, t);
    __syncthreads();
    const int bar = 1 + c;
#pragma unroll 1
    for (int i = 0; i < CLUE_N; ++i) {
        // a^(c)_i = a_{c-i} (i <= c), -a_{512+c-i} (i > c)   SURVEY A.3
        const int a = i <= c ? ca[c - i] : ((CLUE_Q - ca[CLUE_N + c - i]) & (CLUE_Q - 1));
        if (a != 0) cmux_group<F1, G1>(acc, wa, wb, a, bsk1 + (size_t)i * 2 * G1::LEVELS * 2 * F1::N, tb, t, bar);
    }
    __syncthreads();
    // sum of the 7 accumulators (add_element_wise, detector.rs:556)
    for (int e = threadIdx.x; e < 2 * F1::N; e += L1_THREADS) {
        u32 s = 0;
#pragma unroll
        for (int cc = 0; cc < CLUE_COUNT; ++cc) s += smem[(size_t)cc * 4 * F1::N + e];    // 7q < 2^30
        out[(size_t)msg * 2 * F1::N + e] = s % Q1;
    }
}

// ---- K3: second-level blind rotation ------------------------------------------------------------------------------
constexpr int L2_THREADS = 256;
constexpr size_t L2_SMEM = (size_t)4 * F2::N * sizeof(u64) + 672 * sizeof(unsigned short);

__global__ void __launch_bounds__(L2_THREADS, 2)
l2_blind_rotate_kernel(const u32* __restrict__ lwe, const double* __restrict__ bsk2, u64* __restrict__ out, Tables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);
    double* wa = reinterpret_cast<double*>(acc + 2 * F2::N); double* wb = wa + F2::N;
    unsigned short* la = reinterpret_cast<unsigned short*>(wb + F2::N);
    const int msg = blockIdx.x, t = threadIdx.x;
    for (int i = t; i < LWE2_STRIDE_IN; i += L2_THREADS) la[i] = (unsigned short)lwe[(size_t)msg * LWE2_STRIDE_IN + i];
    __syncthreads();
    init_acc<F2>(acc, tb.lut2, la[LWE2_N], t);
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < LWE2_N; ++i) {
        const int a = la[i];
        if (a != 0) cmux_group_f64(acc, wa, wb, a, bsk2 + (size_t)i * 2 * G2::LEVELS * 2 * F2::N, tb, t);
    }
    for (int e = t; e < 2 * F2::N; e += L2_THREADS) out[(size_t)msg * 2 * F2::N + e] = acc[e];
}

// ---- K2: sample extraction + LWE key switch + modulus switch + offset ---------------------------------------------
// out[col] = (0,..,0,b0) - SUM_{i<1024, j<27} d_ij * KSK[i][j][col]  (d = balanced base-2 digits of a'_i), then
// x -> round(x * 4096 / q1) mod 4096, b += 7 * 128.   A {-1,0,1} x u32 integer product: thread = column,
// KS_MB messages per CTA share each key row load.
constexpr int KS_MB = 8, KS_THREADS = 128;

__global__ void __launch_bounds__(KS_THREADS)
keyswitch_kernel(const u32* __restrict__ rlwe, const u32* __restrict__ ksk, u32* __restrict__ out, int B) {
    __shared__ i32 uw[KS_MB][F1::N];
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
#pragma unroll 1
    for (int i = 0; i < F1::N; ++i) {
        i32 w[KS_MB];
#pragma unroll
        for (int m = 0; m < KS_MB; ++m) w[m] = uw[m][i];
        const u32* row = ksk + (size_t)i * KS_LEVELS * KSK_PAD + col;
#pragma unroll 9
        for (int j = 0; j < KS_LEVELS; ++j) {
            const i32 k = (i32)__ldg(row + (size_t)j * KSK_PAD);
#pragma unroll
            for (int m = 0; m < KS_MB; ++m) {
                const i32 d = j < KS_LEVELS - 1 ? ((w[m] >> j) & 1) - 1 : (w[m] >> (KS_LEVELS - 1));
                acc[m] += (i64)d * k;
            }
        }
    }
    if (col > LWE2_N) return;
#pragma unroll
    for (int m = 0; m < KS_MB; ++m) {
        if (m0 + m >= B) break;
        i64 s = acc[m] % (i64)Q1; if (s < 0) s += Q1;
        const u32 base = col == LWE2_N ? rlwe[(size_t)(m0 + m) * 2 * F1::N + F1::N] : 0u;   // b0
        const u32 x = base >= (u32)s ? base - (u32)s : base + Q1 - (u32)s;
        u32 y = (u32)(((u64)2 * LWE2_Q * x + Q1) / (2ull * Q1)) & (LWE2_Q - 1);
        if (col == LWE2_N) y = (y + CLUE_COUNT * (LWE2_Q >> 5)) & (LWE2_Q - 1);
        out[(size_t)(m0 + m) * LWE2_STRIDE_IN + col] = y;
    }
}

// ---- K4: scale by N^-1, homomorphic trace, forward NTT ------------------------------------------------------------
constexpr int TR_THREADS = 256;
constexpr size_t TR_SMEM = (size_t)4 * F2::N * sizeof(u64);

// sigma_d(p)[pos] as a signed value: source index i0 = pos * d^-1 mod 2N
__device__ __forceinline__ i64 automorphed(const u64* p, int pos, u32 dinv) {
    const u32 i0 = ((u32)pos * dinv) & (2 * F2::N - 1);
    return i0 < (u32)F2::N ? (i64)p[i0] : -(i64)p[i0 - F2::N];
}

__global__ void __launch_bounds__(TR_THREADS, 2)
trace_kernel(u64* __restrict__ ct, const u64* __restrict__ trk, Tables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* acc = reinterpret_cast<u64*>(smem_raw);        // [2][N]: a, b
    u64* wa = acc + 2 * F2::N; u64* wb = wa + F2::N;
    typedef F2 F; typedef F::Acc Acc;
    constexpr int N = F::N, NT = N / 8;
    const int t = threadIdx.x;
    u64* g = ct + (size_t)blockIdx.x * 2 * N;
    for (int e = t; e < 2 * N; e += TR_THREADS) acc[e] = F::csub(F::mul_shoup(g[e], tb.n2_inv), F::Q);   // detector.rs:635-636
    __syncthreads();
#pragma unroll 1
    for (int step = 0; step < TR_STEPS; ++step) {
        const u32 dinv = tb.trace_dinv[step];
        i64 u[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            i64 v = automorphed(acc, t + NT * k, dinv);              // in (-q, q)
            constexpr i64 H = (i64)(F::Q >> 1);
            if (v > H) v -= (i64)F::Q;
            if (v < -H) v += (i64)F::Q;
            u[k] = gadget_word<F, GT>(v);
        }
        Acc ma[8], mb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ma[k] = Acc(); mb[k] = Acc(); }
        const u64* key = trk + (size_t)step * TR_LEVELS * 2 * N;
#pragma unroll 1
        for (int r = 0; r < TR_LEVELS; ++r) {
            u64 x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = gadget_digit<F, GT>(u[k], r);
            ntt_forward_regs<F>(x, (r & 1) ? wb : wa, tb.tw2, t, 0);
            const u64* ka = key + (size_t)r * 2 * N; const u64* kb = ka + N;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = PassGeom<F, 3>::idx(t, k);
                F::mac(ma[k], x[k], __ldg(ka + idx));
                F::mac(mb[k], x[k], __ldg(kb + idx));
            }
        }
        u64 ya[8], yb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { ya[k] = F::redc(ma[k]); yb[k] = F::redc(mb[k]); }
        __syncthreads();
        ntt_inverse_regs2<F>(ya, yb, wa, wb, tb.itw2, t, 0);
        // b' = b + ks.b + sigma_d(b): read the permuted b before anyone overwrites it
        u64 sb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { i64 v = automorphed(acc + N, t + NT * k, dinv); sb[k] = (u64)(v < 0 ? v + (i64)F::Q : v); }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int pos = t + NT * k;
            acc[pos] = F::add_canon(acc[pos], ya[k]);
            acc[N + pos] = F::add_canon(F::csub(acc[N + pos] + sb[k], F::Q), yb[k]);
        }
        __syncthreads();
    }
    // to_ntt_rlwe (detector.rs:638): forward NTT of a and b, canonical output
#pragma unroll 1
    for (int p = 0; p < 2; ++p) {
        u64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = acc[p * N + t + NT * k];
        ntt_forward_regs<F>(x, p ? wb : wa, tb.tw2, t, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[p * N + PassGeom<F, 3>::idx(t, k)] = F::canon_lazy(x[k]);
    }
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
constexpr int PACK_THREADS = 256, PACK_CHUNK = 128;
constexpr size_t PACK_SMEM = (size_t)2 * F2::N * sizeof(u64);

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
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64* wa = reinterpret_cast<u64*>(smem_raw); u64* wb = wa + F2::N;
    typedef F2 F; typedef F::Acc Acc;
    constexpr int N = F::N, NT = N / 8;
    const int t = threadIdx.x, cipher = blockIdx.x, chunk = blockIdx.y;
    const size_t m_begin = (size_t)chunk * PACK_CHUNK, m_end = min(count, m_begin + PACK_CHUNK);
    Acc ma[8], mb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ma[k] = Acc(); mb[k] = Acc(); }
    int parity = 0;
#pragma unroll 1
    for (size_t m = m_begin; m < m_end; ++m, parity ^= 1) {
        const u64 gi = index0 + m;
        u64 x[8];
        if (INDICES) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const u32 pos = t + NT * k;
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
            for (int k = 0; k < 8; ++k) {
                const u32 pos = t + NT * k;
                const u32 j = pos / PAYLOAD_LEN, kk = pos % PAYLOAD_LEN;
                u64 v = 0;
                if (j < pa.cmb_per_cipher) {
                    const u32 w = pa.weights[(size_t)(cipher * pa.cmb_per_cipher + j) * pa.weight_stride + gi];
                    v = centred_p(((u32)pl[kk] * w) % OUT_P);
                }
                x[k] = v;
            }
        }
        ntt_forward_regs<F>(x, parity ? wb : wa, tb.tw2, t, 0);
        const u64* pa_ = pv + m * 2 * N; const u64* pb_ = pa_ + N;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int idx = PassGeom<F, 3>::idx(t, k);
            F::mac(ma[k], x[k], __ldg(pa_ + idx));
            F::mac(mb[k], x[k], __ldg(pb_ + idx));
        }
    }
    u64* o = partial + ((size_t)cipher * gridDim.y + chunk) * 2 * N;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int idx = PassGeom<F, 3>::idx(t, k);
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
__global__ void digest_mod_kernel(u64* words, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) words[i] = F2::canon_lazy(words[i]);
}

// ---- standalone batched NTTs (key upload in coefficient form, tests, API completeness) -----------------------------
template <class F, bool INVERSE>
__global__ void __launch_bounds__(F::N / 8) ntt_kernel(typename F::T* data, Tables tb, typename F::TW n_inv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef typename F::T T;
    T* w = reinterpret_cast<T*>(smem_raw);
    constexpr int NT = F::N / 8;
    const int t = threadIdx.x;
    T* g = data + (size_t)blockIdx.x * F::N;
    T x[8];
    if (!INVERSE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = g[t + NT * k];
        ntt_forward_regs<F>(x, w, fwd_tw<F>(tb), t, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[PassGeom<F, 3>::idx(t, k)] = F::canon_lazy(x[k]);
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = g[PassGeom<F, 3>::idx(t, k)];
        ntt_inverse_regs<F>(x, w, inv_tw<F>(tb), t, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[t + NT * k] = F::csub(F::mul_shoup(F::canon_lazy(x[k]), n_inv), F::Q);
    }
}

// key pre-transform: word -> word * c mod q (c = R * N^-1, Shoup pair), optionally re-striding rows (KSK 671 -> 672)
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

// ---- step-0 peak: register-only loop of the Shoup butterfly the NTTs use (the denominator of the integer roofline) --
// every thread runs 8 independent forward butterflies per iteration (= 8 mulmods + their add/sub), no memory traffic.
template <class F>
__global__ void __launch_bounds__(256) mulmod_peak_kernel(typename F::T* sink, typename F::TW w0, int iters) {
    typedef typename F::T T;
    T x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (T)(threadIdx.x * 8 + k + 1);
    typename F::TW w = w0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T u = x[k], v = F::mul_shoup(x[k + 4], w);
            x[k] = u + v; x[k + 4] = u - v + 2 * F::Q;
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            T u = x[k], v = F::mul_shoup(x[k + 1], w);
            x[k] = F::canon_lazy(u + v); x[k + 1] = F::canon_lazy(u - v + 2 * F::Q);
        }
    }
    T acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x[k];
    if (acc == (T)0x12345) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace omr

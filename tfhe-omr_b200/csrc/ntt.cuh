// ntt.cuh — negacyclic NTT / INTT over Z_q[X]/(X^N+1), cooperative over a group of NT threads, E coefficients per
// thread held in registers.  Replaces [UPSTREAM] Primus-fhe algebra::ntt (concrete-ntt plans) as used by blind_rotate /
// trace / transform_slice (detector.rs:325,435,555,623,638).  Same transform as the oracle: Cooley-Tukey
// forward, natural order in, bit-reversed order out; Gentleman-Sande inverse (unscaled: N^-1 lives in the key material).
//
// The log2(N) stages are cut into passes (L1: 4+4+2 with 16 coefficients/thread, L2: 3+3+3+2 with 8).  Inside a pass
// all butterflies are register-to-register; between passes the group exchanges through shared memory.  Each exchange
// has its own padded layout  phys(i) = i + A*(i >> S)  chosen (scripts/layout_check.py) so that both the writer and the
// reader are bank-conflict free AND every address is (per-thread base) + (compile-time constant): the inner loops carry
// no index arithmetic, and the last pass moves 16-byte vectors.  Exchanges alternate between two buffers so one barrier
// per exchange suffices.
#pragma once
#include "field.cuh"

namespace omr {

template <int N_, int LOGN_, int NT_, int E_, int NP_, int NS0, int NS1, int NS2, int NS3, int A0, int S0_, int A1, int S1_, int A2, int S2_, bool WARP_BLOCKED_ = false>
struct GeoT {
    static constexpr int N = N_, LOGN = LOGN_, NT = NT_, E = E_, NPASS = NP_;
    // WARP_BLOCKED: every pass after the first keeps warp w inside coefficients [w * 32 * E, (w + 1) * 32 * E) — the last pass takes
    // its register groups from consecutive 32-thread slices of the warp's own block instead of from NT-strided ones — so all
    // exchanges but the first are private to a warp and need __syncwarp() instead of a barrier over the whole group.
    static constexpr bool WARP_BLOCKED = WARP_BLOCKED_;
    __host__ __device__ static constexpr bool warp_local(int x) { return WARP_BLOCKED && x >= 1; }
    __host__ __device__ static constexpr int ns(int p) { return p == 0 ? NS0 : p == 1 ? NS1 : p == 2 ? NS2 : NS3; }
    __host__ __device__ static constexpr int s0(int p) { return p == 0 ? 0 : p == 1 ? NS0 : p == 2 ? NS0 + NS1 : NS0 + NS1 + NS2; }
    __host__ __device__ static constexpr int pad_a(int x) { return x == 0 ? A0 : x == 1 ? A1 : A2; }
    __host__ __device__ static constexpr int pad_s(int x) { return x == 0 ? S0_ : x == 1 ? S1_ : S2_; }
    __host__ __device__ static constexpr int phys(int x, int i) { return pad_a(x) ? i + pad_a(x) * (i >> pad_s(x)) : i; }
    __host__ __device__ static constexpr int cmax(int a, int b) { return a > b ? a : b; }
    // elements per exchange buffer (max over exchanges, rounded up to 4)
    static constexpr int BUF = (cmax(cmax(phys(0, N - 1), phys(1, N - 1)), NP_ > 3 ? phys(2, N - 1) : 0) + 1 + 3) & ~3;
};
typedef GeoT<1024, 10, 64, 16, 3, 4, 4, 2, 0, 4, 6, 4, 6, 0, 0> GeoL1;     // 4-byte elements
typedef GeoT<2048, 11, 256, 8, 4, 3, 3, 3, 2, 32, 8, 4, 5, 2, 4, true> GeoL2;   // 8-byte elements; every exchange keeps warp w in [288 w, 288 w + 288)

// Pass P of geometry GEO: element k (0..E-1) of thread t:  group g = k / EP, kk = k % EP, vt = t + NT*g,
//   blk = N >> S0, stride = blk / EP, j = vt / stride, i = vt % stride, idx = j*blk + i + kk*stride.
template <class GEO, int P> struct Pass {
    static constexpr int S0 = GEO::s0(P), NS = GEO::ns(P), EP = 1 << NS, G = GEO::E / EP;
    static constexpr int BLK = GEO::N >> S0, STRIDE = BLK / EP;
    static constexpr bool BLOCKED = GEO::WARP_BLOCKED && P == GEO::NPASS - 1;
    // virtual thread of register group g: NT-strided, or (warp-blocked last pass) the g-th 32-thread slice of the warp's block
    __host__ __device__ static constexpr int vt(int t, int g) { return BLOCKED ? (t / 32) * 32 * G + (t % 32) + 32 * g : t + GEO::NT * g; }
    __host__ __device__ static constexpr int idx(int t, int k) {
        return (vt(t, k / EP) / STRIDE) * BLK + (vt(t, k / EP) % STRIDE) + (k % EP) * STRIDE;
    }
    static __device__ __forceinline__ int block_of(int t, int g) { return vt(t, g) / STRIDE; }
    // physical offset of element k relative to element 0 in exchange X (thread independent, checked by layout_check.py)
    template <int X> __host__ __device__ static constexpr int off(int k) { return GEO::phys(X, idx(0, k)) - GEO::phys(X, idx(0, 0)); }
    template <int X> static __device__ __forceinline__ int base(int t) { return GEO::phys(X, idx(t, 0)); }
};

template <class T, int VE> struct VecOf;
template <> struct VecOf<u32, 4> { typedef uint4 V; };
template <> struct VecOf<u64, 2> { typedef ulonglong2 V; };
template <> struct VecOf<double, 2> { typedef double2 V; };

// store / load the E registers of pass P to / from exchange X
template <class GEO, int P, int X, class T> __device__ __forceinline__ void ex_store(const T (&x)[GEO::E], T* buf, int t) {
    typedef Pass<GEO, P> PS;
    T* b = buf + PS::template base<X>(t);
    if constexpr (PS::STRIDE == 1 && PS::EP * sizeof(T) >= 16) {
        constexpr int VE = 16 / sizeof(T);
        typedef typename VecOf<T, VE>::V V;
#pragma unroll
        for (int k = 0; k < GEO::E; k += VE) {
            V v;
            T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int e = 0; e < VE; ++e) pv[e] = x[k + e];
            *reinterpret_cast<V*>(b + PS::template off<X>(k)) = v;
        }
    } else {
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) b[PS::template off<X>(k)] = x[k];
    }
}
template <class GEO, int P, int X, class T> __device__ __forceinline__ void ex_load(T (&x)[GEO::E], const T* buf, int t) {
    typedef Pass<GEO, P> PS;
    const T* b = buf + PS::template base<X>(t);
    if constexpr (PS::STRIDE == 1 && PS::EP * sizeof(T) >= 16) {
        constexpr int VE = 16 / sizeof(T);
        typedef typename VecOf<T, VE>::V V;
#pragma unroll
        for (int k = 0; k < GEO::E; k += VE) {
            const V v = *reinterpret_cast<const V*>(b + PS::template off<X>(k));
            const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < VE; ++e) x[k + e] = pv[e];
        }
    } else {
#pragma unroll
        for (int k = 0; k < GEO::E; ++k) x[k] = b[PS::template off<X>(k)];
    }
}

// ---- arithmetic policies ---------------------------------------------------------------------------------------------
// integer Harvey/Shoup butterflies; forward never reduces (q1: 21q < 2^32 after 10 stages, q2: 23q << 2^64 after 11)
template <class F> struct ArInt {
    typedef typename F::T T; typedef typename F::TW TW;
    static constexpr int LOGN = F::LOGN;
    template <int S0, int E> static __device__ __forceinline__ void pre_fwd(T (&)[E]) {}
    template <int STAGE> static __device__ __forceinline__ void fwd(T& a, T& b, TW w) {
        const T u = a, v = F::mul_shoup(b, w);
        a = F::add_alu(u, v); b = u - v + 2 * F::Q;
    }
    template <int STAGE> static __device__ __forceinline__ void inv(T& a, T& b, TW w) {
        constexpr int done = LOGN - 1 - STAGE;                 // Gentleman-Sande stages completed before this one
        const T u = a, v = b;
        a = F::inv_add(u, v, done); b = F::mul_shoup(F::inv_sub(u, v, done), w);
    }
};
// level 2 on the FP64 pipe (D2 in field.cuh).  Lazy ranges of the forward transform (|mulmod| < 0.76q, growth 0.76q per
// stage): stages 0-5 from |x| <= 65 reach 4.6q (< 8q exact; mulmod inputs <= 3.8q), renormalise at the start of the pass
// that begins at stage 6, stages 6-9 reach 3.6q, and the last stage renormalises its pass-through operand, so outputs
// are <= 1.3q.
struct ArD2 {
    typedef double T; typedef double2 TW;
    static constexpr int LOGN = 11;
    template <int S0, int E> static __device__ __forceinline__ void pre_fwd(T (&x)[E]) {
        if (S0 == 6) {
#pragma unroll
            for (int k = 0; k < E; ++k) x[k] = D2::renorm(x[k]);
        }
    }
    template <int STAGE> static __device__ __forceinline__ void fwd(T& a, T& b, TW w) {
        const T u = (STAGE == LOGN - 1) ? D2::renorm(a) : a;
        const T v = D2::mulmod(b, w.x, w.y);
        a = __dadd_rn(u, v); b = __dadd_rn(u, -v);
    }
    // inverse: inputs <= 0.5q.  Sums are renormalised only after every second Gentleman-Sande stage (and after the last
    // one): <= 0.66q -> 1.32q -> 2.64q stays below the 4q mulmod bound; differences always go through mulmod (< 0.67q).
    template <int STAGE> static __device__ __forceinline__ void inv(T& a, T& b, TW w) {
        constexpr int done = LOGN - 1 - STAGE;                  // stages completed before this one
        const T u = a, v = b;
        const T s = __dadd_rn(u, v);
        a = ((done & 1) || done == LOGN - 1) ? D2::renorm(s) : s;
        b = D2::mulmod(__dadd_rn(u, -v), w.x, w.y);
    }
};
// The twiddles of the first pass are the same for every thread (block index 0): they live in constant memory and reach
// the multiplier as constant-bank operands — no load instruction, no register.  Filled at context creation.
__constant__ uint2 c_tw1_head[16];       // level 1 forward, indices 1..15 (stages 0-3)
__constant__ double2 c_tw2d_head[8];     // level 2 forward (FP64 form), indices 1..7 (stages 0-2)
template <class TW> __device__ __forceinline__ const TW* const_head();
template <> __device__ __forceinline__ const uint2* const_head<uint2>() { return c_tw1_head; }
template <> __device__ __forceinline__ const double2* const_head<double2>() { return c_tw2d_head; }
template <class TW> struct HasConstHead { static constexpr bool value = false; };
template <> struct HasConstHead<uint2> { static constexpr bool value = true; };
template <> struct HasConstHead<double2> { static constexpr bool value = true; };

struct LdGlobal { static constexpr bool USE_CONST_HEAD = false, DEINT = false; template <class TW> static __device__ __forceinline__ TW ld(const TW* p) { return __ldg(p); } };
struct LdShared { static constexpr bool USE_CONST_HEAD = false, DEINT = false; template <class TW> static __device__ __forceinline__ TW ld(const TW* p) { return *p; } };
struct LdSharedC { static constexpr bool USE_CONST_HEAD = true, DEINT = false; template <class TW> static __device__ __forceinline__ TW ld(const TW* p) { return *p; } };
// Shared-memory tables in DE-INTERLEAVED order (filled by fill_twiddles_deint): the twiddle of stage S0 + l of the pass that
// starts at stage S0, block j, sub-butterfly sb sits at 2^(S0+l) + sb 2^S0 + j instead of 2^(S0+l) + (j << l) + sb.  In the last
// pass every lane has its own j, so the interleaved order makes lanes read at twice (l = 1) the vector stride — a 2-way bank
// conflict on every such load; de-interleaved, consecutive lanes read consecutive words.  D = forward (+ constant head), DI = inverse.
struct LdSharedD { static constexpr bool USE_CONST_HEAD = true, DEINT = true; template <class TW> static __device__ __forceinline__ TW ld(const TW* p) { return *p; } };
struct LdSharedDI { static constexpr bool USE_CONST_HEAD = false, DEINT = true; template <class TW> static __device__ __forceinline__ TW ld(const TW* p) { return *p; } };
template <class LD> __device__ __forceinline__ constexpr int tw_index(int s0, int l, int j, int sb) {
    return LD::DEINT ? (1 << (s0 + l)) + (sb << s0) + j : (1 << (s0 + l)) + (j << l) + sb;
}
// copy a bit-reversed twiddle table (global, interleaved) into shared memory in de-interleaved order
template <class GEO, class TW> __device__ __forceinline__ void fill_twiddles_deint(TW* __restrict__ dst, const TW* __restrict__ src, int tid, int nthreads) {
    for (int i = tid; i < GEO::N; i += nthreads) {
        int o = i;
        if (i > 0) {
            const int s = 31 - __clz(i);                                    // stage
            int p = 0;
#pragma unroll
            for (int q = 1; q < GEO::NPASS; ++q) if (s >= GEO::s0(q)) p = q;
            const int s0 = GEO::s0(p), l = s - s0, r = i - (1 << s);
            o = (1 << s) + ((r & ((1 << l) - 1)) << s0) + (r >> l);
        }
        dst[o] = src[i];
    }
}

template <class AR, class GEO, int P, class LD>
__device__ __forceinline__ void fwd_pass(typename AR::T (&x)[GEO::E], const typename AR::TW* __restrict__ tw, int t) {
    typedef Pass<GEO, P> PS;
    AR::template pre_fwd<PS::S0, GEO::E>(x);
#pragma unroll
    for (int g = 0; g < PS::G; ++g) {
        const int j = PS::block_of(t, g);
#pragma unroll
        for (int l = 0; l < PS::NS; ++l) {
            const int half = PS::EP >> (l + 1);
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                typename AR::TW w;
                if constexpr (P == 0 && HasConstHead<typename AR::TW>::value && LD::USE_CONST_HEAD) w = const_head<typename AR::TW>()[(1 << l) + sb];
                else w = LD::ld(&tw[tw_index<LD>(PS::S0, l, j, sb)]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * PS::EP + sb * 2 * half + h, hi = lo + half;
                    // STAGE must be a compile-time constant: dispatch on l (NS <= 4)
                    if (l == 0) AR::template fwd<PS::S0 + 0>(x[lo], x[hi], w);
                    else if (l == 1) AR::template fwd<PS::S0 + 1>(x[lo], x[hi], w);
                    else if (l == 2) AR::template fwd<PS::S0 + 2>(x[lo], x[hi], w);
                    else AR::template fwd<PS::S0 + 3>(x[lo], x[hi], w);
                }
            }
        }
    }
}
template <class AR, class GEO, int P, class LD>
__device__ __forceinline__ void inv_pass(typename AR::T (&x)[GEO::E], const typename AR::TW* __restrict__ itw, int t) {
    typedef Pass<GEO, P> PS;
#pragma unroll
    for (int g = 0; g < PS::G; ++g) {
        const int j = PS::block_of(t, g);
#pragma unroll
        for (int l = PS::NS - 1; l >= 0; --l) {
            const int half = PS::EP >> (l + 1);
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                const typename AR::TW w = LD::ld(&itw[tw_index<LD>(PS::S0, l, j, sb)]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * PS::EP + sb * 2 * half + h, hi = lo + half;
                    if (l == 0) AR::template inv<PS::S0 + 0>(x[lo], x[hi], w);
                    else if (l == 1) AR::template inv<PS::S0 + 1>(x[lo], x[hi], w);
                    else if (l == 2) AR::template inv<PS::S0 + 2>(x[lo], x[hi], w);
                    else AR::template inv<PS::S0 + 3>(x[lo], x[hi], w);
                }
            }
        }
    }
}

// group barrier: id 0 = whole CTA (__syncthreads), otherwise a named barrier over NTHREADS
template <int NTHREADS> __device__ __forceinline__ void group_sync(int bar_id) {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTHREADS) : "memory");
}

// barrier of exchange X: private to the warp when the geometry keeps it inside one warp's block
template <class GEO, int X> __device__ __forceinline__ void exchange_sync(int bar_id) {
    if constexpr (GEO::warp_local(X)) __syncwarp(); else group_sync<GEO::NT>(bar_id);
}
// With warp-private exchanges the warps of a group drift apart between group barriers, and the FIRST exchange of a forward
// transform stores into every warp's region: a group barrier in front of that store keeps it behind the last private exchange of
// the slowest warp.  (The first exchange of an inverse transform stores into the warp's own region and needs none.)
template <class GEO> __device__ __forceinline__ void pre_first_exchange(int bar_id) {
    if constexpr (GEO::WARP_BLOCKED) group_sync<GEO::NT>(bar_id);
}

// two exchange buffers used alternately by EVERY exchange of the group (forward, inverse, any polynomial): a write to
// one buffer is always separated from the last read of it by the barrier of the exchange in between.
template <class T> struct ExBuf {
    T* a; T* b;
    __device__ __forceinline__ T* next() { T* r = a; a = b; b = r; return r; }
};

// Forward NTT: x holds pass-0 elements (idx = t + NT*k) on entry, last-pass elements on exit (lazy, unreduced).
template <class AR, class GEO, class LD>
__device__ __forceinline__ void ntt_forward(typename AR::T (&x)[GEO::E], ExBuf<typename AR::T>& eb, const typename AR::TW* __restrict__ tw, int t, int bar) {
    typedef typename AR::T T;
    fwd_pass<AR, GEO, 0, LD>(x, tw, t);
    pre_first_exchange<GEO>(bar);
    { T* w = eb.next(); ex_store<GEO, 0, 0>(x, w, t); exchange_sync<GEO, 0>(bar); ex_load<GEO, 1, 0>(x, w, t); }
    fwd_pass<AR, GEO, 1, LD>(x, tw, t);
    { T* w = eb.next(); ex_store<GEO, 1, 1>(x, w, t); exchange_sync<GEO, 1>(bar); ex_load<GEO, 2, 1>(x, w, t); }
    fwd_pass<AR, GEO, 2, LD>(x, tw, t);
    if constexpr (GEO::NPASS == 4) {
        { T* w = eb.next(); ex_store<GEO, 2, 2>(x, w, t); exchange_sync<GEO, 2>(bar); ex_load<GEO, 3, 2>(x, w, t); }
        fwd_pass<AR, GEO, 3, LD>(x, tw, t);
    }
}
// Inverse NTT (unscaled): x holds last-pass elements on entry, pass-0 elements on exit.
template <class AR, class GEO, class LD>
__device__ __forceinline__ void ntt_inverse(typename AR::T (&x)[GEO::E], ExBuf<typename AR::T>& eb, const typename AR::TW* __restrict__ itw, int t, int bar) {
    typedef typename AR::T T;
    if constexpr (GEO::NPASS == 4) {
        inv_pass<AR, GEO, 3, LD>(x, itw, t);
        { T* w = eb.next(); ex_store<GEO, 3, 2>(x, w, t); exchange_sync<GEO, 2>(bar); ex_load<GEO, 2, 2>(x, w, t); }
    }
    inv_pass<AR, GEO, 2, LD>(x, itw, t);
    { T* w = eb.next(); ex_store<GEO, 2, 1>(x, w, t); exchange_sync<GEO, 1>(bar); ex_load<GEO, 1, 1>(x, w, t); }
    inv_pass<AR, GEO, 1, LD>(x, itw, t);
    { T* w = eb.next(); ex_store<GEO, 1, 0>(x, w, t); exchange_sync<GEO, 0>(bar); ex_load<GEO, 0, 0>(x, w, t); }
    inv_pass<AR, GEO, 0, LD>(x, itw, t);
}

// Two transforms through only TWO buffers (x always via `bx`, y always via `by`), staggered so that every store to a
// buffer is separated from the previous loads of it by a barrier: store x | bar | load x, store y | bar | load y.
template <class AR, class GEO, int P, int X, class T>
__device__ __forceinline__ void exchange2s(T (&x)[GEO::E], T (&y)[GEO::E], T* bx, T* by, int t, int bar) {
    ex_store<GEO, P, X>(x, bx, t); exchange_sync<GEO, X>(bar);
    ex_load<GEO, P + 1, X>(x, bx, t); ex_store<GEO, P, X>(y, by, t); exchange_sync<GEO, X>(bar);
    ex_load<GEO, P + 1, X>(y, by, t);
}
template <class AR, class GEO, int P, int X, class T>
__device__ __forceinline__ void exchange2s_inv(T (&x)[GEO::E], T (&y)[GEO::E], T* bx, T* by, int t, int bar) {
    ex_store<GEO, P + 1, X>(x, bx, t); exchange_sync<GEO, X>(bar);
    ex_load<GEO, P, X>(x, bx, t); ex_store<GEO, P + 1, X>(y, by, t); exchange_sync<GEO, X>(bar);
    ex_load<GEO, P, X>(y, by, t);
}
template <class AR, class GEO, class LD>
__device__ __forceinline__ void ntt_forward2s(typename AR::T (&x)[GEO::E], typename AR::T (&y)[GEO::E], typename AR::T* bx, typename AR::T* by,
                                              const typename AR::TW* __restrict__ tw, int t, int bar) {
    fwd_pass<AR, GEO, 0, LD>(x, tw, t); fwd_pass<AR, GEO, 0, LD>(y, tw, t);
    pre_first_exchange<GEO>(bar);
    exchange2s<AR, GEO, 0, 0>(x, y, bx, by, t, bar);
    fwd_pass<AR, GEO, 1, LD>(x, tw, t); fwd_pass<AR, GEO, 1, LD>(y, tw, t);
    exchange2s<AR, GEO, 1, 1>(x, y, bx, by, t, bar);
    fwd_pass<AR, GEO, 2, LD>(x, tw, t); fwd_pass<AR, GEO, 2, LD>(y, tw, t);
    if constexpr (GEO::NPASS == 4) {
        exchange2s<AR, GEO, 2, 2>(x, y, bx, by, t, bar);
        fwd_pass<AR, GEO, 3, LD>(x, tw, t); fwd_pass<AR, GEO, 3, LD>(y, tw, t);
    }
}
template <class AR, class GEO, class LD>
__device__ __forceinline__ void ntt_inverse2s(typename AR::T (&x)[GEO::E], typename AR::T (&y)[GEO::E], typename AR::T* bx, typename AR::T* by,
                                              const typename AR::TW* __restrict__ itw, int t, int bar) {
    if constexpr (GEO::NPASS == 4) {
        inv_pass<AR, GEO, 3, LD>(x, itw, t); inv_pass<AR, GEO, 3, LD>(y, itw, t);
        exchange2s_inv<AR, GEO, 2, 2>(x, y, bx, by, t, bar);
    }
    inv_pass<AR, GEO, 2, LD>(x, itw, t); inv_pass<AR, GEO, 2, LD>(y, itw, t);
    exchange2s_inv<AR, GEO, 1, 1>(x, y, bx, by, t, bar);
    inv_pass<AR, GEO, 1, LD>(x, itw, t); inv_pass<AR, GEO, 1, LD>(y, itw, t);
    exchange2s_inv<AR, GEO, 0, 0>(x, y, bx, by, t, bar);
    inv_pass<AR, GEO, 0, LD>(x, itw, t); inv_pass<AR, GEO, 0, LD>(y, itw, t);
}

// index of element k of thread t after the forward transform (= NTT-domain coefficient index, bit-reversed order)
template <class GEO> __host__ __device__ constexpr int out_idx(int t, int k) { return Pass<GEO, GEO::NPASS - 1>::idx(t, k); }

}  // namespace omr

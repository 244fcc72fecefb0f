// ntt.cuh — negacyclic NTT / INTT over Z_q[X]/(X^N+1), cooperative over a group of NT = N/8 threads.
// Replaces [UPSTREAM] Primus-fhe algebra::ntt (concrete-ntt plans) as used by blind_rotate / trace / transform_slice
// (detector.rs:325,435,555,623,638).  Same transform as the oracle: Cooley-Tukey forward, natural order
// in, bit-reversed order out; Gentleman-Sande inverse (unscaled here: N^-1 is folded into the key material).
//
// Each thread owns 8 coefficients in registers.  The log2(N) stages are cut into passes of 3 (last pass: 1 or 2)
// stages; inside a pass all butterflies are register-to-register, between passes the group exchanges through a
// swizzled shared-memory buffer (one barrier per exchange).  Forward butterflies are Harvey/Shoup and never reduce:
// q1 < 2^27 leaves 21q < 2^32 after 10 stages, q2 < 2^50 leaves 23q << 2^64 after 11.
#pragma once
#include "field.cuh"

namespace omr {

// ---- bank-conflict-free swizzles of the exchange buffer (derived by search; checked in tests/test_layout.py) ----
// u32, N=1024, 32 lanes x 4 B:  b0^=i5, b1^=i5, b2^=i6, b3^=i7, b4^=i7
__device__ __forceinline__ int swz(F1, int x) { return x ^ ((((x >> 5) & 1) * 3) | (((x >> 6) & 1) * 4) | (((x >> 7) & 1) * 24)); }
// u64, N=2048, half-warps of 16 lanes x 8 B:  b0^=i4, b1^=i5, b2^=i5, b3^=i6
__device__ __forceinline__ int swz(F2, int x) { return x ^ (((x >> 4) & 1) | (((x >> 5) & 1) * 6) | (((x >> 6) & 1) * 8)); }

template <class F> struct Plan {
    static constexpr int N = F::N, LOGN = F::LOGN, NT = F::N / 8;
    static constexpr int NPASS = 4;
    static constexpr int LAST_NS = LOGN - 9;   // 1 (N=1024) or 2 (N=2048)
    __host__ __device__ static constexpr int ns(int p) { return p < 3 ? 3 : LAST_NS; }
    __host__ __device__ static constexpr int s0(int p) { return 3 * p; }
};

// Index of element k (0..7) of thread t in pass P:  group g = k / EP, kk = k % EP, vt = t + NT*g,
//   blk = N >> S0, stride = blk / EP, j = vt / stride, i = vt % stride, idx = j*blk + i + kk*stride.
template <class F, int P> struct PassGeom {
    typedef Plan<F> PL;
    static constexpr int S0 = PL::s0(P), NS = PL::ns(P), EP = 1 << NS, G = 8 / EP;
    static constexpr int BLK = F::N >> S0, STRIDE = BLK / EP;
    static __device__ __forceinline__ int block_of(int t, int g) { return (t + PL::NT * g) / STRIDE; }
    static __device__ __forceinline__ int idx(int t, int k) {
        const int g = k / EP, kk = k % EP, vt = t + PL::NT * g;
        const int j = vt / STRIDE, i = vt % STRIDE;
        return j * BLK + i + kk * STRIDE;
    }
};

template <class F, int P> __device__ __forceinline__ void pass_load(typename F::T (&x)[8], const typename F::T* w, int t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = w[swz(F(), PassGeom<F, P>::idx(t, k))];
}
template <class F, int P> __device__ __forceinline__ void pass_store(const typename F::T (&x)[8], typename F::T* w, int t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) w[swz(F(), PassGeom<F, P>::idx(t, k))] = x[k];
}

// forward butterflies of pass P on registers
template <class F, int P> __device__ __forceinline__ void fwd_pass(typename F::T (&x)[8], const typename F::TW* __restrict__ tw, int t) {
    typedef PassGeom<F, P> GEO; typedef typename F::T T;
#pragma unroll
    for (int g = 0; g < GEO::G; ++g) {
        const int j = GEO::block_of(t, g);
#pragma unroll
        for (int l = 0; l < GEO::NS; ++l) {
            const int half = GEO::EP >> (l + 1);
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                const typename F::TW w = __ldg(&tw[(1 << (GEO::S0 + l)) + (j << l) + sb]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * GEO::EP + sb * 2 * half + h, hi = lo + half;
                    T u = x[lo], v = F::mul_shoup(x[hi], w);
                    x[lo] = u + v; x[hi] = u - v + 2 * F::Q;
                }
            }
        }
    }
}

// inverse (Gentleman-Sande) butterflies of pass P on registers; stages run in reverse order
template <class F, int P> __device__ __forceinline__ void inv_pass(typename F::T (&x)[8], const typename F::TW* __restrict__ itw, int t) {
    typedef PassGeom<F, P> GEO; typedef typename F::T T;
#pragma unroll
    for (int g = 0; g < GEO::G; ++g) {
        const int j = GEO::block_of(t, g);
#pragma unroll
        for (int l = GEO::NS - 1; l >= 0; --l) {
            const int half = GEO::EP >> (l + 1);
            const int done = F::LOGN - 1 - (GEO::S0 + l);     // GS stages completed before this one
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                const typename F::TW w = __ldg(&itw[(1 << (GEO::S0 + l)) + (j << l) + sb]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * GEO::EP + sb * 2 * half + h, hi = lo + half;
                    T u = x[lo], v = x[hi];
                    x[lo] = F::inv_add(u, v, done);
                    x[hi] = F::mul_shoup(F::inv_sub(u, v, done), w);
                }
            }
        }
    }
}

// group barrier: id 0 = whole CTA (__syncthreads), otherwise a named barrier over `nthreads`
template <int NTHREADS> __device__ __forceinline__ void group_sync(int bar_id) {
    if (bar_id == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NTHREADS) : "memory");
}

// Forward NTT: x holds pass-0 elements on entry (idx = t + NT*k), pass-3 elements on exit (lazy, unreduced).
// w is the exchange buffer (N elements).  3 barriers.
template <class F> __device__ __forceinline__ void ntt_forward_regs(typename F::T (&x)[8], typename F::T* w,
                                                                      const typename F::TW* __restrict__ tw, int t, int bar) {
    constexpr int NT = Plan<F>::NT;
    fwd_pass<F, 0>(x, tw, t); pass_store<F, 0>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 1>(x, w, t); fwd_pass<F, 1>(x, tw, t); pass_store<F, 1>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 2>(x, w, t); fwd_pass<F, 2>(x, tw, t); pass_store<F, 2>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 3>(x, w, t); fwd_pass<F, 3>(x, tw, t);
}
// Inverse NTT (unscaled): x holds pass-3 elements (< 2q) on entry, pass-0 elements on exit.  3 barriers.
template <class F> __device__ __forceinline__ void ntt_inverse_regs(typename F::T (&x)[8], typename F::T* w,
                                                                      const typename F::TW* __restrict__ itw, int t, int bar) {
    constexpr int NT = Plan<F>::NT;
    inv_pass<F, 3>(x, itw, t); pass_store<F, 3>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 2>(x, w, t); inv_pass<F, 2>(x, itw, t); pass_store<F, 2>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 1>(x, w, t); inv_pass<F, 1>(x, itw, t); pass_store<F, 1>(x, w, t); group_sync<NT>(bar);
    pass_load<F, 0>(x, w, t); inv_pass<F, 0>(x, itw, t);
}
// two inverse NTTs sharing the barriers (buffers wa, wb)
template <class F> __device__ __forceinline__ void ntt_inverse_regs2(typename F::T (&xa)[8], typename F::T (&xb)[8], typename F::T* wa,
                                                                       typename F::T* wb, const typename F::TW* __restrict__ itw, int t, int bar) {
    constexpr int NT = Plan<F>::NT;
    inv_pass<F, 3>(xa, itw, t); inv_pass<F, 3>(xb, itw, t); pass_store<F, 3>(xa, wa, t); pass_store<F, 3>(xb, wb, t); group_sync<NT>(bar);
    pass_load<F, 2>(xa, wa, t); pass_load<F, 2>(xb, wb, t); inv_pass<F, 2>(xa, itw, t); inv_pass<F, 2>(xb, itw, t);
    pass_store<F, 2>(xa, wa, t); pass_store<F, 2>(xb, wb, t); group_sync<NT>(bar);
    pass_load<F, 1>(xa, wa, t); pass_load<F, 1>(xb, wb, t); inv_pass<F, 1>(xa, itw, t); inv_pass<F, 1>(xb, itw, t);
    pass_store<F, 1>(xa, wa, t); pass_store<F, 1>(xb, wb, t); group_sync<NT>(bar);
    pass_load<F, 0>(xa, wa, t); pass_load<F, 0>(xb, wb, t); inv_pass<F, 0>(xa, itw, t); inv_pass<F, 0>(xb, itw, t);
}

// ---- FP64 variants for level 2 (same pass geometry / swizzle as F2; elements are integer-valued doubles) ------------
// Lazy-range schedule of the forward transform (T = mulmod output, |T| < 0.76q; growth 0.76q per stage):
//   stages 0-5 from |x| <= 65 reach 4.56q (< 8q exact; mulmod inputs <= 3.8q), renormalise at the start of pass 2,
//   stages 6-9 reach 3.54q, and the last stage renormalises its pass-through operand so outputs are <= 1.26q.
template <int P> __device__ __forceinline__ void pass_load_d(double (&x)[8], const double* w, int t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = w[swz(F2(), PassGeom<F2, P>::idx(t, k))];
}
template <int P> __device__ __forceinline__ void pass_store_d(const double (&x)[8], double* w, int t) {
#pragma unroll
    for (int k = 0; k < 8; ++k) w[swz(F2(), PassGeom<F2, P>::idx(t, k))] = x[k];
}
template <int P> __device__ __forceinline__ void fwd_pass_d(double (&x)[8], const double2* __restrict__ tw, int t) {
    typedef PassGeom<F2, P> GEO;
    if (P == 2) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = D2::renorm(x[k]);
    }
#pragma unroll
    for (int g = 0; g < GEO::G; ++g) {
        const int j = GEO::block_of(t, g);
#pragma unroll
        for (int l = 0; l < GEO::NS; ++l) {
            const int half = GEO::EP >> (l + 1);
            const bool last = (GEO::S0 + l == F2::LOGN - 1);
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                const double2 w = __ldg(&tw[(1 << (GEO::S0 + l)) + (j << l) + sb]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * GEO::EP + sb * 2 * half + h, hi = lo + half;
                    const double u = last ? D2::renorm(x[lo]) : x[lo];
                    const double v = D2::mulmod(x[hi], w.x, w.y);
                    x[lo] = __dadd_rn(u, v); x[hi] = __dadd_rn(u, -v);
                }
            }
        }
    }
}
// inverse (Gentleman-Sande): inputs |x| < 0.76q; sums are renormalised, differences (< 1.52q) go through mulmod
template <int P> __device__ __forceinline__ void inv_pass_d(double (&x)[8], const double2* __restrict__ itw, int t) {
    typedef PassGeom<F2, P> GEO;
#pragma unroll
    for (int g = 0; g < GEO::G; ++g) {
        const int j = GEO::block_of(t, g);
#pragma unroll
        for (int l = GEO::NS - 1; l >= 0; --l) {
            const int half = GEO::EP >> (l + 1);
#pragma unroll
            for (int sb = 0; sb < (1 << l); ++sb) {
                const double2 w = __ldg(&itw[(1 << (GEO::S0 + l)) + (j << l) + sb]);
#pragma unroll
                for (int h = 0; h < half; ++h) {
                    const int lo = g * GEO::EP + sb * 2 * half + h, hi = lo + half;
                    const double u = x[lo], v = x[hi];
                    x[lo] = D2::renorm(__dadd_rn(u, v));
                    x[hi] = D2::mulmod(__dadd_rn(u, -v), w.x, w.y);
                }
            }
        }
    }
}
__device__ __forceinline__ void ntt_forward_regs_d(double (&x)[8], double* w, const double2* __restrict__ tw, int t) {
    fwd_pass_d<0>(x, tw, t); pass_store_d<0>(x, w, t); __syncthreads();
    pass_load_d<1>(x, w, t); fwd_pass_d<1>(x, tw, t); pass_store_d<1>(x, w, t); __syncthreads();
    pass_load_d<2>(x, w, t); fwd_pass_d<2>(x, tw, t); pass_store_d<2>(x, w, t); __syncthreads();
    pass_load_d<3>(x, w, t); fwd_pass_d<3>(x, tw, t);
}
__device__ __forceinline__ void ntt_inverse_regs2_d(double (&xa)[8], double (&xb)[8], double* wa, double* wb,
                                                    const double2* __restrict__ itw, int t) {
    inv_pass_d<3>(xa, itw, t); inv_pass_d<3>(xb, itw, t); pass_store_d<3>(xa, wa, t); pass_store_d<3>(xb, wb, t); __syncthreads();
    pass_load_d<2>(xa, wa, t); pass_load_d<2>(xb, wb, t); inv_pass_d<2>(xa, itw, t); inv_pass_d<2>(xb, itw, t);
    pass_store_d<2>(xa, wa, t); pass_store_d<2>(xb, wb, t); __syncthreads();
    pass_load_d<1>(xa, wa, t); pass_load_d<1>(xb, wb, t); inv_pass_d<1>(xa, itw, t); inv_pass_d<1>(xb, itw, t);
    pass_store_d<1>(xa, wa, t); pass_store_d<1>(xb, wb, t); __syncthreads();
    pass_load_d<0>(xa, wa, t); pass_load_d<0>(xb, wb, t); inv_pass_d<0>(xa, itw, t); inv_pass_d<0>(xb, itw, t);
}

}  // namespace omr

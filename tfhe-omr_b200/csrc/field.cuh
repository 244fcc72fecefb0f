// field.cuh — modular arithmetic for the two prime fields of the path, device side.
//   F1 = Z_q1, q1 = 134215681 = 2^27 - 2047 (u32)      parameters/mod.rs:18  (FirstLevelField)
//   F2 = Z_q2, q2 = 1125899906826241 = 2^50 - 16383 (u64) parameters/mod.rs:21 (SecondLevelField)
// Replaces [UPSTREAM] Primus-fhe algebra::{Field, modulus::*, reduce::*} as reached from detector.rs.
// All arithmetic is exact mod q; intermediates are kept lazily (unreduced) wherever that is free, and every
// stage boundary is canonical [0,q) so results are bit-identical to the oracle.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace omr {

typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;
typedef unsigned __int128 u128;

constexpr u32 Q1 = 134215681u;
constexpr u64 Q2 = 1125899906826241ull;

// ptxas balances integer adds and moves between the ALU pipe and the FMA pipe (IMAD.IADD, IMAD.MOV) as if IMAD could issue on both
// FMA halves; on B200 every IMAD runs on the heavy half only (ncu: fmalite 0.02 % in the level-1 kernel), the pipe that bounds
// that kernel.  A three-input add cannot be an IMAD: adding this opaque zero (a constant-bank operand: no register, no extra
// instruction) keeps an addition on the ALU pipe.
__constant__ uint32_t c_opaque_zero = 0;

// ---- level 1: 32-bit ---------------------------------------------------------------------------------------------
struct F1 {
    typedef u32 T;
    typedef uint2 TW;    // (w, floor(w * 2^32 / q))
    typedef u64 Acc;     // MAC accumulator: 8 products of (<2^31.4) x (<2^27) fit
    typedef i32 S;
    static constexpr int N = 1024, LOGN = 10;
    static constexpr T Q = Q1;
    static constexpr int QBITS = 27;
    static constexpr T QINV_NEG = 130021375u;   // -q^-1 mod 2^32 (checked at context creation)

    static __device__ __forceinline__ T mul_shoup(T x, TW w) {     // any 32-bit x -> [0, 2q)
        T h = __umulhi(x, w.y);
        return x * w.x - h * Q;
    }
    // x < 32q  ->  [0, q + 2047*32): 2^27 = 2047 (mod q)
    static __device__ __forceinline__ T fold(T x) { return (x & ((1u << 27) - 1)) + (x >> 27) * 2047u; }
    static __device__ __forceinline__ T csub(T x, T m) { return min(x, x - m); }   // x < 2m -> x mod-ish m
    static __device__ __forceinline__ T canon_lazy(T x) { return csub(fold(x), Q); } // x < 32q -> [0,q)
    static __device__ __forceinline__ T add_alu(T a, T b) { return a + b + c_opaque_zero; }    // IADD3 on the ALU pipe, never IMAD.IADD
    static __device__ __forceinline__ void mac(Acc& acc, T x, T k) { acc += (u64)x * k; }
    // Montgomery REDC: acc < 2^62 -> acc * 2^-32 mod q, lazily in [0, acc/2^32 + q)
    static __device__ __forceinline__ T redc(Acc acc) {
        u32 m = (u32)acc * QINV_NEG;
        return (T)((acc + (u64)m * Q) >> 32);
    }
    // inverse-NTT butterfly support: keep sums < 2q
    static __device__ __forceinline__ T inv_add(T x, T y, int) { T s = add_alu(x, y); return min(s, s - 2 * Q); }
    static __device__ __forceinline__ T inv_sub(T x, T y, int) { return x - y + 2 * Q; }
    static __device__ __forceinline__ T inv_prepare(T x) { return fold(x); }        // redc output (<8q) -> < 2q
    // acc (canonical) + delta (inverse output, < 2q) -> canonical
    static __device__ __forceinline__ T add_canon(T a, T d) { T v = add_alu(a, d); v = csub(v, 2 * Q); return csub(v, Q); }
};

// ---- level 2: 64-bit ---------------------------------------------------------------------------------------------
struct F2 {
    typedef u64 T;
    typedef ulonglong2 TW;  // (w, floor(w * 2^64 / q))
    struct Acc { u64 lo, hi; };
    typedef i64 S;
    static constexpr int N = 2048, LOGN = 11;
    static constexpr T Q = Q2;
    static constexpr int QBITS = 50;
    static constexpr T QINV_NEG = 18375807981263503359ull;  // -q^-1 mod 2^64 (checked at context creation)

    static __device__ __forceinline__ T mul_shoup(T x, TW w) {     // any 64-bit x -> [0, 2q)
        T h = __umul64hi(x, w.y);
        return x * w.x - h * Q;
    }
    // x < 2^64 -> [0, 2^50 + 2^28): 2^50 = 16383 (mod q)
    static __device__ __forceinline__ T fold(T x) { return (x & ((1ull << 50) - 1)) + (x >> 50) * 16383ull; }
    static __device__ __forceinline__ T csub(T x, T m) { return x >= m ? x - m : x; }
    static __device__ __forceinline__ T canon_lazy(T x) { return csub(fold(x), Q); }
    static __device__ __forceinline__ T add_alu(T a, T b) { return a + b; }
    static __device__ __forceinline__ void mac(Acc& acc, T x, T k) {
        u64 lo = x * k, hi = __umul64hi(x, k);
        acc.lo += lo; acc.hi += hi + (acc.lo < lo);
    }
    // Montgomery REDC: acc < 2^64 * q -> acc * 2^-64 mod q in [0, acc.hi + q]
    static __device__ __forceinline__ T redc(Acc acc) {
        u64 m = acc.lo * QINV_NEG;
        return acc.hi + __umul64hi(m, Q) + (acc.lo != 0);
    }
    // inverse butterflies: 14 bits of headroom, sums may double for all 11 stages (bound 2q * 2^stage)
    static __device__ __forceinline__ T inv_add(T x, T y, int) { return x + y; }
    static __device__ __forceinline__ T inv_sub(T x, T y, int stage_done) { return x - y + ((2 * Q) << stage_done); }
    static __device__ __forceinline__ T inv_prepare(T x) { return x; }              // redc output already < 2q
    static __device__ __forceinline__ T add_canon(T a, T d) { T v = a + fold(d); v = csub(v, Q); return csub(v, Q); }
};

// ---- level 2 on the FP64 pipe ---------------------------------------------------------------------------------------
// B200 issues 64 DFMA/clk/SM (measured, profiles/r1_pipe_microbench.txt) while a 64-bit integer mulmod costs ~48 issue
// slots (mul.hi.u64 runs at 7.6/clk/SM).  Residues mod q2 < 2^50 are therefore carried as integer-valued doubles and
// multiplied with error-free transformations (two-product by FMA + quotient by the 1.5*2^52 rounding constant): every
// result is an exact integer, so the arithmetic is still exact mod q2 and bit-identical to the integer oracle.
// Invariants: every value is an integer with |x| < 2^53; inputs of mulmod have |x| < 2^52 (= 4q).
struct D2 {
    static constexpr double Q = 1125899906826241.0;
    static constexpr double QINV = 1.0 / 1125899906826241.0;
    static constexpr double MAGIC = 6755399441055744.0;        // 1.5 * 2^52: (v + MAGIC) - MAGIC = rint(v) for |v| <= 2^51
    // x * w mod q for a constant w (|w| <= q/2) with winv = w/q; |x| < 4q  ->  |result| < 0.76 q
    static __device__ __forceinline__ double mulmod(double x, double w, double winv) {
        const double h = __dmul_rn(x, w);
        const double l = __fma_rn(x, w, -h);                                   // x*w = h + l exactly
        const double k = __dadd_rn(__fma_rn(x, winv, MAGIC), -MAGIC);          // rint(x*w/q) +- 0.25
        return __dadd_rn(__fma_rn(-k, Q, h), l);                               // exact: both terms are small integers
    }
    // x * key mod q for a key word without a precomputed quotient (|x| <= 1.3q, |key| <= q/2) -> |result| < 0.66 q
    static __device__ __forceinline__ double mulmod_key(double x, double key) {
        const double h = __dmul_rn(x, key);
        const double l = __fma_rn(x, key, -h);
        const double k = __dadd_rn(__fma_rn(h, QINV, MAGIC), -MAGIC);
        return __dadd_rn(__fma_rn(-k, Q, h), l);
    }
    // x -> x - rint(x/q) q : |x| < 8q -> |result| <= q/2 (+1)
    static __device__ __forceinline__ double renorm(double x) {
        const double k = __dadd_rn(__fma_rn(x, QINV, MAGIC), -MAGIC);
        return __fma_rn(-k, Q, x);
    }
    // small signed integer (|d| < 2^31) -> double without a cvt instruction
    static __device__ __forceinline__ double from_small(int d) {
        return __dadd_rn(__hiloint2double(0x43300000, d + (1 << 20)), -(4503599627370496.0 + 1048576.0));
    }
    // integer-valued double, |x| < 2^51 -> int64 without a cvt instruction
    static __device__ __forceinline__ i64 to_i64(double x) {
        const u64 b = (u64)__double_as_longlong(__dadd_rn(x, MAGIC));          // mantissa = 2^51 + x
        return (i64)(b & ((1ull << 52) - 1)) - ((i64)1 << 51);
    }
};

}  // namespace omr

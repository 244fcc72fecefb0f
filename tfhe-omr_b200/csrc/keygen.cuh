// keygen.cuh — detection-key generation on the GPU (SecretKeyPack::generate_detection_key,
// key_gen/secret.rs:118-178): BSK1 = 512 RGSW_{z1}(s0[i]) (:124-131), KSK = 27 648 LWE_{s2}(z1[i] 2^j) mod q1 (:133-147),
// BSK2 = 670 RGSW_{z2}(s2[i]) (:149-156), trace key = 11 x 25 RLWE_{z2}(-sigma_d(z2) 4^j) (:158-165), all written in the flat
// NTT-native layouts of omr_key_blobs.  The reference draws from an `R: Rng + CryptoRng`; here every draw comes from ChaCha12
// keyed by a caller-supplied 32-byte seed in counter mode, domain-separated through the nonce words, so the key is a pure
// function of (secret, seed) and the CPU oracle reproduces it bit for bit (gen_detection_key_cb):
//   uniform ring / mask element e of a key array (flat index of the a-part, NTT order): block e / 4 of the array's "A" domain,
//     words 4(e % 4) ..: q1: four 27-bit candidates, the first below q1 wins; q2: two 50-bit candidates (lo | hi << 32); if
//     every candidate is rejected (probability < 2^-60) the last one is reduced mod q;
//   error e of a key array (flat index, coefficient order): block e / 8 of the "E" domain, 64 bits h = w[2(e % 8)] | w[..+1] << 32,
//     rounded Gaussian by an integer cumulative table on the low 32 bits, sign = bit 63;
//   KSK error (sigma = 2.0329 * 2^10 = 2081.7, parameters/mod.rs:58-66): 512 x + y with x the table Gaussian of sigma 4.0556 and
//     y = ((h >> 32) & 511) - 256 uniform — variance 512^2 (4.0556^2 + 1/12) = 2081.7^2.
// Sigmas: 3.1859 (first level, parameters/mod.rs:54), 0.3908 (second level and trace, :80,88).
#pragma once
#include "kernels.cuh"

namespace omr {

enum KeygenDomain : u32 { KG_BSK1_A = 16, KG_BSK1_E = 17, KG_KSK_A = 18, KG_KSK_E = 19, KG_BSK2_A = 20, KG_BSK2_E = 21, KG_TRK_A = 22, KG_TRK_E = 23 };

// P(|e| <= k) 2^32 for the rounded Gaussians (scripts/make_cdt.py)
__constant__ u32 KG_CDT_L1[21] = {535621359u, 1555783112u, 2436857004u, 3126965323u, 3617176249u, 3932973494u, 4117471559u, 4215224899u,
                                  4262195442u, 4282663250u, 4290751715u, 4293650440u, 4294592526u, 4294870186u, 4294944398u, 4294962385u,
                                  4294966338u, 4294967126u, 4294967269u, 4294967292u, 4294967295u};
__constant__ u32 KG_CDT_L2[3] = {3432766375u, 4294435154u, 4294967295u};
__constant__ u32 KG_CDT_KS[26] = {421424793u, 1239163273u, 1985968725u, 2627960021u, 3147452532u, 3543144983u, 3826848252u, 4018317373u,
                                  4139952964u, 4212689020u, 4253630692u, 4275323096u, 4286141809u, 4291220698u, 4293465027u, 4294398558u,
                                  4294764063u, 4294898767u, 4294945497u, 4294960755u, 4294965445u, 4294966802u, 4294967172u, 4294967267u,
                                  4294967289u, 4294967295u};
template <int LEN> __device__ __forceinline__ int kg_gauss(const u32 (&cdt)[LEN], u64 h) {
    const u32 u = (u32)h; int m = 0;
#pragma unroll
    for (int k = 0; k < LEN; ++k) m += u >= cdt[k];
    return (h >> 63) ? -m : m;
}
__device__ __forceinline__ u64 kg_draw64(const ChaChaKey& key, u32 domain, u64 e) {
    u32 w[16]; chacha12_block(key, e >> 3, domain, 0u, w);
    u64 h = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) if ((int)(e & 7) == k) h = (u64)w[2 * k] | ((u64)w[2 * k + 1] << 32);
    return h;
}
// 4 consecutive uniform elements (e0 = 4 * block) of a domain
__device__ __forceinline__ void kg_uniform4_q1(const ChaChaKey& key, u32 domain, u64 block, u32 (&out)[4]) {
    u32 w[16]; chacha12_block(key, block, domain, 0u, w);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        u32 v = w[4 * s + 3] & ((1u << 27) - 1);
#pragma unroll
        for (int c = 2; c >= 0; --c) { const u32 cand = w[4 * s + c] & ((1u << 27) - 1); if (cand < Q1) v = cand; }
        out[s] = v >= Q1 ? v - Q1 : v;
    }
}
__device__ __forceinline__ void kg_uniform4_q2(const ChaChaKey& key, u32 domain, u64 block, u64 (&out)[4]) {
    u32 w[16]; chacha12_block(key, block, domain, 0u, w);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const u64 c0 = ((u64)w[4 * s] | ((u64)w[4 * s + 1] << 32)) & ((1ull << 50) - 1), c1 = ((u64)w[4 * s + 2] | ((u64)w[4 * s + 3] << 32)) & ((1ull << 50) - 1);
        const u64 v = c0 < Q2 ? c0 : c1;
        out[s] = v >= Q2 ? v - Q2 : v;
    }
}
__device__ __forceinline__ u32 kg_mulmod1(u32 a, u32 b) { return (u32)(((u64)a * b) % Q1); }
// a b mod q2 for canonical a, b: product < 2^100 = top 2^50 + low, 2^50 = 16383 (mod q2)
__device__ __forceinline__ u64 kg_mulmod2(u64 a, u64 b) {
    const u64 lo = a * b, hi = __umul64hi(a, b);
    const u64 top = (hi << 14) | (lo >> 50), low = lo & ((1ull << 50) - 1);
    return F2::canon_lazy(low + top * 16383ull);
}
template <class F> __device__ __forceinline__ typename F::T kg_mulmod(typename F::T a, typename F::T b);
template <> __device__ __forceinline__ u32 kg_mulmod<F1>(u32 a, u32 b) { return kg_mulmod1(a, b); }
template <> __device__ __forceinline__ u64 kg_mulmod<F2>(u64 a, u64 b) { return kg_mulmod2(a, b); }
template <class F> __device__ __forceinline__ typename F::T kg_addmod(typename F::T a, typename F::T b) { const typename F::T s = a + b; return s >= F::Q ? s - F::Q : s; }

// secret polynomial (coefficients in {-1,0,1} as i32) -> canonical residues
template <class F> __global__ void kg_lift_kernel(const i32* __restrict__ z, typename F::T* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = z[i] < 0 ? (typename F::T)(F::Q - (typename F::T)(-z[i])) : (typename F::T)z[i];
}
// out[t] = sigma_d(z2) for d = 2^(11-t)+1 (coefficient form, canonical): coefficient i -> position i d mod 2N, negated past N
__global__ void kg_automorph_kernel(const u64* __restrict__ z /*[N2] canonical*/, u64* __restrict__ out /*[11][N2]*/) {
    const int t = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F2::N) return;
    const u32 d = (1u << (TR_STEPS - t)) + 1, p = ((u32)i * d) & (2 * F2::N - 1);
    const u64 v = z[i];
    if (p < (u32)F2::N) out[(size_t)t * F2::N + p] = v; else out[(size_t)t * F2::N + p - F2::N] = v ? Q2 - v : 0;
}

// One CTA per RLWE row: e <- Gaussian (coefficient form), NTT(e); a <- uniform (NTT order); b = a z + e + msg.
//   MODE 0: RGSW rows of a blind-rotation key: row = (i, r), r in [0, 2L): msg = -z m g_j (r < L) or m g_j (r >= L), m = bits[i],
//           g_j = 2^(DROP + LOGB j), j = r % L.    MODE 1: trace rows: row = (t, j): msg = -zs[t] 4^j.
template <class F, class G, int MODE, int CDT_LEN>
__global__ void __launch_bounds__(256)
kg_rlwe_rows_kernel(ChaChaKey key, u32 dom_a, u32 dom_e, const typename F::T* __restrict__ z_ntt /*[N]*/, const i32* __restrict__ bits /*MODE 0*/,
                    const typename F::T* __restrict__ zs_ntt /*MODE 1: [steps][N]*/, typename F::T* __restrict__ out /*[rows][2][N]*/, Tables tb) {
    typedef typename F::T T; typedef typename GeoOf<F>::G GEO; typedef ArInt<F> AR;
    constexpr int N = F::N, E = GEO::E, L = G::LEVELS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* bufs = reinterpret_cast<T*>(smem_raw);
    ExBuf<T> eb{bufs, bufs + GEO::BUF};
    const int t = threadIdx.x; const size_t row = blockIdx.x;
    T x[E];
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const u64 h = kg_draw64(key, dom_e, row * N + (u64)(t + GEO::NT * k));
        const int g = CDT_LEN == 21 ? kg_gauss(KG_CDT_L1, h) : kg_gauss(KG_CDT_L2, h);
        x[k] = g < 0 ? (T)(F::Q - (T)(-g)) : (T)g;
    }
    ntt_forward<AR, GEO, LdGlobal>(x, eb, fwd_tw<F>(tb), t, 0);
    // message scale for this row
    T scale; bool times_z, negate; const T* zsrc = z_ntt;
    if (MODE == 0) {
        const int i = (int)(row / (2 * L)), r = (int)(row % (2 * L)), j = r % L;
        const int sh = G::DROP + G::LOGB * j;                                 // < QBITS: 2^sh is canonical
        scale = bits[i] ? (T)((T)1 << sh) : (T)0;
        times_z = r < L; negate = r < L;
    } else {
        const int step = (int)(row / TR_LEVELS), j = (int)(row % TR_LEVELS);
        scale = (T)((T)1 << (GT::DROP + GT::LOGB * j));                        // 4^j, j <= 24: < 2^50
        if (scale >= F::Q) scale -= F::Q;
        times_z = true; negate = true; zsrc = zs_ntt + (size_t)step * N;
    }
    T* oa = out + row * 2 * N; T* ob = oa + N;
    // out_idx(t, k) for k = 0..E-1 covers E/4 (resp. E/2 ... ) runs of consecutive indices; draw the uniform a in blocks of 4
    // consecutive NTT-order elements: element index e = row * N + idx
#pragma unroll
    for (int k = 0; k < E; ++k) {
        const int idx = out_idx<GEO>(t, k);
        const u64 e = row * N + (u64)idx;
        T a4[4];
        if constexpr (sizeof(T) == 4) kg_uniform4_q1(key, dom_a, e >> 2, a4); else kg_uniform4_q2(key, dom_a, e >> 2, a4);
        T a = a4[0];
#pragma unroll
        for (int s = 1; s < 4; ++s) if ((int)(e & 3) == s) a = a4[s];
        T msg = scale;
        if (times_z) msg = kg_mulmod<F>(zsrc[idx], scale);
        if (negate) msg = msg ? F::Q - msg : 0;
        const T en = F::canon_lazy(x[k]);
        oa[idx] = a;
        ob[idx] = kg_addmod<F>(kg_addmod<F>(kg_mulmod<F>(a, z_ntt[idx]), en), msg);
    }
}

// KSK rows (i, j): a[k] uniform (k < 670), b = <a, s2> + e + z1[i] 2^j mod q1 (z1 lifted with -1 -> q1 - 1, secret.rs:134-138)
constexpr int KG_KSK_THREADS = 128;
__global__ void __launch_bounds__(KG_KSK_THREADS)
kg_ksk_rows_kernel(ChaChaKey key, const i32* __restrict__ s2 /*[670]*/, const i32* __restrict__ z1 /*[1024]*/, u32* __restrict__ out /*[rows][671]*/) {
    __shared__ u64 red[KG_KSK_THREADS / 32];
    const size_t row = blockIdx.x; const int i = (int)(row / KS_LEVELS), j = (int)(row % KS_LEVELS), t = threadIdx.x;
    u32* o = out + row * LWE2_STRIDE_IN;
    u64 dot = 0;
    // element index of a[k] in the KSK "A" domain: row * 672 + k (rows padded to a multiple of 4 so that blocks never straddle rows)
    for (int b = t; b < KSK_PAD / 4; b += KG_KSK_THREADS) {
        u32 a4[4]; kg_uniform4_q1(key, KG_KSK_A, (row * KSK_PAD) / 4 + b, a4);
#pragma unroll
        for (int s = 0; s < 4; ++s) { const int k = 4 * b + s; if (k < LWE2_N) { o[k] = a4[s]; if (s2[k]) dot += a4[s]; } }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_down_sync(0xffffffffu, dot, off);
    if ((t & 31) == 0) red[t >> 5] = dot;
    __syncthreads();
    if (t == 0) {
        u64 tot = 0;
        for (int w = 0; w < KG_KSK_THREADS / 32; ++w) tot += red[w];
        const u64 h = kg_draw64(key, KG_KSK_E, row);
        const int x = kg_gauss(KG_CDT_KS, h);
        const i64 e = (i64)512 * x + (i64)((h >> 32) & 511) - 256;
        const u32 ef = e < 0 ? (u32)((i64)Q1 + e) : (u32)e;
        const u32 zi = z1[i] < 0 ? Q1 - 1 : (u32)z1[i];
        const u32 m = kg_mulmod1(zi, (u32)(1u << j));                           // j <= 26: 2^j < q1
        o[LWE2_N] = kg_addmod<F1>(kg_addmod<F1>((u32)(tot % Q1), ef), m);
    }
}

}  // namespace omr

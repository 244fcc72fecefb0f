// ============================================================================
// omr_oracle.hpp — CPU ORACLE for the InstantOMR detection hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build, load or call it.  The CUDA library never links it.
//
// What it restates (reference = /root/reference, crate omr_core):
//   * Detector::detect                       omr_core/src/detector.rs:135-166
//   * first/second level LUTs                omr_core/src/detector.rs:457-503, lut.rs:12-65
//   * extract_clues_and_modulus_switch       omr_core/src/detector.rs:505-531
//   * first_level_bootstrapping              omr_core/src/detector.rs:533-597
//   * second_level_bootstrapping             omr_core/src/detector.rs:599-624
//   * hom_trace                              omr_core/src/detector.rs:626-639
//   * encode_pertinent_indices / _payloads   omr_core/src/detector.rs:223-339, 341-453
//   * Retriever (decode)                     omr_core/src/retriever.rs:63-130,188-260,318-387
//   * solve_matrix_mod_257                   omr_core/src/matrix.rs:164-247
//   * parameters                             omr_core/src/parameters/mod.rs:39-105
//   * RetrievalParams::new                   omr_core/src/parameters/retrieval_params.rs:50-106
//   * key material (who encrypts what)       omr_core/src/key_gen/secret.rs:46-209
//
// PARITY UNPINNED.  Every ring/LWE primitive the reference calls (NTT, gadget
// decomposition, RGSW external product, blind rotation, key switch, modulus
// switch, trace, public-key encryption) lives in Primus-fhe (crates `algebra`,
// `lattice`, `fhe_core`; git branch `omr2`, no rev, no Cargo.lock) with NTTs from
// `concrete-ntt`; none of that source is under /root/reference, there is no
// Rust toolchain in this image, and the reference ships no golden vectors, no
// fixed seeds and no KATs for this path.  The conventions used here are the
// ones fixed in SURVEY.md Appendix A.  What IS pinned, and is checked in tests/:
//   * the reference's own acceptance assertions: omr_core/examples/omd.rs:52-58
//     (decrypted[0]==1, rest 0 / all 0) and omr_time_analyze2.rs:220-240
//     (decoded index set == planted set, solved payloads == originals);
//   * the only golden table in the repo: INV_MOD_257 (matrix.rs:28-41);
//   * every constant derivable from parameters/mod.rs.
// ============================================================================
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <array>
#include <algorithm>
#include <stdexcept>
#include <set>

namespace orc {

using u8 = uint8_t;
using u16 = uint16_t;
using u32 = uint32_t;
using u64 = uint64_t;
using i32 = int32_t;
using i64 = int64_t;
using u128 = unsigned __int128;
using i128 = __int128;

// ---------------------------------------------------------------------------
// Parameters — omr_core/src/parameters/mod.rs:39-105 (OmrParameters::new)
// ---------------------------------------------------------------------------
constexpr int CLUE_N = 512;          // clue LWE dimension              mod.rs:41
constexpr u32 CLUE_Q = 2048;         // clue cipher modulus (pow of 2)  mod.rs:43
constexpr u32 CLUE_T = 8;            // clue plain modulus              mod.rs:42
constexpr double CLUE_SIGMA = 0.8293;   //                              mod.rs:45
constexpr int CLUE_COUNT = 7;        //                                 mod.rs:48

constexpr u32 Q1 = 134215681u;       // FirstLevelField modulus         mod.rs:18
constexpr int N1 = 1024;             //                                 mod.rs:51
constexpr double SIGMA1 = 3.1859;    //                                 mod.rs:54
constexpr int BS1_LOGB = 5, BS1_LEVELS = 4;   // basis(q1,5,Some(4))    mod.rs:55
constexpr int Q1_BITS = 27;
constexpr int BS1_DROP = Q1_BITS - BS1_LOGB * BS1_LEVELS;   // 7

constexpr int KS_LOGB = 1, KS_LEVELS = 27;    // log_modulus 27, base 2 mod.rs:58-66
constexpr double KS_SIGMA = 2.0329 * 1024.0;  //                        mod.rs:65

constexpr int LWE2_N = 670;          // intermediate LWE dimension      mod.rs:69
constexpr u32 LWE2_Q = 4096;         // intermediate cipher modulus     mod.rs:71
constexpr u32 LWE2_T = 32;           // intermediate plain modulus      mod.rs:70

constexpr u64 Q2 = 1125899906826241ull;  // SecondLevelField modulus    mod.rs:21
constexpr int N2 = 2048;             //                                 mod.rs:77
constexpr double SIGMA2 = 0.3908;    //                                 mod.rs:80
constexpr int BS2_LOGB = 7, BS2_LEVELS = 6;   // basis(q2,7,Some(6))    mod.rs:81
constexpr int Q2_BITS = 50;
constexpr int BS2_DROP = Q2_BITS - BS2_LOGB * BS2_LEVELS;   // 8

constexpr int TR_LOGB = 2, TR_LEVELS = 25;    // basis(q2,2,None)       mod.rs:84-90
constexpr int TR_DROP = 0;
constexpr int TR_STEPS = 11;         // log2(N2)

constexpr u64 OUT_P = 257;           // output plain modulus            mod.rs:93
constexpr int PAYLOAD_LEN = 612;     // payload.rs:8

// retrieval layout constants hard-coded by SecretKeyPack::generate_retriever (secret.rs:196-203)
constexpr int BUCKETS_PER_SEGMENT = 130;
constexpr int SEGMENT_COUNT = 25;
constexpr int CMB_PER_CIPHER = 2;

// key blob sizes (elements)
constexpr size_t BSK1_ROWS = 2 * BS1_LEVELS;                 // 8
constexpr size_t BSK1_ELEMS = (size_t)CLUE_N * BSK1_ROWS * 2 * N1;
constexpr size_t KSK_STRIDE = LWE2_N + 1;                    // 671
constexpr size_t KSK_ELEMS = (size_t)N1 * KS_LEVELS * KSK_STRIDE;
constexpr size_t BSK2_ROWS = 2 * BS2_LEVELS;                 // 12
constexpr size_t BSK2_ELEMS = (size_t)LWE2_N * BSK2_ROWS * 2 * N2;
constexpr size_t TRK_ELEMS = (size_t)TR_STEPS * TR_LEVELS * 2 * N2;

// ---------------------------------------------------------------------------
// Modular arithmetic (naive, obviously-correct forms; fast paths are checked
// against these in tests/test_oracle_primitives.py)
// ---------------------------------------------------------------------------
inline u32 addmod(u32 a, u32 b, u32 q) { u32 s = a + b; return s >= q ? s - q : s; }
inline u32 submod(u32 a, u32 b, u32 q) { return a >= b ? a - b : a + q - b; }
inline u32 mulmod(u32 a, u32 b, u32 q) { return (u32)((u64)a * b % q); }
inline u64 addmod(u64 a, u64 b, u64 q) { u64 s = a + b; return s >= q ? s - q : s; }
inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }
inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64)((u128)a * b % q); }

template <class T> inline T powmod(T b, u64 e, T q) {
    T r = 1;
    while (e) { if (e & 1) r = mulmod(r, b, q); b = mulmod(b, b, q); e >>= 1; }
    return r;
}
template <class T> inline T invmod(T a, T q) { return powmod<T>(a, (u64)q - 2, q); }

// Special-form reductions: q1 = 2^27 - 2047, q2 = 2^50 - 16383.
inline u64 reduce128_q2(u128 x) {          // x < 2^114
    const u64 M50 = (1ull << 50) - 1;
    x = (x >> 50) * 16383u + (u64)(x & M50);
    x = (x >> 50) * 16383u + (u64)(x & M50);
    x = (x >> 50) * 16383u + (u64)(x & M50);
    u64 r = (u64)x;
    return r >= Q2 ? r - Q2 : r;
}

inline unsigned bitrev(unsigned x, int bits) {
    unsigned r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// ---------------------------------------------------------------------------
// Negacyclic NTT — SURVEY A.2 [UPSTREAM convention: HEXL/SEAL/concrete-ntt]
//   forward: Cooley-Tukey, natural-order in, bit-reversed-order out,
//            out[k] = a(psi^(2*brv(k)+1)); inverse: Gentleman-Sande, then * N^-1.
//   psi = g^((q-1)/2N) for the smallest g >= 2 whose power has order exactly 2N.
// Reference call sites: detector.rs:325,435 (transform_slice), omd.rs:48,
//   retriever.rs:80,340 (inverse_transform_*); tables: parameters/mod.rs:174-181,238-245.
// ---------------------------------------------------------------------------
template <class T> struct Wide;
template <> struct Wide<u32> { using type = u64; static constexpr int BITS = 32; };
template <> struct Wide<u64> { using type = u128; static constexpr int BITS = 64; };

template <class T> struct NttTable {
    using W = typename Wide<T>::type;
    static constexpr int BITS = Wide<T>::BITS;
    int n = 0, logn = 0;
    T q = 0, psi = 0, psi_inv = 0, n_inv = 0, n_inv_shoup = 0;
    std::vector<T> tw, tw_shoup, itw, itw_shoup;   // bit-reversed power tables

    static T shoup(T w, T q) { return (T)(((W)w << BITS) / q); }
    static T mul_shoup(T x, T w, T ws, T q) {          // result in [0, 2q)
        T hi = (T)(((W)x * ws) >> BITS);
        return (T)(x * w - hi * q);
    }

    NttTable() {}
    NttTable(int n_, T q_) : n(n_), q(q_) {
        logn = 0; while ((1 << logn) < n) ++logn;
        u64 e = ((u64)q - 1) / (2 * (u64)n);
        for (T g = 2;; ++g) {
            T c = powmod<T>(g, e, q);
            if (powmod<T>(c, (u64)n, q) == q - 1) { psi = c; break; }
        }
        psi_inv = invmod<T>(psi, q);
        n_inv = invmod<T>((T)n, q);
        n_inv_shoup = shoup(n_inv, q);
        tw.resize(n); tw_shoup.resize(n); itw.resize(n); itw_shoup.resize(n);
        std::vector<T> pw(n), ipw(n);
        pw[0] = 1; ipw[0] = 1;
        for (int i = 1; i < n; ++i) { pw[i] = mulmod(pw[i - 1], psi, q); ipw[i] = mulmod(ipw[i - 1], psi_inv, q); }
        for (int i = 0; i < n; ++i) {
            unsigned r = bitrev(i, logn);
            tw[i] = pw[r]; tw_shoup[i] = shoup(tw[i], q);
            itw[i] = ipw[r]; itw_shoup[i] = shoup(itw[i], q);
        }
    }

    // in-place forward; input in [0,q) (or lazily < 4q), output canonical [0,q)
    void forward(T* a) const {
        const T two_q = 2 * q;
        int t = n;
        for (int m = 1; m < n; m <<= 1) {
            t >>= 1;
            for (int i = 0; i < m; ++i) {
                const T w = tw[m + i], ws = tw_shoup[m + i];
                T* x = a + 2 * i * t; T* y = x + t;
                for (int j = 0; j < t; ++j) {
                    T u = x[j]; u = u >= two_q ? u - two_q : u;
                    T v = mul_shoup(y[j], w, ws, q);
                    x[j] = u + v; y[j] = u - v + two_q;
                }
            }
        }
        for (int i = 0; i < n; ++i) {
            T v = a[i]; v = v >= two_q ? v - two_q : v; a[i] = v >= q ? v - q : v;
        }
    }

    // forward without the final canonicalisation: output in [0, 4q)
    void forward_lazy(T* a) const {
        const T two_q = 2 * q;
        int t = n;
        for (int m = 1; m < n; m <<= 1) {
            t >>= 1;
            for (int i = 0; i < m; ++i) {
                const T w = tw[m + i], ws = tw_shoup[m + i];
                T* x = a + 2 * i * t; T* y = x + t;
                for (int j = 0; j < t; ++j) {
                    T u = x[j]; u = u >= two_q ? u - two_q : u;
                    T v = mul_shoup(y[j], w, ws, q);
                    x[j] = u + v; y[j] = u - v + two_q;
                }
            }
        }
    }

    // in-place inverse; input in [0,2q), output canonical [0,q)
    void inverse(T* a) const {
        const T two_q = 2 * q;
        int t = 1;
        for (int m = n; m > 1; m >>= 1) {
            int h = m >> 1;
            for (int i = 0; i < h; ++i) {
                const T w = itw[h + i], ws = itw_shoup[h + i];
                T* x = a + 2 * i * t; T* y = x + t;
                for (int j = 0; j < t; ++j) {
                    T u = x[j], v = y[j];
                    T s = u + v; s = s >= two_q ? s - two_q : s;
                    x[j] = s; y[j] = mul_shoup(u - v + two_q, w, ws, q);
                }
            }
            t <<= 1;
        }
        for (int i = 0; i < n; ++i) {
            T v = mul_shoup(a[i], n_inv, n_inv_shoup, q); a[i] = v >= q ? v - q : v;
        }
    }
};

// Schoolbook negacyclic product (test oracle for the NTT itself).
template <class T> std::vector<T> negacyclic_schoolbook(const T* a, const T* b, int n, T q) {
    std::vector<T> c(n, 0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) {
        T p = mulmod(a[i], b[j], q); int k = i + j;
        if (k < n) c[k] = addmod(c[k], p, q); else c[k - n] = submod(c[k - n], p, q);
    }
    return c;
}

// ---------------------------------------------------------------------------
// Signed approximate gadget decomposition — SURVEY A.4 [UPSTREAM: algebra::decompose::
// NonPowOf2ApproxSignedBasis; convention adopted, parity unpinned]
//   centre; round away `drop` low bits; L-1 balanced digits in [-B/2,B/2); the top
//   digit absorbs the remainder.  Row j of a gadget key encrypts m * 2^(drop + w*j).
// ---------------------------------------------------------------------------
template <class T> inline void gadget_decompose(T x, T q, int logb, int levels, int drop, i64* digits) {
    i64 v = (x > (q >> 1)) ? (i64)x - (i64)q : (i64)x;
    if (drop > 0) v = (v + ((i64)1 << (drop - 1))) >> drop;
    const i64 B = (i64)1 << logb, half = B >> 1;
    for (int j = 0; j < levels - 1; ++j) {
        i64 d = v & (B - 1);
        if (d >= half) d -= B;
        digits[j] = d; v = (v - d) >> logb;
    }
    digits[levels - 1] = v;
}
template <class T> inline T signed_to_field(i64 d, T q) { return d >= 0 ? (T)d : (T)((i64)q + d); }

// ---------------------------------------------------------------------------
// Deterministic RNG for keys / clues / test inputs (xoshiro256** seeded by splitmix64).
// The reference draws everything from rand::thread_rng() (omr.rs:73, omd.rs:20) — there
// are no seeds to mirror; this stream is ours.
// ---------------------------------------------------------------------------
struct Rng {
    u64 s[4]; bool has_spare = false; double spare = 0;
    static u64 splitmix(u64& x) {
        x += 0x9E3779B97F4A7C15ull; u64 z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Rng(u64 seed, u64 stream = 0) {
        u64 x = seed ^ (stream * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
        for (auto& v : s) v = splitmix(x);
    }
    static u64 rotl(u64 x, int k) { return (x << k) | (x >> (64 - k)); }
    u64 next() {
        u64 r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    u64 below(u64 q) {               // uniform in [0,q) by mask rejection
        u64 m = q - 1; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; m |= m >> 32;
        for (;;) { u64 v = next() & m; if (v < q) return v; }
    }
    double gauss() {
        if (has_spare) { has_spare = false; return spare; }
        double u1 = ((next() >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        double u2 = (next() >> 11) * (1.0 / 9007199254740992.0);
        double r = std::sqrt(-2.0 * std::log(u1)), th = 6.283185307179586476925 * u2;
        spare = r * std::sin(th); has_spare = true;
        return r * std::cos(th);
    }
    i64 rounded_gauss(double sigma) { return (i64)std::llround(gauss() * sigma); }
};

// ---------------------------------------------------------------------------
// Shared tables (built once)
// ---------------------------------------------------------------------------
struct Tables {
    NttTable<u32> t1; NttTable<u64> t2;
    std::vector<u32> lut1; std::vector<u64> lut2;
    u64 n2_inv;
    Tables() : t1(N1, Q1), t2(N2, Q2) {
        // first_level_lut — detector.rs:457-476 with lut.rs:12-27:
        //   [s,0,0,0,-s].negacyclic_lut(N1, log2(8)=3): chunks of N1>>3=128 coefficients
        //   take v0,v1,v1,v2,v2,v3,v3,v4.
        {
            const int log_out = 5 - 1;                       // trailing_zeros(32) - 1
            const u32 s = ((Q1 >> log_out) + 1) >> 1;        // 4194240
            const u32 vals[5] = {s, 0, 0, 0, Q1 - s};
            lut1.assign(N1, 0);
            const int hd = N1 >> 3;
            for (int c = 0; c < N1 / hd; ++c) {
                int vi = (c + 1) / 2;                        // interleave(v, v[1..])
                if (vi < 5) for (int j = 0; j < hd; ++j) lut1[c * hd + j] = vals[vi];
            }
        }
        // second_level_lut — detector.rs:479-503: 32-entry table, entry clue_count*2 = 14
        //   set to round_half_up(q2/257); negacyclic_lut(N2, log2(32)=5): chunks of 64.
        {
            const u64 s = (2 * Q2 + OUT_P) / (2 * OUT_P);    // round half up of q2/257 = 4380933489596
            std::vector<u64> data(LWE2_T, 0); data[CLUE_COUNT * 2] = s;
            lut2.assign(N2, 0);
            const int hd = N2 >> 5;
            for (int c = 0; c < N2 / hd; ++c) {
                int vi = (c + 1) / 2;
                if (vi < (int)LWE2_T) for (int j = 0; j < hd; ++j) lut2[c * hd + j] = data[vi];
            }
        }
        n2_inv = invmod<u64>((u64)N2, Q2);                  // secret.rs:167-168
    }
};
inline const Tables& tables() { static Tables t; return t; }

// ---------------------------------------------------------------------------
// Key material — key_gen/secret.rs:46-178
// ---------------------------------------------------------------------------
struct SecretKeyPack {
    std::vector<i32> s0;   // clue LWE secret, binary          secret.rs:46 (params mod.rs:44)
    std::vector<i32> z1;   // first-level ring secret, ternary secret.rs:51-58
    std::vector<i32> s2;   // intermediate LWE secret, binary  secret.rs:47-48
    std::vector<i32> z2;   // second-level ring secret, ternary secret.rs:67-75
    std::vector<u64> z2_ntt;
};

inline SecretKeyPack gen_secret_key(u64 seed) {
    SecretKeyPack sk; Rng r(seed, 1);
    sk.s0.resize(CLUE_N); sk.z1.resize(N1); sk.s2.resize(LWE2_N); sk.z2.resize(N2);
    for (auto& v : sk.s0) v = (i32)r.below(2);
    for (auto& v : sk.s2) v = (i32)r.below(2);
    for (auto& v : sk.z1) v = (i32)r.below(3) - 1;
    for (auto& v : sk.z2) v = (i32)r.below(3) - 1;
    sk.z2_ntt.resize(N2);
    for (int i = 0; i < N2; ++i) sk.z2_ntt[i] = signed_to_field<u64>(sk.z2[i], Q2);
    tables().t2.forward(sk.z2_ntt.data());
    return sk;
}

// Clue public key: LwePublicKeyRlweMode over Z_2048[X]/(X^512+1) — secret.rs:99-106 [UPSTREAM]
struct ClueKey { std::vector<u16> pa, pb; };

inline void negacyclic_mul_small(const u16* a, const i32* s, int n, u32 qmask, u16* out) {
    std::vector<i32> acc(n, 0);
    for (int j = 0; j < n; ++j) {
        if (!s[j]) continue;
        const i32 sj = s[j];
        for (int i = 0; i < n - j; ++i) acc[i + j] += sj * (i32)a[i];
        for (int i = n - j; i < n; ++i) acc[i + j - n] -= sj * (i32)a[i];
    }
    for (int i = 0; i < n; ++i) out[i] = (u16)((u32)acc[i] & qmask);
}

inline ClueKey gen_clue_key(const SecretKeyPack& sk, u64 seed) {
    ClueKey k; k.pa.resize(CLUE_N); k.pb.resize(CLUE_N); Rng r(seed, 2);
    for (auto& v : k.pa) v = (u16)r.below(CLUE_Q);
    negacyclic_mul_small(k.pa.data(), sk.s0.data(), CLUE_N, CLUE_Q - 1, k.pb.data());
    for (auto& v : k.pb) v = (u16)(((i64)v + r.rounded_gauss(CLUE_SIGMA)) & (CLUE_Q - 1));
    return k;
}

// ClueKey::gen_clues — key_gen/clue.rs:27-34: encrypt_multi_messages(&[0;7]) [UPSTREAM]; SURVEY A.3.
inline void gen_clue(const ClueKey& k, u64 seed, u64 index, const u32* msgs /*7 values mod 8, may be null = zeros*/,
                     u16* a_out /*512*/, u16* b_out /*7*/) {
    Rng r(seed, 0x1000 + index);
    std::vector<i32> rr(CLUE_N);
    for (auto& v : rr) v = (i32)r.below(2);
    std::vector<u16> u(CLUE_N), v(CLUE_N);
    negacyclic_mul_small(k.pa.data(), rr.data(), CLUE_N, CLUE_Q - 1, u.data());
    negacyclic_mul_small(k.pb.data(), rr.data(), CLUE_N, CLUE_Q - 1, v.data());
    for (int i = 0; i < CLUE_N; ++i) a_out[i] = (u16)(((i64)u[i] + r.rounded_gauss(CLUE_SIGMA)) & (CLUE_Q - 1));
    const u32 delta = CLUE_Q / CLUE_T;
    for (int c = 0; c < CLUE_COUNT; ++c) {
        i64 m = msgs ? (i64)(msgs[c] % CLUE_T) * delta : 0;
        b_out[c] = (u16)(((i64)v[c] + r.rounded_gauss(CLUE_SIGMA) + m) & (CLUE_Q - 1));
    }
}

// Counter-based clue generation: the same public-key encryption as gen_clue, but every random draw comes from
// ChaCha12 keyed by a 32-byte seed in counter mode (the reference requires a CryptoRng here: key_gen/clue.rs:27-30) —
// block(counter = message index, nonce = (domain, block)): domain 1 = the 512 bits of r, domain 2 = e1 (one u64 per draw, 8 per
// block), domain 3 = e2 — and the rounded Gaussian (sigma = 0.8293, parameters/mod.rs:45) comes from an integer cumulative
// table, so the CUDA kernel (csrc: clue_gen_kernel) reproduces it bit for bit (no libm).
//   P(|e| <= k) * 2^32 for k = 0..4, sigma = 0.8293 (tail beyond 5 is < 2^-32)
static const u32 CLUE_CDT[5] = {1947496405u, 3992218608u, 4283915214u, 4294862567u, 4294967049u};
inline void chacha_block(const u32 key[8], u64 counter, u64 stream, int rounds, u32 out[16]);
inline void seed_to_key(const u8 seed[32], u32 key[8]) {
    for (int i = 0; i < 8; ++i) key[i] = (u32)seed[4 * i] | ((u32)seed[4 * i + 1] << 8) | ((u32)seed[4 * i + 2] << 16) | ((u32)seed[4 * i + 3] << 24);
}
inline i32 clue_gauss(u64 h) {
    const u32 u = (u32)h; i32 m = 0;
    for (int k = 0; k < 5; ++k) m += u >= CLUE_CDT[k];
    return (h >> 63) ? -m : m;
}
inline void gen_clue_cb(const ClueKey& k, const u8 seed[32], u64 index, const u8* msgs /*7 values mod 8 or null*/, u16* a_out, u16* b_out) {
    u32 key[8]; seed_to_key(seed, key);
    u32 w[16];
    std::vector<i32> rr(CLUE_N);
    chacha_block(key, index, 1ull, 12, w);
    for (int j = 0; j < CLUE_N; ++j) rr[j] = (i32)((w[j / 32] >> (j % 32)) & 1u);
    std::vector<u16> u(CLUE_N), v(CLUE_N);
    negacyclic_mul_small(k.pa.data(), rr.data(), CLUE_N, CLUE_Q - 1, u.data());
    negacyclic_mul_small(k.pb.data(), rr.data(), CLUE_N, CLUE_Q - 1, v.data());
    for (int i = 0; i < CLUE_N; ++i) {
        if (i % 8 == 0) chacha_block(key, index, 2ull | ((u64)(i / 8) << 32), 12, w);
        const u64 h = (u64)w[2 * (i % 8)] | ((u64)w[2 * (i % 8) + 1] << 32);
        a_out[i] = (u16)(((i64)u[i] + clue_gauss(h)) & (CLUE_Q - 1));
    }
    chacha_block(key, index, 3ull, 12, w);
    const u32 delta = CLUE_Q / CLUE_T;
    for (int c = 0; c < CLUE_COUNT; ++c) {
        const i64 m = msgs ? (i64)(msgs[c] % CLUE_T) * delta : 0;
        const u64 h = (u64)w[2 * c] | ((u64)w[2 * c + 1] << 32);
        b_out[c] = (u16)(((i64)v[c] + clue_gauss(h) + m) & (CLUE_Q - 1));
    }
}

// Flat, NTT-native detection key (layouts = include/omr_b200.h omr_key_blobs):
//   bsk1 [512][8][2][1024] u32   rows 0..3 = RLWE(-z1*m*g_j), rows 4..7 = RLWE(m*g_j); poly 0 = a, 1 = b
//   ksk  [1024][27][671]   u32   (a[670], b)
//   bsk2 [670][12][2][2048] u64
//   trk  [11][25][2][2048]  u64  step index t <-> k = 11 - t
struct DetectionKey {
    std::vector<u32> bsk1, ksk; std::vector<u64> bsk2, trk;
};

template <class T>
inline void rlwe_encrypt_ntt(const NttTable<T>& tab, const T* z_ntt, const T* msg_ntt /*nullable*/,
                             double sigma, Rng& r, T* a_out, T* b_out) {
    const int n = tab.n; const T q = tab.q;
    std::vector<T> e(n);
    for (int i = 0; i < n; ++i) { a_out[i] = (T)r.below(q); e[i] = signed_to_field<T>(r.rounded_gauss(sigma), q); }
    tab.forward(e.data());
    for (int i = 0; i < n; ++i) {
        T v = addmod(mulmod(a_out[i], z_ntt[i], q), e[i], q);
        b_out[i] = msg_ntt ? addmod(v, msg_ntt[i], q) : v;
    }
}

// RGSW(m) for scalar m in {0,1}: BlindRotationKey::generate — secret.rs:124-131,149-156 [UPSTREAM]
template <class T>
inline void rgsw_encrypt_bit(const NttTable<T>& tab, const T* z_ntt, int m, int logb, int levels, int drop,
                             double sigma, Rng& r, T* out /*[2*levels][2][n]*/) {
    const int n = tab.n; const T q = tab.q;
    std::vector<T> msg(n);
    for (int row = 0; row < 2 * levels; ++row) {
        const int j = row % levels;
        T g = powmod<T>(2, (u64)(drop + logb * j), q);
        T gm = m ? g : 0;
        if (row < levels) for (int i = 0; i < n; ++i) msg[i] = submod((T)0, mulmod(z_ntt[i], gm, q), q);  // -z*m*g_j
        else for (int i = 0; i < n; ++i) msg[i] = gm;                                                    // m*g_j (constant poly)
        rlwe_encrypt_ntt(tab, z_ntt, msg.data(), sigma, r, out + ((size_t)row * 2 + 0) * n, out + ((size_t)row * 2 + 1) * n);
    }
}

// sigma_d : X -> X^d on a coefficient vector (d odd).  SURVEY A.5 step 9.
template <class T> inline void automorphism(const T* in, int n, int d, T q, T* out) {
    for (int i = 0; i < n; ++i) {
        unsigned p = ((unsigned long long)i * (unsigned)d) % (2u * n);
        if (p < (unsigned)n) out[p] = in[i]; else out[p - n] = in[i] ? q - in[i] : 0;
    }
}

inline DetectionKey gen_detection_key(const SecretKeyPack& sk, u64 seed) {
    const Tables& tb = tables();
    DetectionKey dk;
    // z1 in NTT form
    std::vector<u32> z1n(N1);
    for (int i = 0; i < N1; ++i) z1n[i] = signed_to_field<u32>(sk.z1[i], Q1);
    tb.t1.forward(z1n.data());
    // BSK1[i] = RGSW_{z1}(s0[i])                                        secret.rs:124-131
    dk.bsk1.resize(BSK1_ELEMS);
    { Rng r(seed, 3);
      for (int i = 0; i < CLUE_N; ++i)
          rgsw_encrypt_bit<u32>(tb.t1, z1n.data(), sk.s0[i], BS1_LOGB, BS1_LEVELS, BS1_DROP, SIGMA1, r,
                                dk.bsk1.data() + (size_t)i * BSK1_ROWS * 2 * N1); }
    // KSK[i][j] = LWE_{s2}(z1[i]*2^j) mod q1, z1 lifted with -1 -> q1-1   secret.rs:133-147
    dk.ksk.resize(KSK_ELEMS);
    { Rng r(seed, 4);
      for (int i = 0; i < N1; ++i) for (int j = 0; j < KS_LEVELS; ++j) {
          u32* row = dk.ksk.data() + ((size_t)i * KS_LEVELS + j) * KSK_STRIDE;
          u64 dot = 0;
          for (int k = 0; k < LWE2_N; ++k) { row[k] = (u32)r.below(Q1); if (sk.s2[k]) dot += row[k]; }
          u32 m = mulmod(signed_to_field<u32>(sk.z1[i], Q1), powmod<u32>(2, j, Q1), Q1);
          u32 e = signed_to_field<u32>(r.rounded_gauss(KS_SIGMA), Q1);
          row[LWE2_N] = addmod(addmod((u32)(dot % Q1), e, Q1), m, Q1);
      } }
    // BSK2[i] = RGSW_{z2}(s2[i])                                        secret.rs:149-156
    dk.bsk2.resize(BSK2_ELEMS);
    { Rng r(seed, 5);
      for (int i = 0; i < LWE2_N; ++i)
          rgsw_encrypt_bit<u64>(tb.t2, sk.z2_ntt.data(), sk.s2[i], BS2_LOGB, BS2_LEVELS, BS2_DROP, SIGMA2, r,
                                dk.bsk2.data() + (size_t)i * BSK2_ROWS * 2 * N2); }
    // TraceKey: for k = 11..1, d = 2^k+1: T_k[j] = RLWE_{z2}(-sigma_d(z2) * 4^j)   secret.rs:158-165
    dk.trk.resize(TRK_ELEMS);
    { Rng r(seed, 6);
      std::vector<u64> zc(N2), zs(N2), msg(N2);
      for (int i = 0; i < N2; ++i) zc[i] = signed_to_field<u64>(sk.z2[i], Q2);
      for (int t = 0; t < TR_STEPS; ++t) {
          const int k = TR_STEPS - t, d = (1 << k) + 1;
          automorphism<u64>(zc.data(), N2, d, Q2, zs.data());
          tb.t2.forward(zs.data());
          for (int j = 0; j < TR_LEVELS; ++j) {
              u64 g = powmod<u64>(2, (u64)(TR_DROP + TR_LOGB * j), Q2);
              for (int i = 0; i < N2; ++i) msg[i] = submod((u64)0, mulmod(zs[i], g, Q2), Q2);
              u64* row = dk.trk.data() + ((size_t)t * TR_LEVELS + j) * 2 * N2;
              rlwe_encrypt_ntt<u64>(tb.t2, sk.z2_ntt.data(), msg.data(), SIGMA2, r, row, row + N2);
          }
      } }
    return dk;
}

// Counter-based detection-key generation: the same keys as gen_detection_key (secret.rs:118-178), but every draw
// comes from ChaCha12 keyed by a 32-byte seed in counter mode (the reference requires a CryptoRng), so that the CUDA key generator
// (csrc/keygen.cuh) can be checked bit for bit.  Stream layout — block(counter, nonce = (domain, 0)):
//   domains 16/17 = BSK1 masks/errors, 18/19 = KSK, 20/21 = BSK2, 22/23 = trace key;
//   mask element e (flat index into the a-parts, NTT order; KSK: row * 672 + k): block e / 4, four words per element: q1 takes the first
//     of four 27-bit candidates below q1, q2 the first of two 50-bit candidates (lo | hi << 32); all rejected -> last one minus q;
//   error e (flat index, coefficient order; KSK: the row): block e / 8, h = w[2(e%8)] | w[2(e%8)+1] << 32; rounded Gaussian from an
//     integer cumulative table on the low 32 bits, sign = bit 63; KSK: 512 x + ((h >> 32) & 511) - 256 with x of sigma 4.0556.
static const u32 KG_CDT_L1[21] = {535621359u, 1555783112u, 2436857004u, 3126965323u, 3617176249u, 3932973494u, 4117471559u, 4215224899u,
                                  4262195442u, 4282663250u, 4290751715u, 4293650440u, 4294592526u, 4294870186u, 4294944398u, 4294962385u,
                                  4294966338u, 4294967126u, 4294967269u, 4294967292u, 4294967295u};     // sigma 3.1859
static const u32 KG_CDT_L2[3] = {3432766375u, 4294435154u, 4294967295u};                               // sigma 0.3908
static const u32 KG_CDT_KS[26] = {421424793u, 1239163273u, 1985968725u, 2627960021u, 3147452532u, 3543144983u, 3826848252u, 4018317373u,
                                  4139952964u, 4212689020u, 4253630692u, 4275323096u, 4286141809u, 4291220698u, 4293465027u, 4294398558u,
                                  4294764063u, 4294898767u, 4294945497u, 4294960755u, 4294965445u, 4294966802u, 4294967172u, 4294967267u,
                                  4294967289u, 4294967295u};                                           // sigma 4.0556 (x 512 + uniform = 2081.7)
struct CbStream {                       // one (key, domain): cached current block
    u32 key[8]; u32 domain; u64 cur = ~0ull; u32 w[16];
    CbStream(const u8 seed[32], u32 dom) : domain(dom) { seed_to_key(seed, key); }
    const u32* block(u64 b) { if (b != cur) { chacha_block(key, b, (u64)domain, 12, w); cur = b; } return w; }
    u64 draw64(u64 e) { const u32* x = block(e >> 3); return (u64)x[2 * (e & 7)] | ((u64)x[2 * (e & 7) + 1] << 32); }
    u32 uniform_q1(u64 e) {
        const u32* x = block(e >> 2) + 4 * (e & 3);
        for (int c = 0; c < 3; ++c) { u32 v = x[c] & ((1u << 27) - 1); if (v < Q1) return v; }
        u32 v = x[3] & ((1u << 27) - 1); return v >= Q1 ? v - Q1 : v;
    }
    u64 uniform_q2(u64 e) {
        const u32* x = block(e >> 2) + 4 * (e & 3);
        const u64 c0 = ((u64)x[0] | ((u64)x[1] << 32)) & ((1ull << 50) - 1), c1 = ((u64)x[2] | ((u64)x[3] << 32)) & ((1ull << 50) - 1);
        const u64 v = c0 < Q2 ? c0 : c1; return v >= Q2 ? v - Q2 : v;
    }
};
template <int LEN> inline i32 cdt_gauss(const u32 (&cdt)[LEN], u64 h) {
    const u32 u = (u32)h; i32 m = 0;
    for (int k = 0; k < LEN; ++k) m += u >= cdt[k];
    return (h >> 63) ? -m : m;
}
// one RLWE row of a key array: a uniform (NTT order), e Gaussian (coefficient order) -> NTT, b = a z + e + msg
template <class T, int LEN>
inline void rlwe_row_cb(const NttTable<T>& tab, const T* z_ntt, const T* msg_ntt, const u32 (&cdt)[LEN], CbStream& sa, CbStream& se, u64 row,
                        T* a_out, T* b_out) {
    const int n = tab.n; const T q = tab.q;
    std::vector<T> e(n);
    for (int i = 0; i < n; ++i) e[i] = signed_to_field<T>(cdt_gauss(cdt, se.draw64(row * n + i)), q);
    tab.forward(e.data());
    for (int i = 0; i < n; ++i) {
        const u64 el = row * n + i;
        T a; if constexpr (sizeof(T) == 4) a = sa.uniform_q1(el); else a = sa.uniform_q2(el);
        a_out[i] = a;
        b_out[i] = addmod(addmod(mulmod(a, z_ntt[i], q), e[i], q), msg_ntt[i], q);
    }
}
template <class T, int LEN>
inline void rgsw_rows_cb(const NttTable<T>& tab, const T* z_ntt, int m, int logb, int levels, int drop, const u32 (&cdt)[LEN],
                         CbStream& sa, CbStream& se, u64 row0, T* out /*[2*levels][2][n]*/) {
    const int n = tab.n; const T q = tab.q;
    std::vector<T> msg(n);
    for (int row = 0; row < 2 * levels; ++row) {
        const T gm = m ? powmod<T>(2, (u64)(drop + logb * (row % levels)), q) : 0;
        for (int i = 0; i < n; ++i) msg[i] = row < levels ? submod((T)0, mulmod(z_ntt[i], gm, q), q) : gm;
        rlwe_row_cb<T, LEN>(tab, z_ntt, msg.data(), cdt, sa, se, row0 + row, out + ((size_t)row * 2) * n, out + ((size_t)row * 2 + 1) * n);
    }
}
inline DetectionKey gen_detection_key_cb(const SecretKeyPack& sk, const u8 seed[32]) {
    const Tables& tb = tables();
    DetectionKey dk;
    std::vector<u32> z1n(N1);
    for (int i = 0; i < N1; ++i) z1n[i] = signed_to_field<u32>(sk.z1[i], Q1);
    tb.t1.forward(z1n.data());
    dk.bsk1.resize(BSK1_ELEMS);
    { CbStream sa(seed, 16), se(seed, 17);
      for (int i = 0; i < CLUE_N; ++i)
          rgsw_rows_cb<u32, 21>(tb.t1, z1n.data(), sk.s0[i], BS1_LOGB, BS1_LEVELS, BS1_DROP, KG_CDT_L1, sa, se, (u64)i * BSK1_ROWS,
                                dk.bsk1.data() + (size_t)i * BSK1_ROWS * 2 * N1); }
    dk.ksk.resize(KSK_ELEMS);
    { CbStream sa(seed, 18), se(seed, 19);
      for (int i = 0; i < N1; ++i) for (int j = 0; j < KS_LEVELS; ++j) {
          const u64 r = (u64)i * KS_LEVELS + j;
          u32* row = dk.ksk.data() + r * KSK_STRIDE;
          u64 dot = 0;
          for (int k = 0; k < LWE2_N; ++k) { row[k] = sa.uniform_q1(r * 672 + k); if (sk.s2[k]) dot += row[k]; }
          const u64 h = se.draw64(r);
          const i64 e = (i64)512 * cdt_gauss(KG_CDT_KS, h) + (i64)((h >> 32) & 511) - 256;
          const u32 m = mulmod(signed_to_field<u32>(sk.z1[i], Q1), powmod<u32>(2, j, Q1), Q1);
          row[LWE2_N] = addmod(addmod((u32)(dot % Q1), signed_to_field<u32>(e, Q1), Q1), m, Q1);
      } }
    dk.bsk2.resize(BSK2_ELEMS);
    { CbStream sa(seed, 20), se(seed, 21);
      for (int i = 0; i < LWE2_N; ++i)
          rgsw_rows_cb<u64, 3>(tb.t2, sk.z2_ntt.data(), sk.s2[i], BS2_LOGB, BS2_LEVELS, BS2_DROP, KG_CDT_L2, sa, se, (u64)i * BSK2_ROWS,
                               dk.bsk2.data() + (size_t)i * BSK2_ROWS * 2 * N2); }
    dk.trk.resize(TRK_ELEMS);
    { CbStream sa(seed, 22), se(seed, 23);
      std::vector<u64> zc(N2), zs(N2), msg(N2);
      for (int i = 0; i < N2; ++i) zc[i] = signed_to_field<u64>(sk.z2[i], Q2);
      for (int t = 0; t < TR_STEPS; ++t) {
          const int d = (1 << (TR_STEPS - t)) + 1;
          automorphism<u64>(zc.data(), N2, d, Q2, zs.data());
          tb.t2.forward(zs.data());
          for (int j = 0; j < TR_LEVELS; ++j) {
              const u64 g = powmod<u64>(2, (u64)(TR_DROP + TR_LOGB * j), Q2);
              for (int i = 0; i < N2; ++i) msg[i] = submod((u64)0, mulmod(zs[i], g, Q2), Q2);
              u64* row = dk.trk.data() + ((size_t)t * TR_LEVELS + j) * 2 * N2;
              rlwe_row_cb<u64, 3>(tb.t2, sk.z2_ntt.data(), msg.data(), KG_CDT_L2, sa, se, (u64)t * TR_LEVELS + j, row, row + N2);
          }
      } }
    return dk;
}

// ---------------------------------------------------------------------------
// The hot path.
// ---------------------------------------------------------------------------

// a3. CmLwe::extract_all — detector.rs:505-531 (modulus switch skipped: 2048 == 2*N1, :519-529)
inline void extract_clues(const u16* a, const u16* b, u16* out_a /*[7][512]*/, u16* out_b /*[7]*/) {
    for (int c = 0; c < CLUE_COUNT; ++c) {
        for (int j = 0; j < CLUE_N; ++j)
            out_a[c * CLUE_N + j] = j <= c ? a[c - j] : (u16)((CLUE_Q - a[CLUE_N + c - j]) & (CLUE_Q - 1));
        out_b[c] = b[c];
    }
}

// p * X^k, k in [0,2N) — SURVEY A.3
template <class T> inline void monomial_mul(const T* p, int n, unsigned k, T q, T* out) {
    for (int i = 0; i < n; ++i) {
        unsigned pos = (i + k) % (2u * n);
        T v = p[i];
        if (pos < (unsigned)n) out[pos] = v; else out[pos - n] = v ? q - v : 0;
    }
}

// One CMux step: acc += ((X^a - 1) * acc) [x] RGSW.  SURVEY A.5 step 2.  Plain form (the specification; the
// production form below is checked against it in tests/test_oracle_primitives.py).
template <class T>
inline void cmux_step_simple(const NttTable<T>& tab, T* acc_a, T* acc_b, unsigned a, const T* rgsw /*[2L][2][n]*/,
                      int logb, int levels, int drop) {
    const int n = tab.n; const T q = tab.q;
    using W = typename Wide<T>::type;
    if (a == 0) return;                                     // (X^0 - 1) = 0: bit-identical to not skipping
    std::vector<T> ta(n), tb_(n), dig(n);
    monomial_mul(acc_a, n, a, q, ta.data()); monomial_mul(acc_b, n, a, q, tb_.data());
    for (int i = 0; i < n; ++i) { ta[i] = submod(ta[i], acc_a[i], q); tb_[i] = submod(tb_[i], acc_b[i], q); }
    std::vector<W> sa(n, 0), sb(n, 0);
    std::vector<i64> d(levels);
    std::vector<std::vector<T>> digs(levels, std::vector<T>(n));
    for (int half = 0; half < 2; ++half) {
        const T* src = half ? tb_.data() : ta.data();
        for (int i = 0; i < n; ++i) {
            gadget_decompose<T>(src[i], q, logb, levels, drop, d.data());
            for (int j = 0; j < levels; ++j) digs[j][i] = signed_to_field<T>(d[j], q);
        }
        for (int j = 0; j < levels; ++j) {
            tab.forward(digs[j].data());
            const T* ka = rgsw + ((size_t)(half * levels + j) * 2 + 0) * n;
            const T* kb = ka + n;
            const T* x = digs[j].data();
            for (int i = 0; i < n; ++i) { sa[i] += (W)x[i] * ka[i]; sb[i] += (W)x[i] * kb[i]; }
        }
    }
    for (int i = 0; i < n; ++i) { ta[i] = (T)(sa[i] % q); tb_[i] = (T)(sb[i] % q); }
    tab.inverse(ta.data()); tab.inverse(tb_.data());
    for (int i = 0; i < n; ++i) { acc_a[i] = addmod(acc_a[i], ta[i], q); acc_b[i] = addmod(acc_b[i], tb_[i], q); }
}

// Production form of the same step for the CPU baseline: no allocation, offset-word digit extraction (one centred word
// per coefficient, digits by shift/mask, top digit absorbs the remainder — identical digits to gadget_decompose), lazy
// forward transforms (< 4q into the multiply-accumulate) and the special-form reduction for q2.
template <class T> struct CmuxWs { std::vector<T> ta, tb, dig; std::vector<typename Wide<T>::type> sa, sb; std::vector<i64> wa, wb;
    void ensure(int n) { if ((int)ta.size() != n) { ta.resize(n); tb.resize(n); dig.resize(n); sa.resize(n); sb.resize(n); wa.resize(n); wb.resize(n); } } };
template <class T> inline T reduce_acc(typename Wide<T>::type v, T q);
template <> inline u32 reduce_acc<u32>(u64 v, u32 q) { return (u32)(v % q); }
template <> inline u64 reduce_acc<u64>(u128 v, u64) { return reduce128_q2(v); }
template <class T>
inline void cmux_step(const NttTable<T>& tab, T* acc_a, T* acc_b, unsigned a, const T* rgsw /*[2L][2][n]*/,
                      int logb, int levels, int drop) {
    const int n = tab.n; const T q = tab.q;
    using W = typename Wide<T>::type;
    if (a == 0) return;                                     // (X^0 - 1) = 0: bit-identical to not skipping
    static thread_local CmuxWs<T> ws; ws.ensure(n);
    monomial_mul(acc_a, n, a, q, ws.ta.data()); monomial_mul(acc_b, n, a, q, ws.tb.data());
    i64 off = 0;
    for (int j = 0; j < levels - 1; ++j) off += ((i64)1 << (logb - 1)) << (logb * j);
    const i64 half_q = (i64)(q >> 1), rnd = drop > 0 ? (i64)1 << (drop - 1) : 0;
    for (int i = 0; i < n; ++i) {
        i64 va = (i64)submod(ws.ta[i], acc_a[i], q), vb = (i64)submod(ws.tb[i], acc_b[i], q);
        if (va > half_q) va -= (i64)q;
        if (vb > half_q) vb -= (i64)q;
        ws.wa[i] = ((va + rnd) >> drop) + off; ws.wb[i] = ((vb + rnd) >> drop) + off;
        ws.sa[i] = 0; ws.sb[i] = 0;
    }
    const i64 mask = ((i64)1 << logb) - 1, hb = (i64)1 << (logb - 1);
    for (int half = 0; half < 2; ++half) {
        const i64* wsrc = half ? ws.wb.data() : ws.wa.data();
        for (int j = 0; j < levels; ++j) {
            T* x = ws.dig.data();
            if (j < levels - 1) for (int i = 0; i < n; ++i) { i64 d = ((wsrc[i] >> (logb * j)) & mask) - hb; x[i] = (T)(d < 0 ? (i64)q + d : d); }
            else for (int i = 0; i < n; ++i) { i64 d = wsrc[i] >> (logb * (levels - 1)); x[i] = (T)(d < 0 ? (i64)q + d : d); }
            tab.forward_lazy(x);
            const T* ka = rgsw + ((size_t)(half * levels + j) * 2 + 0) * n;
            const T* kb = ka + n;
            W* sa = ws.sa.data(); W* sb = ws.sb.data();
            for (int i = 0; i < n; ++i) { sa[i] += (W)x[i] * ka[i]; sb[i] += (W)x[i] * kb[i]; }
        }
    }
    for (int i = 0; i < n; ++i) { ws.ta[i] = reduce_acc<T>(ws.sa[i], q); ws.tb[i] = reduce_acc<T>(ws.sb[i], q); }
    tab.inverse(ws.ta.data()); tab.inverse(ws.tb.data());
    for (int i = 0; i < n; ++i) { acc_a[i] = addmod(acc_a[i], ws.ta[i], q); acc_b[i] = addmod(acc_b[i], ws.tb[i], q); }
}

// BlindRotationKey::blind_rotate [UPSTREAM] — detector.rs:555, 623.  acc = (0, LUT * X^(2N - b)).
template <class T, class A>
inline void blind_rotate(const NttTable<T>& tab, const T* lut, const A* lwe_a, int lwe_n, unsigned lwe_b,
                         const T* bsk, int logb, int levels, int drop, T* acc_a, T* acc_b) {
    const int n = tab.n;
    std::fill(acc_a, acc_a + n, (T)0);
    monomial_mul(lut, n, (2u * n - lwe_b) % (2u * n), tab.q, acc_b);
    for (int i = 0; i < lwe_n; ++i)
        cmux_step<T>(tab, acc_a, acc_b, (unsigned)lwe_a[i], bsk + (size_t)i * 2 * levels * 2 * n, logb, levels, drop);
}

// a4. first_level blind rotations + sum — detector.rs:553-557.  out = RLWE (a[1024], b[1024]) mod q1.
inline void l1_blind_rotate_sum(const DetectionKey& dk, const u16* clue_a, const u16* clue_b, u32* out_a, u32* out_b) {
    const Tables& tb = tables();
    std::vector<u16> ea(CLUE_COUNT * CLUE_N), eb(CLUE_COUNT);
    extract_clues(clue_a, clue_b, ea.data(), eb.data());
    std::fill(out_a, out_a + N1, 0u); std::fill(out_b, out_b + N1, 0u);
    std::vector<u32> aa(N1), ab(N1);
    for (int c = 0; c < CLUE_COUNT; ++c) {
        blind_rotate<u32, u16>(tb.t1, tb.lut1.data(), ea.data() + c * CLUE_N, CLUE_N, eb[c], dk.bsk1.data(),
                               BS1_LOGB, BS1_LEVELS, BS1_DROP, aa.data(), ab.data());
        for (int i = 0; i < N1; ++i) { out_a[i] = addmod(out_a[i], aa[i], Q1); out_b[i] = addmod(out_b[i], ab[i], Q1); }
    }
}

// a5+a6. extract_lwe_locally + key_switch + lwe_modulus_switch + offset — detector.rs:560-596.
//   out: 671 values in [0,4096): a[670], b.
inline void keyswitch_modswitch(const DetectionKey& dk, const u32* rl_a, const u32* rl_b, u32* out /*[671]*/) {
    std::vector<u64> acc(KSK_STRIDE, 0);     // accumulates SUM d*KSK as (pos - neg) tracked mod q1
    std::vector<u64> neg(KSK_STRIDE, 0);
    i64 d[KS_LEVELS];
    for (int i = 0; i < N1; ++i) {
        // sample extraction of the constant term: a' = (a0, -a_{N-1}, ..., -a_1)   SURVEY A.3
        u32 ai = i == 0 ? rl_a[0] : (rl_a[N1 - i] ? Q1 - rl_a[N1 - i] : 0);
        gadget_decompose<u32>(ai, Q1, KS_LOGB, KS_LEVELS, 0, d);
        for (int j = 0; j < KS_LEVELS; ++j) {
            if (!d[j]) continue;
            const u32* row = dk.ksk.data() + ((size_t)i * KS_LEVELS + j) * KSK_STRIDE;
            if (d[j] == 1) for (size_t k = 0; k < KSK_STRIDE; ++k) acc[k] += row[k];
            else if (d[j] == -1) for (size_t k = 0; k < KSK_STRIDE; ++k) neg[k] += row[k];
            else throw std::runtime_error("ks digit out of range");
        }
    }
    for (size_t k = 0; k < KSK_STRIDE; ++k) {
        u32 s = submod((u32)(acc[k] % Q1), (u32)(neg[k] % Q1), Q1);      // SUM d*KSK
        u32 base = k == (size_t)LWE2_N ? rl_b[0] : 0;                    // (0,...,0,b)
        u32 x = submod(base, s, Q1);
        // modulus switch q1 -> 4096, round half up (no ties: q1 odd)   SURVEY A.5 step 6
        u32 y = (u32)(((u64)2 * LWE2_Q * x + Q1) / (2ull * Q1)) & (LWE2_Q - 1);
        out[k] = y;
    }
    // b += clue_count * (4096 >> 5)                                    detector.rs:577-594
    out[LWE2_N] = (out[LWE2_N] + CLUE_COUNT * (LWE2_Q >> 5)) & (LWE2_Q - 1);
}

// a7. second-level blind rotation — detector.rs:599-624 (mod switch skipped: 4096 == 2*N2)
inline void l2_blind_rotate(const DetectionKey& dk, const u32* lwe /*[671]*/, u64* out_a, u64* out_b) {
    const Tables& tb = tables();
    blind_rotate<u64, u32>(tb.t2, tb.lut2.data(), lwe, LWE2_N, lwe[LWE2_N], dk.bsk2.data(),
                           BS2_LOGB, BS2_LEVELS, BS2_DROP, out_a, out_b);
}

// a8. hom_trace — detector.rs:626-639: (a,b) *= N^-1; for k=11..1: c += KS_k(sigma_{2^k+1}(c)); to NTT.
inline void trace_to_ntt(const DetectionKey& dk, u64* a, u64* b) {
    const Tables& tb = tables();
    for (int i = 0; i < N2; ++i) { a[i] = mulmod(a[i], tb.n2_inv, Q2); b[i] = mulmod(b[i], tb.n2_inv, Q2); }
    std::vector<u64> sa(N2), sb(N2), dig(N2);
    std::vector<u128> ra(N2), rb(N2);
    i64 d[TR_LEVELS];
    std::vector<std::vector<u64>> digs(TR_LEVELS, std::vector<u64>(N2));
    for (int t = 0; t < TR_STEPS; ++t) {
        const int k = TR_STEPS - t, deg = (1 << k) + 1;
        automorphism<u64>(a, N2, deg, Q2, sa.data());
        automorphism<u64>(b, N2, deg, Q2, sb.data());
        for (int i = 0; i < N2; ++i) {
            gadget_decompose<u64>(sa[i], Q2, TR_LOGB, TR_LEVELS, TR_DROP, d);
            for (int j = 0; j < TR_LEVELS; ++j) digs[j][i] = signed_to_field<u64>(d[j], Q2);
        }
        std::fill(ra.begin(), ra.end(), (u128)0); std::fill(rb.begin(), rb.end(), (u128)0);
        for (int j = 0; j < TR_LEVELS; ++j) {
            tb.t2.forward(digs[j].data());
            const u64* ka = dk.trk.data() + ((size_t)t * TR_LEVELS + j) * 2 * N2; const u64* kb = ka + N2;
            for (int i = 0; i < N2; ++i) { ra[i] += (u128)digs[j][i] * ka[i]; rb[i] += (u128)digs[j][i] * kb[i]; }
        }
        for (int i = 0; i < N2; ++i) { sa[i] = (u64)(ra[i] % Q2); dig[i] = (u64)(rb[i] % Q2); }
        tb.t2.inverse(sa.data()); tb.t2.inverse(dig.data());
        for (int i = 0; i < N2; ++i) {
            a[i] = addmod(a[i], sa[i], Q2);
            b[i] = addmod(b[i], addmod(dig[i], sb[i], Q2), Q2);
        }
    }
    tb.t2.forward(a); tb.t2.forward(b);
}

// a1. Detector::detect — detector.rs:135-166.  pv = [2][2048] u64, NTT domain.
inline void detect(const DetectionKey& dk, const u16* clue_a, const u16* clue_b, u64* pv) {
    std::vector<u32> ra(N1), rb(N1), lwe(KSK_STRIDE);
    l1_blind_rotate_sum(dk, clue_a, clue_b, ra.data(), rb.data());
    keyswitch_modswitch(dk, ra.data(), rb.data(), lwe.data());
    l2_blind_rotate(dk, lwe.data(), pv, pv + N2);
    trace_to_ntt(dk, pv, pv + N2);
}

// ---------------------------------------------------------------------------
// Retrieval layout — parameters/retrieval_params.rs:50-106
// ---------------------------------------------------------------------------
struct RetrievalParams {
    u64 index_modulus; int polynomial_size, bucket_count_per_segment, slots_per_bucket, slots_per_segment,
        segment_count, segment_per_cipher, max_encode_indices_cipher_count, pertinent_count, combination_count,
        cmb_count_per_cipher; size_t all_payloads_count;
    RetrievalParams(size_t all_payloads, int pertinent, u64 p = OUT_P, int poly = N2, int buckets = BUCKETS_PER_SEGMENT,
                    int segments = SEGMENT_COUNT, int cmb_per_cipher = CMB_PER_CIPHER) {
        index_modulus = p; polynomial_size = poly; bucket_count_per_segment = buckets; segment_count = segments;
        cmb_count_per_cipher = cmb_per_cipher; all_payloads_count = all_payloads; pertinent_count = pertinent;
        // non-power-of-two branch (p = 257): smallest e >= 1 with p^e >= D            :64-75
        int e = 1; u64 pw = p;
        while (pw < all_payloads) { pw *= p; ++e; }
        slots_per_bucket = e + 1;                                                    // :77
        slots_per_segment = slots_per_bucket * buckets;                              // :78
        segment_per_cipher = poly / slots_per_segment;                               // :80
        max_encode_indices_cipher_count = segments / segment_per_cipher;             // :81
        combination_count = pertinent + 5;                                           // :85-89 (p not pow2)
    }
    int payload_cipher_count() const { return (combination_count + cmb_count_per_cipher - 1) / cmb_count_per_cipher; }
};

// Index-bucket RNG.  The reference draws buckets from rand::thread_rng() (detector.rs:262) and is
// therefore not reproducible run to run; ours is a counter-based hash keyed by
// (seed, cipher index, global message index, segment) so results do not depend on batching or GPU
// count.  Shared bit-for-bit with the CUDA kernel.
inline u64 mix64(u64 z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
inline u32 bucket_of(u64 seed, u32 cipher_idx, u64 msg, u32 seg, u32 buckets) {
    u64 h = mix64(seed + 0x9E3779B97F4A7C15ull * (msg + 1));
    h = mix64(h ^ (((u64)cipher_idx << 32) | seg) * 0xD1342543DE82EF95ull);
    return (u32)(((h >> 32) * buckets) >> 32);
}
inline u64 centred_p(u64 v) { return v < ((OUT_P + 1) >> 1) ? v : Q2 - OUT_P + v; }   // detector.rs:254,309

// a9. encode_pertinent_indices — detector.rs:223-339.  pv: [count][2][2048]; global index of pv[0] = index0.
inline void encode_indices(const RetrievalParams& rp, const u64* pv, size_t count, u64 index0, u64 seed,
                           u32 cipher_idx, u64* out /*[2][2048], accumulated into (caller zeroes)*/) {
    const Tables& tb = tables();
    std::vector<u64> poly(N2);
    for (size_t m = 0; m < count; ++m) {
        std::fill(poly.begin(), poly.end(), 0ull);
        const u64 gi = index0 + m;
        for (int s = 0; s < rp.segment_per_cipher; ++s) {
            u64* chunk = poly.data() + (size_t)s * rp.slots_per_segment;
            const u32 bucket = bucket_of(seed, cipher_idx, gi, (u32)s, (u32)rp.bucket_count_per_segment);
            const size_t address = (size_t)bucket * rp.slots_per_bucket;
            u64 i = gi; int k = 0;
            while (i != 0) { u64 v = i % rp.index_modulus; chunk[address + k] = centred_p(v); i = (i - v) / rp.index_modulus; ++k; }
            chunk[address + rp.slots_per_bucket - 1] = 1;
        }
        tb.t2.forward(poly.data());
        const u64* pa = pv + m * 2 * N2; const u64* pb = pa + N2;
        for (int i = 0; i < N2; ++i) {
            out[i] = addmod(out[i], mulmod(pa[i], poly[i], Q2), Q2);
            out[N2 + i] = addmod(out[N2 + i], mulmod(pb[i], poly[i], Q2), Q2);
        }
    }
}

// a10. encode_pertinent_payloads — detector.rs:341-453.  weights: [rows][D] u16 row-major (rows >= 2*ncipher),
//   payloads: [count][612] u16, message m has global index index0+m (weight column).
inline void encode_payloads(const u64* pv, const u16* payloads, size_t count, u64 index0, const u16* weights,
                            size_t weight_stride, int n_cipher, int cmb_per_cipher, u64* out /*[n_cipher][2][2048] accumulated*/) {
    const Tables& tb = tables();
    std::vector<u64> poly(N2);
    for (int c = 0; c < n_cipher; ++c) {
        u64* oc = out + (size_t)c * 2 * N2;
        for (size_t m = 0; m < count; ++m) {
            std::fill(poly.begin(), poly.end(), 0ull);
            for (int j = 0; j < cmb_per_cipher; ++j) {
                const u32 w = weights[(size_t)(c * cmb_per_cipher + j) * weight_stride + index0 + m];
                for (int k = 0; k < PAYLOAD_LEN; ++k)
                    poly[(size_t)j * PAYLOAD_LEN + k] = centred_p(((u32)payloads[m * PAYLOAD_LEN + k] * w) % OUT_P);
            }
            tb.t2.forward(poly.data());
            const u64* pa = pv + m * 2 * N2; const u64* pb = pa + N2;
            for (int i = 0; i < N2; ++i) {
                oc[i] = addmod(oc[i], mulmod(pa[i], poly[i], Q2), Q2);
                oc[N2 + i] = addmod(oc[N2 + i], mulmod(pb[i], poly[i], Q2), Q2);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Recipient side (needed because the pinned parity target is the decoded result)
// ---------------------------------------------------------------------------

// decrypt + decode an NTT-domain RLWE: b - a*z, INTT, round_half_up(c*p/q), fold t>=p -> t-p.
//   retriever.rs:79-90, 339-355; omd.rs:25,48.
inline void decrypt_decode(const SecretKeyPack& sk, const u64* ct /*[2][2048]*/, u64* out /*[2048] values mod p*/) {
    const Tables& tb = tables();
    std::vector<u64> d(N2);
    for (int i = 0; i < N2; ++i) d[i] = submod(ct[N2 + i], mulmod(ct[i], sk.z2_ntt[i], Q2), Q2);
    tb.t2.inverse(d.data());
    for (int i = 0; i < N2; ++i) {
        u64 t = (u64)(((u128)2 * OUT_P * d[i] + Q2) / ((u128)2 * Q2));
        if (t >= OUT_P) t -= OUT_P;
        out[i] = t;
    }
}
inline void decrypt_raw(const SecretKeyPack& sk, const u64* ct, u64* out /*[2048] mod q2*/) {
    const Tables& tb = tables();
    for (int i = 0; i < N2; ++i) out[i] = submod(ct[N2 + i], mulmod(ct[i], sk.z2_ntt[i], Q2), Q2);
    tb.t2.inverse(out);
}

// Retriever::decode_pertinent_indices — retriever.rs:63-130
inline void decode_indices(const SecretKeyPack& sk, const RetrievalParams& rp, const u64* ct, std::set<size_t>& set) {
    std::vector<u64> dec(N2);
    decrypt_decode(sk, ct, dec.data());
    for (int s = 0; s < rp.segment_per_cipher; ++s) for (int bkt = 0; bkt < rp.bucket_count_per_segment; ++bkt) {
        const u64* bucket = dec.data() + (size_t)s * rp.slots_per_segment + (size_t)bkt * rp.slots_per_bucket;
        if (bucket[rp.slots_per_bucket - 1] == 1) {
            u64 idx = 0;
            for (int k = rp.slots_per_bucket - 2; k >= 0; --k) idx = idx * rp.index_modulus + bucket[k];
            set.insert((size_t)idx);
        }
    }
}

// the only golden table of the reference — matrix.rs:28-41 (checked in tests against Fermat inverses)
static const u16 INV_MOD_257[257] = {
    0, 1, 129, 86, 193, 103, 43, 147, 225, 200, 180, 187, 150, 178, 202, 120, 241, 121, 100, 230,
    90, 49, 222, 190, 75, 72, 89, 238, 101, 195, 60, 199, 249, 148, 189, 235, 50, 132, 115, 145,
    45, 163, 153, 6, 111, 40, 95, 175, 166, 21, 36, 126, 173, 97, 119, 243, 179, 248, 226, 61, 30,
    59, 228, 102, 253, 87, 74, 234, 223, 149, 246, 181, 25, 169, 66, 24, 186, 247, 201, 244, 151,
    165, 210, 96, 205, 127, 3, 65, 184, 26, 20, 209, 176, 152, 216, 46, 83, 53, 139, 135, 18, 28,
    63, 5, 215, 164, 177, 245, 188, 224, 250, 44, 218, 116, 124, 38, 113, 134, 159, 54, 15, 17,
    158, 140, 114, 220, 51, 85, 255, 2, 172, 206, 37, 143, 117, 99, 240, 242, 203, 98, 123, 144,
    219, 133, 141, 39, 213, 7, 33, 69, 12, 80, 93, 42, 252, 194, 229, 239, 122, 118, 204, 174, 211,
    41, 105, 81, 48, 237, 231, 73, 192, 254, 130, 52, 161, 47, 92, 106, 13, 56, 10, 71, 233, 191,
    88, 232, 76, 11, 108, 34, 23, 183, 170, 4, 155, 29, 198, 227, 196, 31, 9, 78, 14, 138, 160, 84,
    131, 221, 236, 91, 82, 162, 217, 146, 251, 104, 94, 212, 112, 142, 125, 207, 22, 68, 109, 8,
    58, 197, 62, 156, 19, 168, 185, 182, 67, 35, 208, 167, 27, 157, 136, 16, 137, 55, 79, 107, 70,
    77, 57, 32, 110, 214, 154, 64, 171, 128, 256,
};

// solve_matrix_mod_257 — matrix.rs:164-247.  matrix [rows][cols], payloads [rows][612]; returns false if singular.
inline bool solve_matrix_mod_257(std::vector<std::vector<u16>>& mat, std::vector<std::array<u16, PAYLOAD_LEN>>& pl,
                                 std::vector<std::array<u16, PAYLOAD_LEN>>& out) {
    const u32 P = 257; const size_t rows = mat.size(), cols = mat[0].size();
    if (rows < cols) return false;
    auto mul_row = [&](std::array<u16, PAYLOAD_LEN>& r, u32 c) { for (auto& v : r) v = (u16)(v * c % P); };
    auto sub_mul = [&](std::array<u16, PAYLOAD_LEN>& dst, const std::array<u16, PAYLOAD_LEN>& src, u32 c) {
        for (int k = 0; k < PAYLOAD_LEN; ++k) { u32 t = src[k] * c % P; dst[k] = (u16)((dst[k] + P - t) % P); } };
    for (size_t i = 0; i < cols; ++i) {
        size_t pick = rows;
        for (size_t j = i; j < rows; ++j) if (mat[j][i] != 0) { pick = j; break; }
        if (pick == rows) return false;                     // OmrError::InvertibleMatrix (error.rs:4-8)
        if (pick != i) { std::swap(mat[i], mat[pick]); std::swap(pl[i], pl[pick]); }
        u32 v = mat[i][i];
        if (v != 1) {
            u32 inv = INV_MOD_257[v];
            mat[i][i] = 1;
            for (size_t c = i + 1; c < cols; ++c) mat[i][c] = (u16)(mat[i][c] * inv % P);
            mul_row(pl[i], inv);
        }
        if (i == cols - 1) break;
        for (size_t r = i + 1; r < rows; ++r) {
            u32 c = mat[r][i];
            if (c) {
                for (size_t cc = i; cc < cols; ++cc) { u32 t = mat[i][cc] * c % P; mat[r][cc] = (u16)((mat[r][cc] + P - t) % P); }
                sub_mul(pl[r], pl[i], c);
            }
        }
    }
    for (size_t ic = cols - 1; ic >= 1; --ic) for (size_t ir = 0; ir < ic; ++ir) {
        u32 c = mat[ir][ic];
        if (c) { sub_mul(pl[ir], pl[ic], c); mat[ir][ic] = 0; }
    }
    out.assign(pl.begin(), pl.begin() + cols);
    return true;
}

// Retriever::decode_digest — retriever.rs:188-260.  weights [combination_count][D] row-major (explicit; the
// reference regenerates them from the 32-byte seed with StdRng+Uniform, retriever.rs:215-226 — see chacha12_weights).
// returns 0 ok, 1 = singular matrix.
inline int decode_digest(const SecretKeyPack& sk, const RetrievalParams& rp, const u64* index_cts, int n_index_cts,
                         const u64* payload_cts, const u16* weights, size_t weight_stride,
                         std::vector<size_t>& indices, std::vector<std::array<u16, PAYLOAD_LEN>>& solved) {
    std::set<size_t> set;
    for (int c = 0; c < n_index_cts; ++c) {
        decode_indices(sk, rp, index_cts + (size_t)c * 2 * N2, set);
        if ((int)set.size() == rp.pertinent_count) break;                            // :200-204, :125-129
    }
    indices.assign(set.begin(), set.end());                                          // sorted (:207-211)
    const size_t pc = indices.size();
    if (pc == 0) { solved.clear(); return 0; }
    std::vector<std::vector<u16>> mat(rp.combination_count, std::vector<u16>(pc));
    for (int r = 0; r < rp.combination_count; ++r) for (size_t k = 0; k < pc; ++k) {
        size_t col = indices[k];
        mat[r][k] = col < rp.all_payloads_count ? weights[(size_t)r * weight_stride + col] : 0;
    }
    // decode_combined_payloads — retriever.rs:318-362
    std::vector<std::array<u16, PAYLOAD_LEN>> comb(rp.combination_count);
    std::vector<u64> dec(N2);
    for (int c = 0; c < rp.payload_cipher_count(); ++c) {
        decrypt_decode(sk, payload_cts + (size_t)c * 2 * N2, dec.data());
        for (int j = 0; j < rp.cmb_count_per_cipher; ++j) {
            int r = c * rp.cmb_count_per_cipher + j;
            if (r >= rp.combination_count) break;
            for (int k = 0; k < PAYLOAD_LEN; ++k) comb[r][k] = (u16)dec[(size_t)j * PAYLOAD_LEN + k];
        }
    }
    if (!solve_matrix_mod_257(mat, comb, solved)) return 1;
    return 0;
}

// ---------------------------------------------------------------------------
// Weight stream of the reference: StdRng::from_seed(seed) (rand 0.8 => ChaCha12, 64-bit counter, stream 0)
// + Uniform::<u16>::new(0,257).sample (widening-multiply rejection on one u32 per draw) — detector.rs:376-387,
// retriever.rs:215-226 [UPSTREAM rand 0.8 / rand_chacha 0.3].  Restated from the published algorithms; parity
// with the Rust crates is unpinned (no Rust here); the ChaCha core is checked against the published zero-key
// blocks of ChaCha20 (RFC 7539), ChaCha12 and ChaCha8 in tests.
// ---------------------------------------------------------------------------
inline void chacha_block(const u32 key[8], u64 counter, u64 stream, int rounds, u32 out[16]) {
    u32 st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5],
                  key[6], key[7], (u32)counter, (u32)(counter >> 32), (u32)stream, (u32)(stream >> 32)};
    u32 x[16]; std::memcpy(x, st, sizeof x);
    auto rotl = [](u32 v, int c) { return (v << c) | (v >> (32 - c)); };
    auto qr = [&](int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7); };
    for (int r = 0; r < rounds; r += 2) {
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
    }
    for (int i = 0; i < 16; ++i) out[i] = x[i] + st[i];
}
inline void chacha12_weights(const u8 seed[32], u16* out, size_t count, u32 range = 257) {
    u32 key[8]; for (int i = 0; i < 8; ++i) key[i] = (u32)seed[4 * i] | ((u32)seed[4 * i + 1] << 8) | ((u32)seed[4 * i + 2] << 16) | ((u32)seed[4 * i + 3] << 24);
    const u32 ints_to_reject = (u32)((0xFFFFFFFFull - range + 1) % range);
    const u32 zone = 0xFFFFFFFFu - ints_to_reject;
    u32 buf[16]; int pos = 16; u64 ctr = 0; size_t n = 0;
    while (n < count) {
        if (pos == 16) { chacha_block(key, ctr++, 0, 12, buf); pos = 0; }
        u64 m = (u64)buf[pos++] * range;
        if ((u32)m <= zone) out[n++] = (u16)(m >> 32);
    }
}

}  // namespace orc

// extern "C" surface of the CPU oracle (ctypes-loadable).  TEST INFRASTRUCTURE ONLY — see omr_oracle.hpp.
// PARITY UNPINNED against Primus-fhe (see header of omr_oracle.hpp).
#include "omr_oracle.hpp"
#include <cstdio>
#include <thread>
#include <atomic>
#include <string>

// tiny parallel-for (this image's gcc has no libgomp)
template <class F> static void parallel_for(size_t n, int threads, F f) {
    if (threads <= 1 || n <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) pool.emplace_back([&] { for (;;) { size_t i = next.fetch_add(1); if (i >= n) return; f(i); } });
    for (auto& th : pool) th.join();
}

using namespace orc;

struct OrcHandle {
    SecretKeyPack sk; ClueKey ck; DetectionKey dk; bool has_secret = false;
};

extern "C" {

// ---- constants / tables ----------------------------------------------------
uint64_t orc_const(const char* name) {
    std::string s(name);
    if (s == "Q1") return Q1; if (s == "Q2") return Q2; if (s == "N1") return N1; if (s == "N2") return N2;
    if (s == "PSI1") return tables().t1.psi; if (s == "PSI2") return tables().t2.psi;
    if (s == "N2_INV") return tables().n2_inv;
    if (s == "BSK1_ELEMS") return BSK1_ELEMS; if (s == "KSK_ELEMS") return KSK_ELEMS;
    if (s == "BSK2_ELEMS") return BSK2_ELEMS; if (s == "TRK_ELEMS") return TRK_ELEMS;
    if (s == "LUT1_SCALE") return tables().lut1[0];
    if (s == "LUT2_SCALE") return tables().lut2[1728];
    return ~0ull;
}
void orc_lut1(uint32_t* out) { std::memcpy(out, tables().lut1.data(), N1 * 4); }
void orc_lut2(uint64_t* out) { std::memcpy(out, tables().lut2.data(), N2 * 8); }
void orc_twiddles1(uint32_t* fwd, uint32_t* inv) { std::memcpy(fwd, tables().t1.tw.data(), N1 * 4); std::memcpy(inv, tables().t1.itw.data(), N1 * 4); }
void orc_twiddles2(uint64_t* fwd, uint64_t* inv) { std::memcpy(fwd, tables().t2.tw.data(), N2 * 8); std::memcpy(inv, tables().t2.itw.data(), N2 * 8); }
uint16_t orc_inv_mod_257(int i) { return INV_MOD_257[i]; }

// ---- primitives --------------------------------------------------------------
void orc_ntt1_forward(uint32_t* a, size_t batch) { for (size_t i = 0; i < batch; ++i) tables().t1.forward(a + i * N1); }
void orc_ntt1_inverse(uint32_t* a, size_t batch) { for (size_t i = 0; i < batch; ++i) tables().t1.inverse(a + i * N1); }
void orc_ntt2_forward(uint64_t* a, size_t batch) { for (size_t i = 0; i < batch; ++i) tables().t2.forward(a + i * N2); }
void orc_ntt2_inverse(uint64_t* a, size_t batch) { for (size_t i = 0; i < batch; ++i) tables().t2.inverse(a + i * N2); }
void orc_negacyclic1(const uint32_t* a, const uint32_t* b, uint32_t* c) { auto r = negacyclic_schoolbook<u32>(a, b, N1, Q1); std::memcpy(c, r.data(), N1 * 4); }
void orc_negacyclic2(const uint64_t* a, const uint64_t* b, uint64_t* c) { auto r = negacyclic_schoolbook<u64>(a, b, N2, Q2); std::memcpy(c, r.data(), N2 * 8); }
// which: 0 = BSK1 basis (q1,5,4), 1 = KS basis (q1,1,27), 2 = BSK2 basis (q2,7,6), 3 = trace basis (q2,2,25)
void orc_decompose(int which, const uint64_t* x, size_t n, int64_t* digits /*[n][levels]*/) {
    for (size_t i = 0; i < n; ++i) {
        switch (which) {
            case 0: gadget_decompose<u32>((u32)x[i], Q1, BS1_LOGB, BS1_LEVELS, BS1_DROP, digits + i * BS1_LEVELS); break;
            case 1: gadget_decompose<u32>((u32)x[i], Q1, KS_LOGB, KS_LEVELS, 0, digits + i * KS_LEVELS); break;
            case 2: gadget_decompose<u64>(x[i], Q2, BS2_LOGB, BS2_LEVELS, BS2_DROP, digits + i * BS2_LEVELS); break;
            default: gadget_decompose<u64>(x[i], Q2, TR_LOGB, TR_LEVELS, TR_DROP, digits + i * TR_LEVELS); break;
        }
    }
}
uint64_t orc_reduce128_q2(uint64_t hi, uint64_t lo) { return reduce128_q2(((u128)hi << 64) | lo); }
uint64_t orc_mod128_q2(uint64_t hi, uint64_t lo) { return (u64)((((u128)hi << 64) | lo) % Q2); }
void orc_chacha_block(const uint32_t* key, uint64_t counter, uint64_t stream, int rounds, uint32_t* out) { chacha_block(key, counter, stream, rounds, out); }
void orc_chacha12_weights(const uint8_t* seed, uint16_t* out, size_t count) { chacha12_weights(seed, out, count); }
uint32_t orc_bucket_of(uint64_t seed, uint32_t cipher_idx, uint64_t msg, uint32_t seg) { return bucket_of(seed, cipher_idx, msg, seg, BUCKETS_PER_SEGMENT); }
// out: slots_per_bucket, slots_per_segment, segment_per_cipher, max_encode_indices_cipher_count, combination_count, payload_cipher_count
void orc_retrieval_params(size_t all_payloads, int pertinent, int* out) {
    RetrievalParams rp(all_payloads, pertinent);
    out[0] = rp.slots_per_bucket; out[1] = rp.slots_per_segment; out[2] = rp.segment_per_cipher;
    out[3] = rp.max_encode_indices_cipher_count; out[4] = rp.combination_count; out[5] = rp.payload_cipher_count();
}

// ---- keys --------------------------------------------------------------------
void* orc_keygen(uint64_t seed) {
    auto* h = new OrcHandle;
    h->sk = gen_secret_key(seed); h->ck = gen_clue_key(h->sk, seed); h->dk = gen_detection_key(h->sk, seed);
    h->has_secret = true;
    return h;
}
// only the clue side of a key pack (the decoy sender of examples/omr.rs:76,81)
void* orc_keygen_sender_only(uint64_t seed) {
    auto* h = new OrcHandle;
    h->sk = gen_secret_key(seed); h->ck = gen_clue_key(h->sk, seed); h->has_secret = true;
    return h;
}
// a detection key from caller-supplied flat blobs (e.g. uniformly random ones for arithmetic parity tests)
void* orc_key_from_blobs(const uint32_t* bsk1, const uint32_t* ksk, const uint64_t* bsk2, const uint64_t* trk) {
    auto* h = new OrcHandle;
    h->dk.bsk1.assign(bsk1, bsk1 + BSK1_ELEMS); h->dk.ksk.assign(ksk, ksk + KSK_ELEMS);
    h->dk.bsk2.assign(bsk2, bsk2 + BSK2_ELEMS); h->dk.trk.assign(trk, trk + TRK_ELEMS);
    return h;
}
// the same secrets (and clue key) with a detection key from the counter-based generator (bit-exact twin of csrc/keygen.cuh)
void* orc_keygen_cb(void* hh, const uint8_t* seed32) {
    auto* src = (OrcHandle*)hh;
    auto* h = new OrcHandle;
    h->sk = src->sk; h->ck = src->ck; h->has_secret = true;
    h->dk = gen_detection_key_cb(h->sk, seed32);
    return h;
}
void orc_free(void* h) { delete (OrcHandle*)h; }
const uint32_t* orc_bsk1(void* h) { return ((OrcHandle*)h)->dk.bsk1.data(); }
const uint32_t* orc_ksk(void* h) { return ((OrcHandle*)h)->dk.ksk.data(); }
const uint64_t* orc_bsk2(void* h) { return ((OrcHandle*)h)->dk.bsk2.data(); }
const uint64_t* orc_trk(void* h) { return ((OrcHandle*)h)->dk.trk.data(); }
void orc_secret(void* hh, int32_t* s0, int32_t* z1, int32_t* s2, int32_t* z2) {
    auto* h = (OrcHandle*)hh;
    if (s0) std::memcpy(s0, h->sk.s0.data(), CLUE_N * 4); if (z1) std::memcpy(z1, h->sk.z1.data(), N1 * 4);
    if (s2) std::memcpy(s2, h->sk.s2.data(), LWE2_N * 4); if (z2) std::memcpy(z2, h->sk.z2.data(), N2 * 4);
}

// ---- clues -------------------------------------------------------------------
void orc_gen_clues(void* hh, uint64_t seed, uint64_t index0, size_t count, uint16_t* a /*[count][512]*/, uint16_t* b /*[count][7]*/, int threads) {
    auto* h = (OrcHandle*)hh;
    parallel_for((size_t)count, threads, [&](size_t i) { gen_clue(h->ck, seed, index0 + i, nullptr, a + i * CLUE_N, b + i * CLUE_COUNT); });
}
// counter-based variant (bit-exact twin of the CUDA clue_gen_kernel)
void orc_gen_clues_cb(void* hh, const uint8_t* seed /*32 bytes*/, uint64_t index0, size_t count, const uint8_t* msgs /*nullable [count][7]*/, uint16_t* a, uint16_t* b, int threads) {
    auto* h = (OrcHandle*)hh;
    parallel_for((size_t)count, threads, [&](size_t i) { gen_clue_cb(h->ck, seed, index0 + i, msgs ? msgs + i * CLUE_COUNT : nullptr, a + i * CLUE_N, b + i * CLUE_COUNT); });
}
void orc_clue_key(void* hh, uint16_t* pa, uint16_t* pb) {
    auto* h = (OrcHandle*)hh;
    std::memcpy(pa, h->ck.pa.data(), CLUE_N * 2); std::memcpy(pb, h->ck.pb.data(), CLUE_N * 2);
}
void orc_gen_clue_msgs(void* hh, uint64_t seed, uint64_t index, const uint32_t* msgs, uint16_t* a, uint16_t* b) {
    gen_clue(((OrcHandle*)hh)->ck, seed, index, msgs, a, b);
}
// decrypt the 7 clue plaintexts (phase / 256 rounded mod 8) with s0 — SecretKeyPack::decrypt_clue secret.rs:266-270
void orc_decrypt_clue(void* hh, const uint16_t* a, const uint16_t* b, uint32_t* out /*7*/) {
    auto* h = (OrcHandle*)hh;
    std::vector<u16> ea(CLUE_COUNT * CLUE_N), eb(CLUE_COUNT);
    extract_clues(a, b, ea.data(), eb.data());
    for (int c = 0; c < CLUE_COUNT; ++c) {
        i64 ph = eb[c];
        for (int j = 0; j < CLUE_N; ++j) ph -= (i64)ea[c * CLUE_N + j] * h->sk.s0[j];
        u32 p = (u32)(ph & (CLUE_Q - 1));
        out[c] = ((p + CLUE_Q / CLUE_T / 2) / (CLUE_Q / CLUE_T)) % CLUE_T;
    }
}

// ---- detect and its stages ---------------------------------------------------
void orc_detect(void* hh, const uint16_t* a, const uint16_t* b, size_t count, uint64_t* pv /*[count][2][2048]*/, int threads) {
    auto* h = (OrcHandle*)hh; tables();
    parallel_for((size_t)count, threads, [&](size_t i) { detect(h->dk, a + i * CLUE_N, b + i * CLUE_COUNT, pv + i * 2 * N2); });
}
void orc_l1(void* hh, const uint16_t* a, const uint16_t* b, size_t count, uint32_t* out /*[count][2][1024]*/, int threads) {
    auto* h = (OrcHandle*)hh; tables();
    parallel_for((size_t)count, threads, [&](size_t i) { l1_blind_rotate_sum(h->dk, a + i * CLUE_N, b + i * CLUE_COUNT, out + i * 2 * N1, out + i * 2 * N1 + N1); });
}
void orc_keyswitch(void* hh, const uint32_t* rlwe /*[count][2][1024]*/, size_t count, uint32_t* out /*[count][671]*/, int threads) {
    auto* h = (OrcHandle*)hh;
    parallel_for((size_t)count, threads, [&](size_t i) { keyswitch_modswitch(h->dk, rlwe + i * 2 * N1, rlwe + i * 2 * N1 + N1, out + i * KSK_STRIDE); });
}
void orc_l2(void* hh, const uint32_t* lwe /*[count][671]*/, size_t count, uint64_t* out /*[count][2][2048]*/, int threads) {
    auto* h = (OrcHandle*)hh; tables();
    parallel_for((size_t)count, threads, [&](size_t i) { l2_blind_rotate(h->dk, lwe + i * KSK_STRIDE, out + i * 2 * N2, out + i * 2 * N2 + N2); });
}
void orc_trace(void* hh, uint64_t* ct /*[count][2][2048] in place*/, size_t count, int threads) {
    auto* h = (OrcHandle*)hh; tables();
    parallel_for((size_t)count, threads, [&](size_t i) { trace_to_ntt(h->dk, ct + i * 2 * N2, ct + i * 2 * N2 + N2); });
}
// single CMux steps for unit-level parity
void orc_cmux1(void* hh, uint32_t* acc /*[2][1024]*/, unsigned a, int key_index) {
    auto* h = (OrcHandle*)hh;
    cmux_step<u32>(tables().t1, acc, acc + N1, a, h->dk.bsk1.data() + (size_t)key_index * BSK1_ROWS * 2 * N1, BS1_LOGB, BS1_LEVELS, BS1_DROP);
}
// the plain (specification) form of the same step, for checking the production form against it
void orc_cmux1_simple(void* hh, uint32_t* acc, unsigned a, int key_index) {
    auto* h = (OrcHandle*)hh;
    cmux_step_simple<u32>(tables().t1, acc, acc + N1, a, h->dk.bsk1.data() + (size_t)key_index * BSK1_ROWS * 2 * N1, BS1_LOGB, BS1_LEVELS, BS1_DROP);
}
void orc_cmux2_simple(void* hh, uint64_t* acc, unsigned a, int key_index) {
    auto* h = (OrcHandle*)hh;
    cmux_step_simple<u64>(tables().t2, acc, acc + N2, a, h->dk.bsk2.data() + (size_t)key_index * BSK2_ROWS * 2 * N2, BS2_LOGB, BS2_LEVELS, BS2_DROP);
}
void orc_cmux2(void* hh, uint64_t* acc /*[2][2048]*/, unsigned a, int key_index) {
    auto* h = (OrcHandle*)hh;
    cmux_step<u64>(tables().t2, acc, acc + N2, a, h->dk.bsk2.data() + (size_t)key_index * BSK2_ROWS * 2 * N2, BS2_LOGB, BS2_LEVELS, BS2_DROP);
}

// ---- digest packing ------------------------------------------------------------
void orc_encode_indices(size_t all_payloads, int pertinent, const uint64_t* pv, size_t count, uint64_t index0, uint64_t seed,
                        uint32_t cipher_idx, uint64_t* out /*[2][2048], zeroed here*/) {
    RetrievalParams rp(all_payloads, pertinent);
    std::fill(out, out + 2 * N2, 0ull);
    encode_indices(rp, pv, count, index0, seed, cipher_idx, out);
}
void orc_encode_payloads(const uint64_t* pv, const uint16_t* payloads, size_t count, uint64_t index0, const uint16_t* weights,
                         size_t weight_stride, int n_cipher, int cmb_per_cipher, uint64_t* out /*zeroed here*/, int threads) {
    std::fill(out, out + (size_t)n_cipher * 2 * N2, 0ull);
    tables();
    parallel_for((size_t)n_cipher, threads, [&](size_t c) { encode_payloads(pv, payloads, count, index0, weights + (size_t)c * cmb_per_cipher * weight_stride, weight_stride, 1, cmb_per_cipher,
                        out + (size_t)c * 2 * N2); });
}

// ---- recipient -----------------------------------------------------------------
void orc_decrypt_decode(void* hh, const uint64_t* ct, uint64_t* out) { decrypt_decode(((OrcHandle*)hh)->sk, ct, out); }
void orc_decrypt_raw(void* hh, const uint64_t* ct, uint64_t* out) { decrypt_raw(((OrcHandle*)hh)->sk, ct, out); }
// returns status (0 ok, 1 singular); n_found written; indices[<=pertinent], payloads[n_found][612]
int orc_decode_digest(void* hh, size_t all_payloads, int pertinent, const uint64_t* index_cts, int n_index_cts, const uint64_t* payload_cts,
                      const uint16_t* weights, size_t weight_stride, uint64_t* indices_out, int* n_found, uint16_t* payloads_out) {
    RetrievalParams rp(all_payloads, pertinent);
    std::vector<size_t> idx; std::vector<std::array<u16, PAYLOAD_LEN>> solved;
    int st = decode_digest(((OrcHandle*)hh)->sk, rp, index_cts, n_index_cts, payload_cts, weights, weight_stride, idx, solved);
    *n_found = (int)idx.size();
    for (size_t i = 0; i < idx.size(); ++i) indices_out[i] = idx[i];
    if (st == 0) for (size_t i = 0; i < solved.size(); ++i) std::memcpy(payloads_out + i * PAYLOAD_LEN, solved[i].data(), PAYLOAD_LEN * 2);
    return st;
}
// phase of an LWE mod q1 under z1 (debug/semantic checks: SURVEY A.9)
uint32_t orc_phase_l1(void* hh, const uint32_t* rlwe /*[2][1024]*/) {
    auto* h = (OrcHandle*)hh;
    i128 ph = rlwe[N1];
    for (int i = 0; i < N1; ++i) {
        u32 ai = i == 0 ? rlwe[0] : (rlwe[N1 - i] ? Q1 - rlwe[N1 - i] : 0);
        ph -= (i128)ai * h->sk.z1[i];
    }
    i64 r = (i64)(ph % (i128)Q1); if (r < 0) r += Q1; return (u32)r;
}
uint32_t orc_phase_lwe2(void* hh, const uint32_t* lwe /*[671] mod 4096*/) {
    auto* h = (OrcHandle*)hh; i64 ph = lwe[LWE2_N];
    for (int i = 0; i < LWE2_N; ++i) ph -= (i64)lwe[i] * h->sk.s2[i];
    return (u32)(ph & (LWE2_Q - 1));
}


// ---- blobs: the oracle's own reader/writer of the OMRB200 container ------------------------------------------
// header: "OMRB200\0" u32 version u32 kind u64 count u64 index0 u64 aux u64 payload_bytes u32 domain 12 reserved; then raw arrays.
// The oracle does not know the kinds: it moves the payload as one byte string (the tests slice it), which is enough to check
// that the three implementations (Python, CUDA library, oracle) agree on the container.
// returns 0 ok; 1 cannot open / short; 2 bad magic / version / size
int orc_blob_read(const char* path, uint64_t* hdr /*[8]: version kind count index0 aux payload_bytes domain 0*/, uint8_t* payload, uint64_t payload_cap) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return 1;
    unsigned char raw[64];
    if (std::fread(raw, 1, 64, f) != 64) { std::fclose(f); return 1; }
    if (std::memcmp(raw, "OMRB200\0", 8) != 0) { std::fclose(f); return 2; }
    auto rd = [&](int off, int n) { uint64_t v = 0; for (int i = 0; i < n; ++i) v |= (uint64_t)raw[off + i] << (8 * i); return v; };
    hdr[0] = rd(8, 4); hdr[1] = rd(12, 4); hdr[2] = rd(16, 8); hdr[3] = rd(24, 8); hdr[4] = rd(32, 8); hdr[5] = rd(40, 8); hdr[6] = rd(48, 4); hdr[7] = 0;
    if (hdr[0] != 1) { std::fclose(f); return 2; }
    int st = 0;
    if (payload) {
        if (hdr[5] > payload_cap) st = 2;
        else if (std::fread(payload, 1, hdr[5], f) != hdr[5]) st = 1;
        else if (std::fgetc(f) != EOF) st = 2;
    }
    std::fclose(f);
    return st;
}
int orc_blob_write(const char* path, uint32_t kind, uint64_t count, uint64_t index0, uint64_t aux, uint32_t domain, const uint8_t* payload, uint64_t payload_bytes) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return 1;
    unsigned char raw[64] = {0};
    std::memcpy(raw, "OMRB200\0", 8);
    auto wr = [&](int off, int n, uint64_t v) { for (int i = 0; i < n; ++i) raw[off + i] = (unsigned char)(v >> (8 * i)); };
    wr(8, 4, 1); wr(12, 4, kind); wr(16, 8, count); wr(24, 8, index0); wr(32, 8, aux); wr(40, 8, payload_bytes); wr(48, 4, domain);
    bool ok = std::fwrite(raw, 1, 64, f) == 64 && std::fwrite(payload, 1, payload_bytes, f) == payload_bytes;
    return (std::fclose(f) == 0 && ok) ? 0 : 1;
}

}  // extern "C"

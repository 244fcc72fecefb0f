// blob.cu — versioned flat blobs for the data either side of the hot path, host code only.
// The reference serialises nothing (it only reports byte counts through `Size`: key_gen/detection.rs:81-88, sender.rs:36);
// these little-endian containers are what the Rust shim (ffi/omr-b200-sys), the CPU oracle and this library exchange: keys,
// clues, stage outputs, pertinency vectors, digests.  Same format as tfhe-omr_b200/blobs.py:
//   header (64 bytes): magic "OMRB200\0" | u32 version | u32 kind | u64 count | u64 index0 | u64 aux | u64 payload bytes |
//                      u32 domain | 12 B reserved (zero)
//   payload: the arrays of the kind in declaration order, C-contiguous, little-endian
#include "../../include/omr_b200.h"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace omr { void set_global_error(const std::string& m); }

namespace {
struct Field { const char* name; uint32_t elem_bytes; uint64_t fixed; uint64_t per_count; };   // elements = fixed + per_count * count
struct Kind { uint32_t id; const char* name; std::vector<Field> fields; };
const std::vector<Kind>& kinds() {
    static const std::vector<Kind> k = {
        {OMR_BLOB_DETECTION_KEY, "detection_key", {{"bsk1", 4, 512ull * 8 * 2 * 1024, 0}, {"ksk", 4, 1024ull * 27 * 671, 0},
                                                   {"bsk2", 8, 670ull * 12 * 2 * 2048, 0}, {"trace", 8, 11ull * 25 * 2 * 2048, 0}}},
        {OMR_BLOB_CLUES, "clues", {{"a", 2, 0, 512}, {"b", 2, 0, 7}}},
        {OMR_BLOB_PERTINENCY_VECTOR, "pertinency_vector", {{"pv", 8, 0, 2 * 2048}}},
        {OMR_BLOB_DIGEST, "digest", {{"ct", 8, 0, 2 * 2048}}},
        {OMR_BLOB_PAYLOADS, "payloads", {{"payloads", 2, 0, 612}}},
        {OMR_BLOB_SECRET_KEY, "secret_key", {{"s0", 4, 512, 0}, {"z1", 4, 1024, 0}, {"s2", 4, 670, 0}, {"z2", 4, 2048, 0}}},
        {OMR_BLOB_RLWE1, "rlwe1", {{"ct", 4, 0, 2 * 1024}}},
        {OMR_BLOB_LWE2, "lwe2", {{"ct", 4, 0, 671}}},
        {OMR_BLOB_RLWE2, "rlwe2", {{"ct", 8, 0, 2 * 2048}}},
        {OMR_BLOB_CLUE_KEY, "clue_key", {{"pa", 2, 512, 0}, {"pb", 2, 512, 0}}},
    };
    return k;
}
const Kind* find_kind(uint32_t id) { for (auto& k : kinds()) if (k.id == id) return &k; return nullptr; }
const char MAGIC[8] = {'O', 'M', 'R', 'B', '2', '0', '0', '\0'};
int fail(const std::string& m) { omr::set_global_error(m); return OMR_ERR_INVALID; }
void put32(unsigned char* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (unsigned char)(v >> (8 * i)); }
void put64(unsigned char* p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (unsigned char)(v >> (8 * i)); }
uint32_t get32(const unsigned char* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)p[i] << (8 * i); return v; }
uint64_t get64(const unsigned char* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i); return v; }
static_assert(__BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__, "array payloads are written as they lie in memory");

int parse_header(FILE* f, omr_blob_header* h, const std::string& path) {
    unsigned char raw[64];
    if (fread(raw, 1, 64, f) != 64) return fail(path + ": truncated header");
    if (memcmp(raw, MAGIC, 8) != 0) return fail(path + ": not an OMRB200 blob");
    h->version = get32(raw + 8); h->kind = get32(raw + 12); h->count = get64(raw + 16); h->index0 = get64(raw + 24);
    h->aux = get64(raw + 32); h->payload_bytes = get64(raw + 40); h->domain = get32(raw + 48);
    h->reserved[0] = get32(raw + 52); h->reserved[1] = get32(raw + 56); h->reserved[2] = get32(raw + 60);
    if (h->version != OMR_BLOB_VERSION) return fail(path + ": unsupported blob version " + std::to_string(h->version));
    const Kind* k = find_kind(h->kind);
    if (!k) return fail(path + ": unknown blob kind " + std::to_string(h->kind));
    if (h->count > (1ull << 40)) return fail(path + ": implausible count");
    uint64_t total = 0;
    for (auto& fl : k->fields) total += (fl.fixed + fl.per_count * h->count) * fl.elem_bytes;
    if (total != h->payload_bytes) return fail(path + ": payload size mismatch");
    return OMR_OK;
}
}  // namespace

extern "C" {

size_t omr_blob_field_bytes(uint32_t kind, uint32_t field, uint64_t count) {
    const Kind* k = find_kind(kind);
    if (!k || field >= k->fields.size()) return 0;
    return (size_t)((k->fields[field].fixed + k->fields[field].per_count * count) * k->fields[field].elem_bytes);
}
uint32_t omr_blob_field_count(uint32_t kind) { const Kind* k = find_kind(kind); return k ? (uint32_t)k->fields.size() : 0; }

int omr_blob_write(const char* path, uint32_t kind, uint64_t count, uint64_t index0, uint64_t aux, uint32_t domain,
                   const void* const* arrays, uint32_t n_arrays) {
    const Kind* k = find_kind(kind);
    if (!path || !arrays || !k) return fail("blob_write: bad argument / unknown kind");
    if (n_arrays != k->fields.size()) return fail(std::string("blob_write: kind ") + k->name + " has " + std::to_string(k->fields.size()) + " arrays");
    if (domain != OMR_OUT_NTT_NATIVE && domain != OMR_OUT_COEFF) return fail("blob_write: unknown domain");
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_arrays; ++i) {
        if (!arrays[i] && omr_blob_field_bytes(kind, i, count)) return fail("blob_write: null array");
        total += omr_blob_field_bytes(kind, i, count);
    }
    FILE* f = fopen(path, "wb");
    if (!f) return fail(std::string("blob_write: cannot open ") + path);
    unsigned char raw[64] = {0};
    memcpy(raw, MAGIC, 8); put32(raw + 8, OMR_BLOB_VERSION); put32(raw + 12, kind); put64(raw + 16, count); put64(raw + 24, index0);
    put64(raw + 32, aux); put64(raw + 40, total); put32(raw + 48, domain);
    bool ok = fwrite(raw, 1, 64, f) == 64;
    for (uint32_t i = 0; ok && i < n_arrays; ++i) {
        const size_t n = omr_blob_field_bytes(kind, i, count);
        ok = n == 0 || fwrite(arrays[i], 1, n, f) == n;
    }
    ok = (fclose(f) == 0) && ok;
    return ok ? OMR_OK : fail(std::string("blob_write: short write to ") + path);
}

int omr_blob_read_header(const char* path, omr_blob_header* hdr) {
    if (!path || !hdr) return fail("blob_read_header: null argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(std::string("blob_read: cannot open ") + path);
    int st = parse_header(f, hdr, path);
    if (st == OMR_OK) {                                   // the file must hold exactly header + payload
        if (fseek(f, 0, SEEK_END) != 0 || (uint64_t)ftell(f) != 64 + hdr->payload_bytes) st = fail(std::string(path) + ": truncated payload");
    }
    fclose(f);
    return st;
}

int omr_blob_read(const char* path, omr_blob_header* hdr, void* const* arrays, uint32_t n_arrays) {
    omr_blob_header h;
    int st = omr_blob_read_header(path, &h);
    if (st) return st;
    if (hdr) *hdr = h;
    const Kind* k = find_kind(h.kind);
    if (!arrays || n_arrays != k->fields.size()) return fail(std::string("blob_read: kind ") + k->name + " has " + std::to_string(k->fields.size()) + " arrays");
    FILE* f = fopen(path, "rb");
    if (!f || fseek(f, 64, SEEK_SET) != 0) { if (f) fclose(f); return fail(std::string("blob_read: cannot open ") + path); }
    bool ok = true;
    for (uint32_t i = 0; ok && i < n_arrays; ++i) {
        const size_t n = omr_blob_field_bytes(h.kind, i, h.count);
        if (n && !arrays[i]) { ok = false; break; }
        ok = n == 0 || fread(arrays[i], 1, n, f) == n;
    }
    fclose(f);
    return ok ? OMR_OK : fail(std::string(path) + ": truncated payload");
}

// Detector::new from a detection-key blob on disk (what ffi/omr-b200-sys's dump_vectors example writes)
int omr_ctx_create_from_blob(int device, const char* path, omr_ctx** out) {
    if (!out) return fail("create_from_blob: null argument");
    *out = nullptr;
    omr_blob_header h;
    int st = omr_blob_read_header(path, &h);
    if (st) return st;
    if (h.kind != OMR_BLOB_DETECTION_KEY) return fail(std::string(path) + ": not a detection-key blob");
    std::vector<uint32_t> bsk1(omr_blob_field_bytes(h.kind, 0, 0) / 4), ksk(omr_blob_field_bytes(h.kind, 1, 0) / 4);
    std::vector<uint64_t> bsk2(omr_blob_field_bytes(h.kind, 2, 0) / 8), trk(omr_blob_field_bytes(h.kind, 3, 0) / 8);
    void* arrays[4] = {bsk1.data(), ksk.data(), bsk2.data(), trk.data()};
    if ((st = omr_blob_read(path, nullptr, arrays, 4))) return st;
    omr_key_blobs kb{bsk1.data(), ksk.data(), bsk2.data(), trk.data(), h.domain == OMR_OUT_COEFF ? OMR_KEYS_COEFF : OMR_KEYS_NTT_NATIVE};
    return omr_ctx_create(device, &kb, out);
}

}  // extern "C"

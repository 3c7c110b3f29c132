// tests/native/host_check.cpp -- TEST HARNESS ONLY.
// Compiles the host/device-portable pieces of the product (huff_build.cuh, inflate_core.cuh) with g++ so
// their logic can be compared against the oracle on a machine without a GPU.  Nothing here ships:
// libb2f.so never contains or calls this file.
#include "../../libflate_b200/csrc/huff_build.cuh"
#include <cstring>
#include <vector>

using namespace b2f;

extern "C" {

void hc_code_lengths(const uint32_t *freq, int n, int cap, uint8_t *width) {
    static HuffWork W;
    huff_code_lengths(freq, n, cap, width, W);
}

// returns header bit count; litcode[288], distcode[32], hdr_words[kHdrWords]
uint32_t hc_block_codes(const uint32_t *hist, uint32_t *litcode, uint32_t *distcode, uint32_t *hdr_words) {
    static HuffWork W;
    return build_block_codes(hist, litcode, distcode, hdr_words, W);
}
void hc_fixed_codes(uint32_t *litcode, uint32_t *distcode) { build_fixed_codes(litcode, distcode); }

void hc_length_code(uint32_t len, uint32_t *out3) { length_code(len, out3[0], out3[1], out3[2]); }
void hc_dist_code(uint32_t dist, uint32_t *out3) { dist_code(dist, out3[0], out3[1], out3[2]); }
uint32_t hc_hdr_words(void) { return kHdrWords; }

}  // extern "C"

// ------------------------------------------------------------------ inflate core on the host
#include "../../libflate_b200/csrc/inflate_core.cuh"
namespace {
struct HostOut {
    uint8_t *o; uint64_t capacity;
    uint64_t cap() const { return capacity; }
    void lit(uint64_t pos, uint8_t b) { o[pos] = b; }
    void copy(uint64_t pos, uint32_t len, uint32_t dist) { for (uint32_t k = 0; k < len; k++) o[pos + k] = o[pos + k - dist]; }
    void raw(uint64_t pos, const uint8_t *src, uint64_t n) { memcpy(o + pos, src, n); }
    int fast(BitIn &, const InflateTables &, uint64_t &, uint64_t) { return 0; }
    void block_end(uint64_t s, uint64_t e, uint64_t, bool) { if (nblk < 4096) { bstart[nblk] = s; bend[nblk] = e; } nblk++; }
    uint64_t bstart[4096], bend[4096]; uint32_t nblk = 0;
};
struct NoSync { void operator()() const {} };
}
extern "C" {
// raw DEFLATE stream decode with the product's core. res[0]=status res[1]=out_len res[2]=consumed res[3]=end_bit
void hc_inflate(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap, int64_t *res) {
    static InflateTables T;
    BitIn b; bi_init(b, in, n, 0);
    static HostOut o; o.o = out; o.capacity = cap; o.nblk = 0;
    InflateResult R;
    inflate_blocks(b, T, o, 0, 0, 0xFFFFFFFFu, 0, 1, NoSync(), R);
    res[0] = R.status; res[1] = (int64_t)R.out_len; res[2] = (int64_t)R.consumed; res[3] = (int64_t)R.end_bit;
}
uint32_t hc_len_base(uint32_t k) { return len_base(k) | (len_extra(k) << 16); }
uint32_t hc_dist_base(uint32_t k) { return dist_base(k) | (dist_extra(k) << 16); }
}

#include "../../libflate_b200/csrc/finder_core.cuh"
extern "C" {
// block start bit offsets of a raw DEFLATE stream as the product's inflate core sees them (valid streams)
uint32_t hc_block_starts(const uint8_t *in, uint64_t n, uint8_t *scratch, uint64_t cap, uint64_t *starts, uint32_t max) {
    static InflateTables T; static HostOut o; o.o = scratch; o.capacity = cap; o.nblk = 0;
    BitIn b; bi_init(b, in, n, 0);
    InflateResult R;
    inflate_blocks(b, T, o, 0, 0, 0xFFFFFFFFu, 0, 1, NoSync(), R);
    uint32_t k = o.nblk < max ? o.nblk : max;
    for (uint32_t i = 0; i < k; i++) starts[i] = o.bstart[i];
    return o.nblk;
}
// finder: returns number of accepted candidate bit offsets in [0, 8n); writes the first `max`
uint32_t hc_find_candidates(const uint8_t *in, uint64_t n, uint64_t *cands, uint32_t max) {
    uint32_t k = 0;
    for (uint64_t q = 0; q + 17 <= n * 8; q++) {
        uint64_t w0 = 0, w1 = 0;
        for (int i = 0; i < 16; i++) { uint64_t by = (q >> 3) + i; uint64_t v = by < n ? in[by] : 0; if (i < 8) w0 |= v << (8 * i); else w1 |= v << (8 * (i - 8)); }
        uint64_t w2 = (q >> 3) + 16 < n ? in[(q >> 3) + 16] : 0;
        uint32_t sh = (uint32_t)(q & 7);
        uint64_t a = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
        uint64_t bb = sh ? (w1 >> sh) | (w2 << (64 - sh)) : w1;
        if (!hdr_precheck((uint32_t)a)) continue;
        if (!precode_check(a, bb)) continue;
        if (!validate_dynamic_header(in, n, q)) continue;
        if (k < max) cands[k] = q;
        k++;
    }
    return k;
}
// equivalence of the finder's fast tests with their written-out forms on pseudo-random windows; returns the mismatch count
uint32_t hc_finder_selftest(uint64_t seed, uint32_t n) {
    uint32_t bad = 0;
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto next = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    for (uint32_t i = 0; i < n; i++) {
        uint64_t w0 = next(), w1 = next();
        if (i & 1) { w0 &= next(); w1 &= next(); }                 // sparser patterns: more zero widths, more passes
        if ((i & 7) == 3) w0 = (w0 & ~(0x7ull << 0)) | 4;         // BTYPE = 10
        if (precode_check(w0, w1) != precode_check_loop(w0, w1)) bad++;
        const uint32_t m = hdr_precheck_mask32(w0);
        for (uint32_t j = 0; j < 32; j++) if (((m >> j) & 1u) != (hdr_precheck((uint32_t)(w0 >> j)) ? 1u : 0u)) bad++;
    }
    return bad;
}
}

// finder_core.cuh -- "is there a plausible dynamic-Huffman block header at this bit offset?"
// Used by the block-boundary finder that makes index-less streams decodable block-parallel
// (the reference decodes strictly in order: src/deflate/decode.rs:136-164).  A hit is only a CANDIDATE:
// the decoder trusts it only when the chain of block ends starting from bit 0 lands exactly on it.
// Checks are necessary conditions that every header written by libflate's (and zlib's) encoder satisfies:
//   BTYPE == 10, HLIT <= 29, HDIST <= 29                                  (src/deflate/symbol.rs:343-365)
//   code-length code complete (or a single 1-bit code)                     (huffman.rs:202-209 always yields that)
//   lit/len + distance widths decode to exactly HLIT+HDIST entries, EOB has a code,
//   lit/len code complete (or single symbol), distance code complete (or <= 1 symbol)
#pragma once
#include "common.cuh"

namespace b2f {

B2F_HD bool hdr_precheck(uint32_t w) {            // w = 32 stream bits starting at the candidate (LSB first)
    return ((w >> 1) & 3u) == 2u && ((w >> 3) & 31u) <= 29u && ((w >> 8) & 31u) <= 29u;
}

// bit j of the result = hdr_precheck(w >> j) for j < 32 (uses bits [0, 45) of w): BTYPE is "bit 1 clear, bit 2 set";
// a 5-bit field read LSB first is >= 30 exactly when its upper four bits are all set
B2F_HD uint32_t hdr_precheck_mask32(uint64_t w) {
    const uint64_t btype = ~(w >> 1) & (w >> 2);
    const uint64_t hlit_bad = (w >> 4) & (w >> 5) & (w >> 6) & (w >> 7);
    const uint64_t hdist_bad = (w >> 9) & (w >> 10) & (w >> 11) & (w >> 12);
    return (uint32_t)(btype & ~hlit_bad & ~hdist_bad);
}

// Kraft sum and non-zero count of four packed 3-bit widths (12-bit index): kraft | count << 12
static const uint16_t kKraft4Host[4096] = {
#include "kraft4_table.inc"
};
#if defined(__CUDACC__)
static __device__ const uint16_t kKraft4Dev[4096] = {
#include "kraft4_table.inc"
};
#endif
B2F_HD uint32_t kraft4(uint32_t idx) {
#if defined(__CUDA_ARCH__)
    return __ldg(&kKraft4Dev[idx]);
#else
    return kKraft4Host[idx];
#endif
}

// w0 = bits [0,64), w1 = bits [64,128) from the candidate.  The code-length code must be complete (or a single 1-bit code).
B2F_HD bool precode_check(uint64_t w0, uint64_t w1) {
    const uint32_t hclen = (uint32_t)((w0 >> 13) & 15u) + 4u;
    // the 19 three-bit fields start at bit 17; only the first hclen of them are present in the stream
    uint64_t f = (w0 >> 17) | (w1 << 47);
    f &= (1ull << (3u * hclen)) - 1ull;
    uint32_t acc = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (uint32_t k = 0; k < 5; k++) acc += kraft4((uint32_t)(f >> (12 * k)) & 4095u);
    const uint32_t kraft = acc & 4095u, nz = acc >> 12;
    return kraft == 128u || (nz == 1 && kraft == 64u);
}
// the same test written out field by field (reference for the table version; used by the host checks)
B2F_HD bool precode_check_loop(uint64_t w0, uint64_t w1) {
    const uint32_t hclen = (uint32_t)((w0 >> 13) & 15u) + 4u;
    const uint32_t f0 = (uint32_t)(w0 >> 17) & 0x3FFFFFFFu;
    const uint32_t f1 = (uint32_t)((w0 >> 47) | (w1 << 17)) & 0x7FFFFFFu;
    uint32_t kraft = 0, nz = 0;
    for (uint32_t i = 0; i < 19; i++) {
        const uint32_t len = i < 10 ? (f0 >> (3 * i)) & 7u : (f1 >> (3 * (i - 10))) & 7u;
        if (i < hclen && len) { kraft += 128u >> len; nz++; }
    }
    return kraft == 128u || (nz == 1 && kraft == 64u);
}

// small LSB-first reader for the validation pass (aligned 32-bit loads, zero fill past the end)
struct VBits { const uint8_t *p; uint64_t nbytes; uint64_t next; uint64_t bb; uint32_t bc; uint64_t pos; };
B2F_HD uint32_t vb_load32(const uint8_t *p, uint64_t off, uint64_t nbytes) {
    if (off + 4 <= nbytes && ((reinterpret_cast<uintptr_t>(p + off)) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p + off);
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4; k++) if (off + k < nbytes) v |= (uint32_t)p[off + k] << (8 * k);
    return v;
}
B2F_HD void vb_init(VBits &b, const uint8_t *p, uint64_t nbytes, uint64_t bitpos) {
    b.p = p; b.nbytes = nbytes; b.pos = bitpos; b.next = bitpos >> 3;
    uint32_t drop = (uint32_t)(bitpos & 7);
    uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(p + b.next) & 3);
    if (mis && b.next >= mis) { b.next -= mis; drop += 8 * mis; }
    b.bb = (uint64_t)vb_load32(p, b.next, nbytes) >> drop; b.bc = 32 - drop; b.next += 4;
}
B2F_HD uint32_t vb_get(VBits &b, uint32_t n) {          // n <= 16
    if (b.bc < 32) { b.bb |= (uint64_t)vb_load32(b.p, b.next, b.nbytes) << b.bc; b.bc += 32; b.next += 4; }
    uint32_t v = (uint32_t)b.bb & ((1u << n) - 1u);
    b.bb >>= n; b.bc -= n; b.pos += n;
    return v;
}

// Full header validation (rare path).  Returns true when every check passes.
B2F_HD bool validate_dynamic_header(const uint8_t *p, uint64_t nbytes, uint64_t bitpos) {
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    const uint64_t limit = nbytes * 8;
    if (bitpos + 17 > limit) return false;
    VBits b; vb_init(b, p, nbytes, bitpos);
    vb_get(b, 3);
    const uint32_t hlit = vb_get(b, 5) + 257, hdist = vb_get(b, 5) + 1, hclen = vb_get(b, 4) + 4;
    if (hlit > 286 || hdist > 30) return false;
    if (b.pos + 3 * hclen > limit) return false;
    uint8_t pw[19];
    for (int i = 0; i < 19; i++) pw[i] = 0;
    for (uint32_t k = 0; k < hclen; k++) pw[ORDER[k]] = (uint8_t)vb_get(b, 3);
    uint32_t cnt[8], first[8];
    for (int i = 0; i < 8; i++) cnt[i] = 0;
    for (int i = 0; i < 19; i++) cnt[pw[i]]++;
    cnt[0] = 0;
    uint32_t code = 0; first[0] = 0;
    for (uint32_t l = 1; l < 8; l++) { code = (code + cnt[l - 1]) << 1; first[l] = code; if (code + cnt[l] > (1u << l)) return false; }
    // 7-bit lookup table of the code-length code (symbol | width << 5, 0 = no code): one load per symbol instead of a bit-by-bit
    // canonical search -- the validation is the larger half of the finder's time and nearly all of it is this loop
    uint8_t lut[128];
    for (int i = 0; i < 128; i++) lut[i] = 0;
    {
        uint32_t nxt[8];
        for (int l = 0; l < 8; l++) nxt[l] = first[l];
        for (uint32_t s = 0; s < 19; s++) {
            const uint32_t l = pw[s];
            if (!l) continue;
            const uint32_t r = bitrev(nxt[l]++, l);
            for (uint32_t k = r; k < 128; k += (1u << l)) lut[k] = (uint8_t)(s | (l << 5));
        }
    }
    uint32_t total = 0, want = hlit + hdist, prev = 0;
    uint32_t lit_kraft = 0, dist_kraft = 0, nlit = 0, ndist = 0, eob_len = 0;
    while (total < want) {
        if (b.pos >= limit) return false;
        if (b.bc < 32) { b.bb |= (uint64_t)vb_load32(b.p, b.next, b.nbytes) << b.bc; b.bc += 32; b.next += 4; }
        const uint32_t e = lut[(uint32_t)b.bb & 127u];
        if (!e) return false;
        const uint32_t sym = e & 31u, used = e >> 5;
        b.bb >>= used; b.bc -= used; b.pos += used;
        uint32_t rep = 1, val = sym;
        if (sym == 16) { if (total == 0) return false; rep = vb_get(b, 2) + 3; val = prev; }
        else if (sym == 17) { rep = vb_get(b, 3) + 3; val = 0; }
        else if (sym == 18) { rep = vb_get(b, 7) + 11; val = 0; }
        if (b.pos > limit || total + rep > want) return false;
        if (val) {
            // rep entries of width val, split at the lit/len | distance boundary
            const uint32_t nl = total < hlit ? (total + rep <= hlit ? rep : hlit - total) : 0u, nd = rep - nl;
            lit_kraft += nl * (32768u >> val); nlit += nl;
            dist_kraft += nd * (32768u >> val); ndist += nd;
            if (nl && total <= 256 && total + nl > 256) eob_len = val;
            if (lit_kraft > 32768u || dist_kraft > 32768u) return false;      // over-subscribed: give up early
        }
        total += rep; prev = val;
    }
    if (!eob_len) return false;
    if (!(lit_kraft == 32768u || (nlit == 1 && lit_kraft == 16384u))) return false;
    if (!(dist_kraft == 32768u || (ndist <= 1 && dist_kraft <= 16384u))) return false;
    return true;
}

}  // namespace b2f

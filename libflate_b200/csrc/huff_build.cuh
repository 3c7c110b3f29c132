// huff_build.cuh -- length-limited canonical Huffman construction and the dynamic block header,
// with libflate's exact tie-breaks.  Plain serial code usable from host and device (the device
// runs it on one lane per DEFLATE block; the host build is used by tests/native to check it
// against the oracle without a GPU).
//
// Reference behaviour restated here (NOT ported: the reference builds node lists carrying symbol
// vectors; this is the count-only formulation):
//   height probe        src/huffman.rs:261-274   (max-heap of (-freq, width) tuples)
//   package-merge       src/huffman.rs:307-363   (stable leaves, package only if strictly lighter)
//   canonical codes     src/huffman.rs:35-55, bit-reversed for the LSB-first writer :19-28, :213-216
//   header / RLE        src/deflate/symbol.rs:343-386, 486-540
#pragma once
#include "common.cuh"

namespace b2f {

constexpr int kMaxSyms = 288;

struct HuffWork {
    uint64_t leafw[kMaxSyms];
    uint16_t leafs[kMaxSyms];
    uint64_t cur[2][2 * kMaxSyms];
    uint32_t flags[16][(2 * kMaxSyms + 31) / 32];   // per level: bit i set = item i is a package
    uint16_t lens[16];
    uint64_t hw[kMaxSyms];                          // binary min-heap keyed by (hw, hs)
    uint16_t hs[kMaxSyms];
    int hn;
};

B2F_HD bool heap_less(const HuffWork &W, int a, int b) {
    return W.hw[a] < W.hw[b] || (W.hw[a] == W.hw[b] && W.hs[a] < W.hs[b]);
}
B2F_HD void heap_push(HuffWork &W, uint64_t w, uint16_t s) {
    int i = W.hn++;
    W.hw[i] = w; W.hs[i] = s;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (!heap_less(W, i, p)) break;
        uint64_t tw = W.hw[p]; W.hw[p] = W.hw[i]; W.hw[i] = tw;
        uint16_t ts = W.hs[p]; W.hs[p] = W.hs[i]; W.hs[i] = ts;
        i = p;
    }
}
B2F_HD void heap_pop(HuffWork &W, uint64_t &w, uint16_t &s) {
    w = W.hw[0]; s = W.hs[0];
    int n = --W.hn;
    W.hw[0] = W.hw[n]; W.hs[0] = W.hs[n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && heap_less(W, l, m)) m = l;
        if (r < n && heap_less(W, r, m)) m = r;
        if (m == i) break;
        uint64_t tw = W.hw[m]; W.hw[m] = W.hw[i]; W.hw[i] = tw;
        uint16_t ts = W.hs[m]; W.hs[m] = W.hs[i]; W.hs[i] = ts;
        i = m;
    }
}

// EncoderBuilder::from_frequencies(freq, cap) -> code widths (0 for unused symbols).
B2F_HD void huff_code_lengths(const uint32_t *freq, int n, int cap, uint8_t *width, HuffWork &W) {
    for (int i = 0; i < n; i++) width[i] = 0;
    // --- H1: height of the unrestricted Huffman tree; ties: smaller weight first, then DEEPER first
    W.hn = 0;
    int nu = 0;
    for (int i = 0; i < n; i++) if (freq[i]) { heap_push(W, freq[i], 255); nu++; }
    if (nu == 0) return;
    while (W.hn > 1) {
        uint64_t w1, w2; uint16_t s1, s2;
        heap_pop(W, w1, s1); heap_pop(W, w2, s2);
        uint32_t d1 = 255u - s1, d2 = 255u - s2;
        uint32_t d = 1 + (d1 > d2 ? d1 : d2);
        heap_push(W, w1 + w2, (uint16_t)(255u - d));
    }
    int height = 255 - (int)W.hs[0];
    if (height < 1) height = 1;
    int L = cap < height ? cap : height;
    // --- leaves sorted by (weight, symbol)
    W.hn = 0;
    for (int i = 0; i < n; i++) if (freq[i]) heap_push(W, freq[i], (uint16_t)i);
    for (int r = 0; r < nu; r++) { uint64_t w; uint16_t s; heap_pop(W, w, s); W.leafw[r] = w; W.leafs[r] = s; }
    // --- H2: package-merge, count-only.  level 1 = leaves.
    int curi = 0;
    for (int r = 0; r < nu; r++) W.cur[0][r] = W.leafw[r];
    W.lens[1] = (uint16_t)nu;
    const int fw = (2 * kMaxSyms + 31) / 32;
    for (int k = 2; k <= L; k++) {
        const uint64_t *prev = W.cur[curi]; uint64_t *nxt = W.cur[curi ^ 1];
        int plen = W.lens[k - 1];
        int npk = plen >= 2 ? plen / 2 : plen;           // package(): lists shorter than 2 pass through unchanged
        for (int t = 0; t < fw; t++) W.flags[k][t] = 0;
        int ip = 0, il = 0, o = 0;
        while (ip < npk || il < nu) {
            bool take_pkg;
            uint64_t pw = 0;
            if (ip < npk) pw = plen >= 2 ? prev[2 * ip] + prev[2 * ip + 1] : prev[ip];
            if (ip >= npk) take_pkg = false;
            else if (il >= nu) take_pkg = true;
            else take_pkg = pw < W.leafw[il];             // strictly lighter, else the leaf goes first
            if (take_pkg) { nxt[o] = pw; W.flags[k][o >> 5] |= 1u << (o & 31); ip++; }
            else { nxt[o] = W.leafw[il]; il++; }
            o++;
        }
        W.lens[k] = (uint16_t)o;
        curi ^= 1;
    }
    // --- backward selection
    int lenL = W.lens[L];
    int sel = lenL >= 2 ? 2 * (lenL / 2) : lenL;
    for (int k = L; k >= 1; k--) {
        int npkg = 0;
        if (k >= 2) {
            int full = sel >> 5, rem = sel & 31;
            for (int t = 0; t < full; t++) {
#if defined(__CUDA_ARCH__)
                npkg += __popc(W.flags[k][t]);
#else
                npkg += __builtin_popcount(W.flags[k][t]);
#endif
            }
            if (rem) {
                uint32_t m = W.flags[k][full] & ((1u << rem) - 1);
#if defined(__CUDA_ARCH__)
                npkg += __popc(m);
#else
                npkg += __builtin_popcount(m);
#endif
            }
        }
        int nleaf = sel - npkg;
        for (int r = 0; r < nleaf; r++) width[W.leafs[r]]++;
        sel = 2 * npkg;
    }
}

// canonical code table: out[s] = width<<16 | bit-reversed code (ready for an LSB-first bit stream)
B2F_HD void canonical_codes(const uint8_t *width, int n, uint32_t *out) {
    uint32_t cnt[16];
    for (int w = 0; w < 16; w++) cnt[w] = 0;
    for (int s = 0; s < n; s++) cnt[width[s]]++;
    uint32_t next[16]; uint32_t code = 0; cnt[0] = 0;
    for (int w = 1; w < 16; w++) { code = (code + cnt[w - 1]) << 1; next[w] = code; }
    for (int s = 0; s < n; s++) {
        uint32_t w = width[s];
        if (!w) { out[s] = 0; continue; }
        uint32_t c = next[w]++;
        out[s] = (w << 16) | bitrev(c & ((1u << w) - 1), w);
    }
}

struct BitSink { uint32_t *w; uint32_t nbits; };
B2F_HD void sink_put(BitSink &b, uint32_t v, uint32_t n) {      // n <= 16, words pre-zeroed
    if (!n) return;
    uint32_t wi = b.nbits >> 5, sh = b.nbits & 31;
    b.w[wi] |= v << sh;
    if (sh + n > 32) b.w[wi + 1] |= v >> (32 - sh);
    b.nbits += n;
}

// DynamicHuffmanCodec::save: writes HLIT/HDIST/HCLEN, the code-length code and the RLE'd widths
// into words[] (zeroed here) and returns the number of bits.  litw[286], distw[30].
B2F_HD uint32_t build_dynamic_header(const uint8_t *litw, const uint8_t *distw, uint32_t *words, HuffWork &W) {
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    int lit_count = 257, dist_count = 1;
    for (int s = 285; s >= 257; s--) if (litw[s]) { lit_count = s + 1; break; }
    for (int s = 29; s >= 1; s--) if (distw[s]) { dist_count = s + 1; break; }
    // RLE (restarting at the table boundary).  (code, extra bits, extra value) packed as code | bits<<8 | extra<<16
    uint32_t *codes = (uint32_t *)W.cur[0];                 // reuse workspace: <= 316 entries
    int nc = 0;
    uint32_t cc[19];
    for (int i = 0; i < 19; i++) cc[i] = 0;
    for (int t = 0; t < 2; t++) {
        const uint8_t *wv = t ? distw : litw; int size = t ? dist_count : lit_count;
        int i = 0;
        while (i < size) {
            uint8_t v = wv[i]; int j = i + 1;
            while (j < size && wv[j] == v) j++;
            int c = j - i; i = j;
            if (v == 0) {
                while (c >= 11) { int k = c < 138 ? c : 138; codes[nc++] = 18u | (7u << 8) | ((uint32_t)(k - 11) << 16); cc[18]++; c -= k; }
                if (c >= 3) { codes[nc++] = 17u | (3u << 8) | ((uint32_t)(c - 3) << 16); cc[17]++; c = 0; }
                for (; c > 0; c--) { codes[nc++] = 0; cc[0]++; }
            } else {
                codes[nc++] = v; cc[v]++; c--;
                while (c >= 3) { int k = c < 6 ? c : 6; codes[nc++] = 16u | (2u << 8) | ((uint32_t)(k - 3) << 16); cc[16]++; c -= k; }
                for (; c > 0; c--) { codes[nc++] = v; cc[v]++; }
            }
        }
    }
    // NOTE: huff_code_lengths uses W.cur as scratch, and `codes` aliases W.cur[0]; the precode needs only
    // levels <= 7 over <= 19 leaves => it touches cur[*][0..37], so move the codes out of the way first.
    uint32_t *codes_hi = (uint32_t *)&W.cur[1][64];          // 512 u64 slots left: room for 316 u32
    for (int i = nc - 1; i >= 0; i--) codes_hi[i] = codes[i];
    uint8_t cw[19]; uint32_t ccode[19];
    huff_code_lengths(cc, 19, 7, cw, W);
    canonical_codes(cw, 19, ccode);
    int hclen = 0;
    for (int k = 18; k >= 0; k--) { int i = ORDER[k]; if (cc[i] != 0 && cw[i] > 0) { hclen = k + 1; break; } }
    if (hclen < 4) hclen = 4;
    for (uint32_t i = 0; i < kHdrWords; i++) words[i] = 0;
    BitSink bs = { words, 0 };
    sink_put(bs, (uint32_t)(lit_count - 257), 5);
    sink_put(bs, (uint32_t)(dist_count - 1), 5);
    sink_put(bs, (uint32_t)(hclen - 4), 4);
    for (int k = 0; k < hclen; k++) { int i = ORDER[k]; sink_put(bs, cc[i] == 0 ? 0u : (uint32_t)cw[i], 3); }
    for (int i = 0; i < nc; i++) {
        uint32_t e = codes_hi[i]; uint32_t c = e & 0xFF, nb = (e >> 8) & 0xFF, ex = e >> 16;
        sink_put(bs, ccode[c] & 0xFFFF, ccode[c] >> 16);
        if (nb) sink_put(bs, ex, nb);
    }
    return bs.nbits;
}

// Whole per-block code construction: DynamicHuffmanCodec::build + save (symbol.rs:321-386).
// hist: 286 lit/len counts then 30 distance counts (EOB not yet counted).  Outputs the two code
// tables (width<<16|revbits), the header bit buffer and its length.
B2F_HD uint32_t build_block_codes(const uint32_t *hist, uint32_t *litcode /*288*/, uint32_t *distcode /*32*/, uint32_t *hdr_words, HuffWork &W) {
    uint32_t lf[286], df[30];
    uint8_t lw[kLitStride], dw[kDistStride];
    bool any_dist = false;
    for (int i = 0; i < 286; i++) lf[i] = hist[i];
    lf[256] += 1;                                            // EndOfBlock pushed by CompressBuf::flush (encode.rs:418)
    for (int i = 0; i < 30; i++) { df[i] = hist[286 + i]; any_dist |= df[i] != 0; }
    if (!any_dist) df[0] = 1;                                // dummy distance code (symbol.rs:332-337)
    huff_code_lengths(lf, 286, 15, lw, W);
    huff_code_lengths(df, 30, 15, dw, W);
    lw[286] = lw[287] = 0; dw[30] = dw[31] = 0;
    canonical_codes(lw, 288, litcode);
    canonical_codes(dw, 32, distcode);
    return build_dynamic_header(lw, dw, hdr_words, W);
}

// FixedHuffmanCodec::build (symbol.rs:260-281)
B2F_HD void build_fixed_codes(uint32_t *litcode, uint32_t *distcode) {
    uint8_t lw[kLitStride], dw[kDistStride];
    for (int s = 0; s < 288; s++) lw[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
    for (int s = 0; s < 32; s++) dw[s] = s < 30 ? 5 : 0;
    canonical_codes(lw, 288, litcode);
    // 30 five-bit codes 0..29: canonical_codes over 30 symbols of width 5 gives exactly code i for symbol i
    canonical_codes(dw, 32, distcode);
}

}  // namespace b2f

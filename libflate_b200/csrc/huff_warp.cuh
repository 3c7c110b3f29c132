// huff_warp.cuh -- warp-cooperative form of huff_build.cuh (device only): one warp builds the codes of one DEFLATE block.
//
// Same results as the serial code (which stays the host-checkable statement of the algorithm, tests/native + the oracle), other
// evaluation order:
//   leaves sorted by (weight, symbol)   bitonic sort of 64-bit keys in shared memory           (src/huffman.rs:309-315)
//   height of the unrestricted tree     two-queue merge over the sorted leaves on one lane, ties resolved like the reference's
//                                        heap of (-freq, width): lighter first, then DEEPER first (src/huffman.rs:261-274)
//   package-merge, count only           every level is a merge of two sorted lists: each item's output position is its own index
//                                        plus a binary search in the other list (leaf first on ties = "package only if strictly
//                                        lighter", src/huffman.rs:329-349); 32 items per step
//   canonical codes                     rank among equal widths by __match_any_sync                   (src/huffman.rs:35-55)
//   header                              RLE on one lane (a few hundred steps), bit packing with a warp scan (symbol.rs:343-386,486-540)
// The serial version needs ~0.9 ms per block (the latency of ONE thread); this one ~0.1 ms.
#pragma once
#include "common.cuh"

namespace b2f {

struct HuffWarp {
    uint64_t key[512];                 // sort buffer, then scratch of the height probe (node weights)
    uint64_t lst[2][576];              // package-merge lists of two consecutive levels
    uint64_t pw[288];                  // package weights of the level being built
    uint32_t leafw[288];
    uint32_t flags[16][18];            // per level: bit i set = item i is a package
    uint32_t lf[288], df[32], cc[32];  // frequencies: lit/len, distance, code-length alphabet
    uint32_t next[16];                 // canonical code counters
    uint32_t codes[320];               // RLE'd header: code | extra bit count << 8 | extra value << 16
    uint32_t ccode[32];                // code-length alphabet: width << 16 | reversed code
    uint16_t leafs[288];
    uint16_t lens[16];
    uint8_t lw[288], dw[32], cw[32];   // code widths
    uint8_t depth[288];                // height probe: depth of the internal nodes
};

__device__ __forceinline__ uint32_t hw_lane() { return threadIdx.x & 31u; }

// EncoderBuilder::from_frequencies(freq, cap) -> code widths (0 for unused symbols).  freq, width: shared memory.
__device__ inline void hw_code_lengths(const uint32_t *freq, uint32_t n, uint32_t cap, uint8_t *width, HuffWarp &H) {
    const uint32_t lane = hw_lane();
    const uint32_t N = n > 32 ? 512u : 32u;
    uint32_t nu = 0;
    for (uint32_t i = lane; i < N; i += 32) {
        const uint32_t f = i < n ? freq[i] : 0u;
        if (i < n) width[i] = 0;
        H.key[i] = f ? ((uint64_t)f << 16) | i : ~0ull;
        nu += __popc(__ballot_sync(0xFFFFFFFFu, f != 0));
    }
    __syncwarp();
    if (nu == 0) return;
    // ---- leaves sorted by (weight, symbol)
    for (uint32_t k = 2; k <= N; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = lane; t < N / 2; t += 32) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), p = i + j;
                const uint64_t a = H.key[i], b = H.key[p];
                const bool asc = (i & k) == 0;
                if ((a > b) == asc) { H.key[i] = b; H.key[p] = a; }
            }
            __syncwarp();
        }
    }
    for (uint32_t r = lane; r < nu; r += 32) { const uint64_t kv = H.key[r]; H.leafw[r] = (uint32_t)(kv >> 16); H.leafs[r] = (uint16_t)(kv & 0xFFFFu); }
    __syncwarp();
    if (nu == 1) { if (lane == 0) width[H.leafs[0]] = 1; __syncwarp(); return; }
    // ---- H1: height of the unrestricted Huffman tree (one lane; the internal nodes come out in non-decreasing weight order)
    uint32_t height = 0;
    if (lane == 0) {
        uint64_t *iw = H.key;                                         // sorting is done: reuse as the queue of internal nodes
        uint32_t li = 0, qi = 0, qn = 0;
        auto pop = [&](uint64_t &w, uint32_t &d) {
            const bool has_leaf = li < nu, has_int = qi < qn;
            if (has_int && (!has_leaf || iw[qi] <= H.leafw[li])) {    // same weight: an internal node is deeper than a leaf
                uint32_t m = qi;                                      // among internal nodes of that weight: the deepest
                for (uint32_t t = qi + 1; t < qn && iw[t] == iw[qi]; t++) if (H.depth[t] > H.depth[m]) m = t;
                const uint8_t dm = H.depth[m]; H.depth[m] = H.depth[qi]; H.depth[qi] = dm;
                w = iw[qi]; d = dm; qi++;
            } else { w = H.leafw[li]; d = 0; li++; }
        };
        for (uint32_t r = 0; r + 1 < nu; r++) {
            uint64_t w1, w2; uint32_t d1, d2;
            pop(w1, d1); pop(w2, d2);
            iw[qn] = w1 + w2; H.depth[qn] = (uint8_t)(1 + max(d1, d2)); qn++;
        }
        height = H.depth[qn - 1];
    }
    height = max(1u, __shfl_sync(0xFFFFFFFFu, height, 0));
    const uint32_t L = min(cap, height);
    // ---- H2: package-merge, count only.  level 1 = the leaves.
    for (uint32_t r = lane; r < nu; r += 32) H.lst[0][r] = H.leafw[r];
    if (lane == 0) H.lens[1] = (uint16_t)nu;
    __syncwarp();
    uint32_t curi = 0;
    for (uint32_t k = 2; k <= L; k++) {
        const uint64_t *prev = H.lst[curi]; uint64_t *nxt = H.lst[curi ^ 1];
        const uint32_t plen = H.lens[k - 1];
        const uint32_t npk = plen >= 2 ? plen / 2 : plen;            // package(): lists shorter than 2 pass through unchanged
        for (uint32_t j = lane; j < npk; j += 32) H.pw[j] = plen >= 2 ? prev[2 * j] + prev[2 * j + 1] : prev[j];
        if (lane < 18) H.flags[k][lane] = 0;
        __syncwarp();
        for (uint32_t j = lane; j < npk; j += 32) {                  // a package goes after every leaf that is not heavier
            const uint64_t w = H.pw[j];
            uint32_t lo = 0, hi = nu;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint64_t)H.leafw[mid] <= w) lo = mid + 1; else hi = mid; }
            const uint32_t o = j + lo;
            nxt[o] = w;
            atomicOr(&H.flags[k][o >> 5], 1u << (o & 31));
        }
        for (uint32_t i = lane; i < nu; i += 32) {                   // a leaf goes after every package that is strictly lighter
            const uint64_t w = H.leafw[i];
            uint32_t lo = 0, hi = npk;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (H.pw[mid] < w) lo = mid + 1; else hi = mid; }
            nxt[i + lo] = w;
        }
        if (lane == 0) H.lens[k] = (uint16_t)(npk + nu);
        __syncwarp();
        curi ^= 1;
    }
    // ---- backward selection
    const uint32_t lenL = H.lens[L];
    uint32_t sel = lenL >= 2 ? 2 * (lenL / 2) : lenL;
    for (uint32_t k = L; k >= 1; k--) {
        uint32_t npkg = 0;
        if (k >= 2) {
            uint32_t c = 0;
            if (lane < 18) {
                const uint32_t lo_bit = lane * 32;
                uint32_t wv = H.flags[k][lane];
                if (sel <= lo_bit) wv = 0; else if (sel < lo_bit + 32) wv &= (1u << (sel - lo_bit)) - 1u;
                c = __popc(wv);
            }
            npkg = __reduce_add_sync(0xFFFFFFFFu, c);
        }
        const uint32_t nleaf = sel - npkg;
        for (uint32_t r = lane; r < nleaf; r += 32) width[H.leafs[r]]++;
        __syncwarp();
        sel = 2 * npkg;
    }
}

// canonical code table: out[s] = width<<16 | bit-reversed code (ready for an LSB-first bit stream); out: global or shared
__device__ inline void hw_canonical(const uint8_t *width, uint32_t n, uint32_t *out, HuffWarp &H) {
    const uint32_t lane = hw_lane();
    uint32_t mycnt = 0;                                               // lane w counts the symbols of width w
    for (uint32_t s0 = 0; s0 < n; s0 += 32) {
        const uint32_t wv = s0 + lane < n ? width[s0 + lane] : 0u;
#pragma unroll
        for (uint32_t w = 1; w < 16; w++) { const uint32_t m = __ballot_sync(0xFFFFFFFFu, wv == w); if (lane == w) mycnt += __popc(m); }
    }
    uint32_t code = 0, mynext = 0;
    for (uint32_t w = 1; w < 16; w++) {
        code = (code + __shfl_sync(0xFFFFFFFFu, mycnt, w - 1)) << 1;  // (lane 0 holds 0: width 0 is not a code)
        if (lane == w) mynext = code;
    }
    if (lane < 16) H.next[lane] = mynext;
    __syncwarp();
    for (uint32_t s0 = 0; s0 < n; s0 += 32) {
        const uint32_t s = s0 + lane;
        const uint32_t wv = s < n ? width[s] : 0u;
        const uint32_t same = __match_any_sync(0xFFFFFFFFu, wv);
        const uint32_t c = H.next[wv & 15u] + __popc(same & ((1u << lane) - 1u));
        if (s < n) out[s] = wv ? (wv << 16) | bitrev(c & ((1u << wv) - 1u), wv) : 0u;
        __syncwarp();
        if (wv && (same & ((1u << lane) - 1u)) == 0) H.next[wv] += __popc(same);       // the lowest lane of every width advances its counter
        __syncwarp();
    }
}

// OR v (n <= 32 bits) into the zeroed word buffer at bit position pos (several lanes at once)
__device__ __forceinline__ void hw_put(uint32_t *words, uint32_t pos, uint32_t v, uint32_t n) {
    if (!n) return;
    const uint32_t wi = pos >> 5, sh = pos & 31u;
    atomicOr(&words[wi], v << sh);
    if (sh + n > 32) atomicOr(&words[wi + 1], v >> (32 - sh));
}

// DynamicHuffmanCodec::build + save for one block (symbol.rs:321-386): hist = 286 lit/len counts then 30 distance counts (EOB
// not yet counted), in global memory.  Writes the two code tables, the header bit buffer (kHdrWords words) and returns its length.
__device__ inline uint32_t hw_build_block_codes(const uint32_t *hist, uint32_t *litcode, uint32_t *distcode, uint32_t *hdr_words, HuffWarp &H) {
    const uint32_t lane = hw_lane();
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    for (uint32_t i = lane; i < 288; i += 32) H.lf[i] = i < 286 ? hist[i] + (i == 256 ? 1u : 0u) : 0u;      // + EndOfBlock (encode.rs:418)
    const uint32_t dfv = lane < 30 ? hist[286 + lane] : 0u;
    const bool any_dist = __any_sync(0xFFFFFFFFu, dfv != 0);
    H.df[lane] = (!any_dist && lane == 0) ? 1u : dfv;                 // dummy distance code (symbol.rs:332-337)
    __syncwarp();
    hw_code_lengths(H.lf, 286, 15, H.lw, H);
    hw_code_lengths(H.df, 30, 15, H.dw, H);
    if (lane < 2) { H.lw[286 + lane] = 0; H.dw[30 + lane] = 0; }
    __syncwarp();
    hw_canonical(H.lw, 288, litcode, H);
    hw_canonical(H.dw, 32, distcode, H);
    // ---- header: HLIT / HDIST, RLE of the widths (restarting at the table boundary), code-length code, HCLEN
    const uint32_t ml = __ballot_sync(0xFFFFFFFFu, lane < 29 && H.lw[257 + lane] != 0);
    const uint32_t md = __ballot_sync(0xFFFFFFFFu, lane >= 1 && lane < 30 && H.dw[lane] != 0);
    const uint32_t lit_count = ml ? 257u + (32u - (uint32_t)__clz((int)ml)) : 257u;
    const uint32_t dist_count = md ? 32u - (uint32_t)__clz((int)md) : 1u;
    if (lane < 19) H.cc[lane] = 0;
    __syncwarp();
    uint32_t nc = 0;
    if (lane == 0) {
        for (uint32_t t = 0; t < 2; t++) {
            const uint8_t *wv = t ? H.dw : H.lw; const uint32_t size = t ? dist_count : lit_count;
            uint32_t i = 0;
            while (i < size) {
                const uint32_t v = wv[i]; uint32_t j = i + 1;
                while (j < size && wv[j] == v) j++;
                uint32_t c = j - i; i = j;
                if (v == 0) {
                    while (c >= 11) { const uint32_t k = c < 138 ? c : 138; H.codes[nc++] = 18u | (7u << 8) | ((k - 11) << 16); H.cc[18]++; c -= k; }
                    if (c >= 3) { H.codes[nc++] = 17u | (3u << 8) | ((c - 3) << 16); H.cc[17]++; c = 0; }
                    for (; c > 0; c--) { H.codes[nc++] = 0; H.cc[0]++; }
                } else {
                    H.codes[nc++] = v; H.cc[v]++; c--;
                    while (c >= 3) { const uint32_t k = c < 6 ? c : 6; H.codes[nc++] = 16u | (2u << 8) | ((k - 3) << 16); H.cc[16]++; c -= k; }
                    for (; c > 0; c--) { H.codes[nc++] = v; H.cc[v]++; }
                }
            }
        }
    }
    nc = __shfl_sync(0xFFFFFFFFu, nc, 0);
    __syncwarp();
    hw_code_lengths(H.cc, 19, 7, H.cw, H);
    hw_canonical(H.cw, 19, H.ccode, H);
    const uint32_t oi = lane < 19 ? ORDER[lane] : 0u;
    const uint32_t mu = __ballot_sync(0xFFFFFFFFu, lane < 19 && H.cc[oi] != 0 && H.cw[oi] > 0);
    const uint32_t hclen = max(4u, mu ? 32u - (uint32_t)__clz((int)mu) : 0u);
    for (uint32_t i = lane; i < kHdrWords; i += 32) hdr_words[i] = 0;
    __syncwarp();
    if (lane == 0) { hw_put(hdr_words, 0, lit_count - 257, 5); hw_put(hdr_words, 5, dist_count - 1, 5); hw_put(hdr_words, 10, hclen - 4, 4); }
    if (lane < hclen) hw_put(hdr_words, 14 + 3 * lane, H.cc[oi] == 0 ? 0u : (uint32_t)H.cw[oi], 3);
    uint32_t pos = 14 + 3 * hclen;
    for (uint32_t i0 = 0; i0 < nc; i0 += 32) {
        const uint32_t i = i0 + lane;
        uint32_t v = 0, nb = 0;
        if (i < nc) {
            const uint32_t e = H.codes[i], cd = H.ccode[e & 0xFFu], cwid = cd >> 16, xb = (e >> 8) & 0xFFu;
            v = (cd & 0xFFFFu) | ((e >> 16) << cwid); nb = cwid + xb;             // <= 7 + 7 bits
        }
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
        hw_put(hdr_words, pos + incl - nb, v, nb);
        pos += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    return pos;
}

}  // namespace b2f

// encode_kernels.cu -- the DEFLATE encode hot path as hand-written sm_100a kernels.
//
// Pipeline (one launch each, all on the ctx stream; see DESIGN.md for bytes/unit and rooflines):
//   K1 lz_chain      one warp per <=256 KiB segment: ordered hash chains out of shared memory   (E2/A2)
//   K2 lz_match      CTA per 16 KiB tile, 48 KiB window staged in shared memory: prev + LCP     (E2/L1)
//   K3 parse_exits   thread per 2 KiB tile: right-to-left exit DP of the greedy walk            (E2)
//   K4 parse_stitch  thread per chunk: chain the tile entry points
//   K5 parse_emit    thread per tile: walk, emit symbols, per-block histograms                  (S1/S2)
//   K6 huff_build    one lane per DEFLATE block: code lengths, canonical codes, header bits     (H1-H3/S3/S4)
//   K7 tile_bits     warp per tile: coded size of the tile
//   K8 scan_tiles / scan_blocks: bit offsets of every tile / block / stream                     (B2, E1)
//   K9 write_headers warp per block: BFINAL/BTYPE, dynamic header, EOB, sync marker             (B2)
//   K10 bitpack      warp per tile: LSB-first packing in shared memory, coalesced store         (B1)
// Reference behaviour: libflate_lz77/src/default.rs:59-183, src/deflate/encode.rs:261-426,
// src/deflate/symbol.rs:95-183,321-386,486-540, src/huffman.rs:35-55,192-363, src/bit.rs:25-49.
#include "common.cuh"
#include "huff_build.cuh"
#include "encode_dev.cuh"

namespace b2f {

// index of the chunk whose prefix range contains idx: prefix[c] <= idx < prefix[c+1]
__device__ __forceinline__ uint32_t find_owner(const uint32_t *__restrict__ prefix, uint32_t n, uint32_t idx) {
    uint32_t lo = 0, hi = n;                 // invariant: prefix[lo] <= idx < prefix[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(prefix + mid) <= idx) lo = mid; else hi = mid;
    }
    return lo;
}

// 32-bit load at byte offset `off` of the input that never touches bytes at or beyond in_size (zero fill)
__device__ __forceinline__ uint32_t ld_in32(const uint8_t *__restrict__ in, uint64_t off, uint64_t in_size) {
    if (off + 4 <= in_size) return __ldg(reinterpret_cast<const uint32_t *>(in + off));
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4; k++) if (off + k < in_size) v |= (uint32_t)in[off + k] << (8 * k);
    return v;
}

// =============================================================================== K1 lz_chain
// link[p] = distance to the nearest earlier position of the same chunk whose trigram has the same
// 14-bit hash (0 = none within 32768).  Every position < end is "inserted" exactly once in
// increasing order (default.rs:78, 92-97), so this ordered chain is parse independent.
__global__ void __launch_bounds__(32) k_lz_chain(EncDev E, uint32_t off) {
    extern __shared__ uint32_t head[];       // 1 << kHashBits entries: last position + 1 (exact: a 16-bit modulo entry would alias
                                             // stale buckets into bogus in-window links and send lz_match down unrelated chains)
    const uint32_t lane = threadIdx.x;
    const uint32_t seg = blockIdx.x + off;
    const uint32_t c = find_owner(E.seg0, E.n_chunks, seg);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t n = cd.len;
    const uint32_t end = (n > 3 ? n : 3) - 3;
    const uint32_t s_start = (seg - E.seg0[c]) * kSeg;
    const uint32_t s_end = min(s_start + kSeg, n);
    const uint32_t lim = min(s_end, end);
    const uint32_t ws = s_start > kLookback ? s_start - kLookback : 0;
    for (uint32_t i = lane; i < (1u << kHashBits); i += 32) head[i] = 0;
    __syncwarp();
    uint16_t *__restrict__ lk = E.link + cd.off;
    for (uint32_t q = max(lim, s_start) + lane; q < s_end; q += 32) lk[q] = 0;   // tail without a trigram
    if (ws >= lim) return;
    // The input is consumed as 128-byte blocks: lane L holds the aligned word at block + 4L; blocks are prefetched
    // several iterations ahead, so no load sits on the critical path.  Positions are int64 relative to the chunk because the
    // aligned start can lie up to 3 bytes before position `ws` (those lanes are masked off).
    const int64_t a0 = (int64_t)ws - (int64_t)((cd.off + ws) & 3u);
    const uint64_t g0 = (uint64_t)((int64_t)cd.off + a0);          // 4-byte aligned offset into E.in
    // One shared-memory atomicMax per position publishes pos+1 and returns the previous head.  Positions grow monotonically, so
    // the table always ends up with the most recent position; lanes of one step that share a bucket are serialised by the
    // hardware and, when that happens in ascending lane order, each one receives exactly its predecessor.  Any other order makes
    // some lane see a value above its own position: that is detected once per 128-byte block and repaired exactly (below).
    // The four steps of a block are independent instructions, so their atomics are in flight together.
    // Input blocks are prefetched kChainAhead iterations ahead (a cold HBM read costs about two block iterations of this warp).
    constexpr uint32_t kChainAhead = 4;
    uint32_t pre[kChainAhead];
#pragma unroll
    for (uint32_t k = 0; k < kChainAhead; k++) pre[k] = ld_in32(E.in, g0 + 128ull * k + 4ull * lane, E.in_size);
    uint32_t blk = 0;
    for (int64_t bpos = a0; bpos < (int64_t)lim; bpos += 128, blk++) {
        const uint32_t cur = pre[0], nxt = pre[1];
#pragma unroll
        for (uint32_t k = 0; k + 1 < kChainAhead; k++) pre[k] = pre[k + 1];
        pre[kChainAhead - 1] = ld_in32(E.in, g0 + 128ull * (blk + kChainAhead) + 4ull * lane, E.in_size);
        const uint32_t hn = __shfl_sync(0xFFFFFFFFu, nxt, 0);
        uint32_t hh[4], oo[4];
        bool vv[4];
        bool bad = false;
#pragma unroll
        for (uint32_t s4 = 0; s4 < 4; s4++) {
            const int64_t base = bpos + 32 * s4;
            const uint32_t wi = (32 * s4 + lane) >> 2, bo = lane & 3u;
            const uint32_t lo = __shfl_sync(0xFFFFFFFFu, cur, wi);
            const uint32_t hc = __shfl_sync(0xFFFFFFFFu, cur, (wi + 1) & 31u);
            const uint32_t hi = wi == 31 ? hn : hc;
            const uint32_t t = __funnelshift_r(lo, hi, bo * 8u) & 0xFFFFFFu;
            const int64_t pos64 = base + lane;
            vv[s4] = pos64 >= (int64_t)ws && pos64 < (int64_t)lim;
            hh[s4] = (t * 0x9E3779B1u) >> (32 - kHashBits);
            oo[s4] = vv[s4] ? atomicMax(&head[hh[s4]], (uint32_t)pos64 + 1u) : 0u;
#ifdef B2F_CHAIN_FORCE_REPAIR                                    /* test build: always take the repair path */
            bad = true;
#else
            bad |= vv[s4] && oo[s4] > (uint32_t)pos64;
#endif
        }
        if (__any_sync(0xFFFFFFFFu, bad)) {
            // exact repair: within a step the predecessor of a lane is the nearest lower lane of its bucket; the lowest lane of a
            // bucket gets the head from before the step = the smallest value any lane of the bucket received
#pragma unroll
            for (uint32_t s4 = 0; s4 < 4; s4++) {
                const uint32_t pos = (uint32_t)(bpos + 32 * s4 + lane);
                const uint32_t key = vv[s4] ? hh[s4] : (0x80000000u | lane);
                const uint32_t m = __match_any_sync(0xFFFFFFFFu, key);
                const uint32_t lower = m & ((1u << lane) - 1u);
                uint32_t mn = oo[s4], mm = m & ~(1u << lane);
                while (__any_sync(0xFFFFFFFFu, mm != 0)) {
                    const uint32_t src = mm ? (uint32_t)__ffs((int)mm) - 1u : lane;
                    const uint32_t v = __shfl_sync(0xFFFFFFFFu, oo[s4], src);
                    if (mm) { mn = min(mn, v); mm &= mm - 1u; }
                }
                oo[s4] = lower ? pos + 1u - (lane - (31u - (uint32_t)__clz((int)lower))) : mn;
            }
        }
#pragma unroll
        for (uint32_t s4 = 0; s4 < 4; s4++) {
            const uint32_t pos = (uint32_t)(bpos + 32 * s4 + lane);
            if (vv[s4] && pos >= s_start) {
                uint32_t d = oo[s4] ? pos + 1u - oo[s4] : 0u;
                if (d > kLookback) d = 0;
                lk[pos] = (uint16_t)d;
            }
        }
    }
}

// =============================================================================== K2 lz_match
__device__ __forceinline__ uint32_t ld32u(const uint8_t *s, uint32_t off) {     // unaligned 4-byte read from smem
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}
constexpr uint32_t kMatchSmem = kPTile + kLookback + 261 + 16 + 32;

// md[p] = 0 (literal) or length<<16 | distance of the single candidate libflate would take at p:
// the most recent earlier occurrence of the same 3 bytes, if within `window` (default.rs:79-91, 116-129).
__global__ void __launch_bounds__(512) k_lz_match(EncDev E, uint32_t off) {
    extern __shared__ __align__(16) uint8_t sb[];
    const uint32_t tid = threadIdx.x;
    const uint32_t pt = blockIdx.x + off;
    const uint32_t c = find_owner(E.pt0, E.n_chunks, pt);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t n = cd.len;
    const uint32_t end = (n > 3 ? n : 3) - 3;
    const uint32_t ts = (pt - E.pt0[c]) * kPTile;
    const uint32_t te = min(ts + kPTile, n);
    const uint32_t lo = ts > kLookback ? ts - kLookback : 0;
    const uint32_t hi = min(n, te + 261);
    const uint64_t g_lo = cd.off + lo;
    const uint64_t g_al = g_lo & ~15ull;
    const uint32_t shift = (uint32_t)(g_lo - g_al);
    const uint32_t nvec = (hi - lo + shift + 15) >> 4;
    const uint4 *__restrict__ gsrc = reinterpret_cast<const uint4 *>(E.in + g_al);
    uint4 *sdst = reinterpret_cast<uint4 *>(sb);
    for (uint32_t i = tid; i < nvec; i += 512) {
        const uint64_t off = g_al + 16ull * i;
        if (off + 16 <= E.in_size) sdst[i] = __ldg(gsrc + i);
        else { uint4 v; v.x = ld_in32(E.in, off, E.in_size); v.y = ld_in32(E.in, off + 4, E.in_size); v.z = ld_in32(E.in, off + 8, E.in_size); v.w = ld_in32(E.in, off + 12, E.in_size); sdst[i] = v; }
    }
    __syncthreads();
    const uint16_t *__restrict__ lk = E.link + cd.off;
    uint32_t *__restrict__ md = E.md + cd.off;
    const uint32_t sbase = shift - lo;        // smem index of chunk position x is x + sbase (mod 2^32 arithmetic)
    for (uint32_t pos0 = ts + tid; pos0 < te; pos0 += 4 * 512) {
        uint32_t dpre[4];
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) { const uint32_t q = pos0 + u * 512; dpre[u] = (q < te && q < end) ? lk[q] : 0; }   // independent loads first
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t pos = pos0 + u * 512;
            if (pos >= te) break;
            uint32_t out = 0;
            if (pos < end) {
                const uint32_t si = pos + sbase;
                const uint32_t t = (uint32_t)sb[si] | ((uint32_t)sb[si + 1] << 8) | ((uint32_t)sb[si + 2] << 16);
                uint32_t d = dpre[u], total = 0, j = pos;
                bool found = false;
                while (d) {
                    total += d;
                    if (total > E.window) break;
                    j -= d;
                    const uint32_t sj = j + sbase;
                    const uint32_t tj = (uint32_t)sb[sj] | ((uint32_t)sb[sj + 1] << 8) | ((uint32_t)sb[sj + 2] << 16);
                    if (tj == t) { found = true; break; }
                    d = lk[j];
                }
                if (found) {
                    const uint32_t a = si + 3, b = j + sbase + 3;
                    const uint32_t limit = min(E.max_len - 3, n - (pos + 3));
                    uint32_t k = 0;
                    while (k < limit) {
                        const uint32_t x = ld32u(sb, a + k) ^ ld32u(sb, b + k);
                        if (x) { k += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
                        k += 4;
                    }
                    if (k > limit) k = limit;
                    out = ((3 + k) << 16) | total;
                }
            }
            md[pos] = out;
        }
    }
}

// =============================================================================== K3 parse_exits
// Greedy walk i -> i + step(i), step = match length or 1.  For a tile [ts,te) and each of the <= 258
// positions a previous tile can jump into, exit = (first position >= te reached) - te  (SURVEY App. C).
constexpr uint32_t kRing = 260;
__global__ void __launch_bounds__(64) k_parse_exits(EncDev E, uint32_t off, uint32_t lim) {
    __shared__ uint16_t ring[64 * kRing];
    const uint32_t tile = blockIdx.x * 64 + threadIdx.x + off;
    if (tile >= lim) return;
    const uint32_t c = find_owner(E.tile0, E.n_chunks, tile);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t n = cd.len;
    const uint32_t ts = (tile - E.tile0[c]) * kTile;
    const uint32_t te = min(ts + kTile, n);
    uint16_t *r = ring + threadIdx.x * kRing;
    const uint32_t *__restrict__ m = E.md + cd.off;
    uint16_t *__restrict__ xt = E.exit_tab + (uint64_t)tile * kExitW;
    for (uint32_t i = te; i-- > ts;) {
        const uint32_t v = m[i];
        const uint32_t nx = i + (v ? (v >> 16) : 1u);
        uint32_t ex;
        if (nx >= te) ex = nx - te; else ex = r[(nx - ts) % kRing];
        r[(i - ts) % kRing] = (uint16_t)ex;
        if (i - ts < kExitW) xt[i - ts] = (uint16_t)ex;
    }
}

// =============================================================================== K4 parse_stitch
__global__ void __launch_bounds__(64) k_parse_stitch(EncDev E, uint32_t off, uint32_t lim) {
    const uint32_t c = blockIdx.x * 64 + threadIdx.x + off;
    if (c >= lim) return;
    const uint32_t t0 = E.tile0[c], t1 = E.tile0[c + 1];
    uint32_t p = 0;
    for (uint32_t t = t0; t < t1; t++) {
        E.tile_entry[t] = (uint16_t)p;
        if (t + 1 < t1) p = E.exit_tab[(uint64_t)t * kExitW + p];
    }
}

// =============================================================================== K5 parse_emit
__global__ void __launch_bounds__(64) k_parse_emit(EncDev E, uint32_t off) {
    __shared__ uint32_t sh[kHistStride];
    const uint32_t grp = blockIdx.x + off;
    const uint32_t c = find_owner(E.grp0, E.n_chunks, grp);
    const ChunkDesc cd = E.chunks[c];
    for (uint32_t i = threadIdx.x; i < kHistStride; i += 64) sh[i] = 0;
    __syncthreads();
    const uint32_t n = cd.len;
    const uint32_t tk = (grp - E.grp0[c]) * kGrpTiles + threadIdx.x;      // tile index inside the chunk
    const uint32_t ntile_c = E.tile0[c + 1] - E.tile0[c];
    if (tk < ntile_c) {
        const uint32_t tile = E.tile0[c] + tk;
        const uint32_t ts = tk * kTile, te = min(ts + kTile, n);
        const uint32_t *__restrict__ m = E.md + cd.off;
        const uint8_t *__restrict__ p = E.in + cd.off;
        uint32_t *__restrict__ so = E.sym + cd.off + ts;
        uint32_t i = ts + E.tile_entry[tile], cnt = 0;
        while (i < te) {
            const uint32_t v = m[i];
            const uint32_t b = p[i];
            if (v) {
                uint32_t lc, le, lx, dc, de, dx;
                length_code(v >> 16, lc, le, lx);
                dist_code(v & 0xFFFFu, dc, de, dx);
                atomicAdd(&sh[lc], 1u);
                atomicAdd(&sh[286 + dc], 1u);
                so[cnt++] = kSymPtr | v;
                i += v >> 16;
            } else {
                atomicAdd(&sh[b], 1u);
                so[cnt++] = b;
                i += 1;
            }
        }
        E.tile_nsym[tile] = cnt;
    }
    __syncthreads();
    uint32_t *__restrict__ gh = E.hist + (uint64_t)cd.block * kHistStride;
    for (uint32_t i = threadIdx.x; i < 316; i += 64) if (sh[i]) atomicAdd(&gh[i], sh[i]);
}

// =============================================================================== K6 huff_build
__global__ void __launch_bounds__(32) k_huff_build(EncDev E) {
    __shared__ HuffWork W;
    const uint32_t b = blockIdx.x;
    if (threadIdx.x != 0) return;
    uint32_t *lit = E.litcode + (uint64_t)b * kLitStride;
    uint32_t *dist = E.distcode + (uint64_t)b * kDistStride;
    if (E.blocks[b].fixed) {
        build_fixed_codes(lit, dist);
        E.hdr_bits[b] = 0;
    } else {
        E.hdr_bits[b] = build_block_codes(E.hist + (uint64_t)b * kHistStride, lit, dist, E.hdr_words + (uint64_t)b * kHdrWords, W);
    }
}

// =============================================================================== symbol -> bits
__device__ __forceinline__ void sym_bits(uint32_t s, const uint32_t *__restrict__ lit, const uint32_t *__restrict__ dist, uint64_t &v, uint32_t &nb) {
    if (s & kSymPtr) {
        uint32_t lc, le, lx, dc, de, dx;
        length_code((s >> 16) & 0x1FFu, lc, le, lx);
        dist_code(s & 0xFFFFu, dc, de, dx);
        const uint32_t cl = __ldg(lit + lc), cdist = __ldg(dist + dc);
        const uint32_t wl = cl >> 16, wd = cdist >> 16;
        v = (uint64_t)(cl & 0xFFFFu) | ((uint64_t)lx << wl) | ((uint64_t)(cdist & 0xFFFFu) << (wl + le)) | ((uint64_t)dx << (wl + le + wd));
        nb = wl + le + wd + de;
    } else {
        const uint32_t cl = __ldg(lit + s);
        v = cl & 0xFFFFu; nb = cl >> 16;
    }
}

// =============================================================================== K7 tile_bits
__global__ void __launch_bounds__(256) k_tile_bits(EncDev E) {
    const uint32_t tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (tile >= E.n_tiles) return;
    const uint32_t c = find_owner(E.tile0, E.n_chunks, tile);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t ts = (tile - E.tile0[c]) * kTile;
    const uint32_t *__restrict__ so = E.sym + cd.off + ts;
    const uint32_t *lit = E.litcode + (uint64_t)cd.block * kLitStride;
    const uint32_t *dist = E.distcode + (uint64_t)cd.block * kDistStride;
    const uint32_t nsym = E.tile_nsym[tile];
    uint32_t sum = 0;
    for (uint32_t k = lane; k < nsym; k += 32) { uint64_t v; uint32_t nb; sym_bits(so[k], lit, dist, v, nb); sum += nb; }
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) E.tile_bits[tile] = sum;
}

// =============================================================================== K8a scan_tiles (CTA per block)
__global__ void __launch_bounds__(256) k_scan_tiles(EncDev E) {
    __shared__ uint64_t wsum[8];
    __shared__ uint64_t carry_s;
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const BlockDesc bd = E.blocks[b];
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < bd.ntiles; base += 256) {
        const uint32_t k = base + tid;
        const uint64_t x = k < bd.ntiles ? E.tile_bits[bd.tile0 + k] : 0;
        uint64_t incl = x;
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        uint64_t woff = 0;
        for (uint32_t w = 0; w < wid; w++) woff += wsum[w];
        const uint64_t carry = carry_s;
        if (k < bd.ntiles) E.tile_bitrel[bd.tile0 + k] = carry + woff + incl - x;
        __syncthreads();
        if (tid == 255) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) {
        const uint32_t eobw = E.litcode[(uint64_t)b * kLitStride + 256] >> 16;
        E.blk_bits[b] = 3ull + E.hdr_bits[b] + carry_s + eobw;
    }
}

// =============================================================================== K8b scan_blocks (thread per stream)
__global__ void __launch_bounds__(64) k_scan_blocks(EncDev E) {
    const uint32_t s = blockIdx.x * 64 + threadIdx.x;
    if (s >= E.n_streams) return;
    uint64_t pos = (uint64_t)E.hdr_len[s] * 8;
    for (uint32_t b = E.stream_blk0[s]; b < E.stream_blk0[s + 1]; b++) {
        E.blk_bitoff[b] = pos;
        pos += E.blk_bits[b];
        if (E.blocks[b].sync_after) {       // empty stored block, byte aligned (encode.rs:225-234)
            pos += 3;
            pos = (pos + 7) & ~7ull;
            E.blk_markpos[b] = pos;
            pos += 32;
        }
    }
    E.stream_end_bits[s] = pos;
}

__device__ __forceinline__ void put_bits_global(uint32_t *__restrict__ out, uint64_t bitpos, uint32_t v, uint32_t nb) {
    if (!nb) return;
    const uint64_t wi = bitpos >> 5; const uint32_t sh = (uint32_t)bitpos & 31u;
    atomicOr(out + wi, v << sh);
    if (sh && sh + nb > 32) atomicOr(out + wi + 1, v >> (32 - sh));
}

// =============================================================================== K9 write_headers (warp per block)
__global__ void __launch_bounds__(128) k_write_headers(EncDev E) {
    const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (b >= E.n_blocks) return;
    const BlockDesc bd = E.blocks[b];
    const uint64_t abs0 = E.out_base[bd.stream] * 8 + E.blk_bitoff[b];
    const uint32_t hb = E.hdr_bits[b];
    if (lane == 0) {
        const uint32_t btype = bd.fixed ? 1u : 2u;
        put_bits_global(E.out_words, abs0, (bd.is_final ? 1u : 0u) | (btype << 1), 3);
        const uint32_t eob = E.litcode[(uint64_t)b * kLitStride + 256];
        put_bits_global(E.out_words, abs0 + E.blk_bits[b] - (eob >> 16), eob & 0xFFFFu, eob >> 16);
        if (bd.sync_after) put_bits_global(E.out_words, E.out_base[bd.stream] * 8 + E.blk_markpos[b] + 16, 0xFFFFu, 16);
    }
    const uint32_t *hw = E.hdr_words + (uint64_t)b * kHdrWords;
    for (uint32_t k = lane; k * 32 < hb; k += 32) {
        const uint32_t nb = min(32u, hb - k * 32);
        uint32_t v = hw[k];
        if (nb < 32) v &= (1u << nb) - 1u;
        put_bits_global(E.out_words, abs0 + 3 + (uint64_t)k * 32, v, nb);
    }
}

// =============================================================================== K10 bitpack (warp per tile)
constexpr uint32_t kPackWords = 1040;
__global__ void __launch_bounds__(256) k_bitpack(EncDev E) {
    extern __shared__ uint32_t pk[];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * 8 + wid;
    if (tile >= E.n_tiles) return;
    const uint32_t nbits = E.tile_bits[tile];
    if (!nbits) return;
    const uint32_t c = find_owner(E.tile0, E.n_chunks, tile);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t b = cd.block;
    const uint32_t ts = (tile - E.tile0[c]) * kTile;
    const uint32_t *__restrict__ so = E.sym + cd.off + ts;
    const uint32_t *lit = E.litcode + (uint64_t)b * kLitStride;
    const uint32_t *dist = E.distcode + (uint64_t)b * kDistStride;
    const uint32_t nsym = E.tile_nsym[tile];
    const uint64_t abs0 = E.out_base[E.blocks[b].stream] * 8 + E.blk_bitoff[b] + 3 + E.hdr_bits[b] + E.tile_bitrel[tile];
    const uint64_t word0 = abs0 >> 5;
    const uint32_t sh0 = (uint32_t)abs0 & 31u;
    const uint32_t nwords = (sh0 + nbits + 31) >> 5;
    uint32_t *buf = pk + wid * kPackWords;
    for (uint32_t w = lane; w < nwords + 2; w += 32) buf[w] = 0;
    __syncwarp();
    uint32_t run = sh0;
    for (uint32_t k0 = 0; k0 < nsym; k0 += 32) {
        const uint32_t k = k0 + lane;
        uint64_t v = 0; uint32_t nb = 0;
        if (k < nsym) sym_bits(so[k], lit, dist, v, nb);
        uint32_t incl = nb;
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
        if (nb) {
            const uint32_t off = run + incl - nb;
            const uint32_t wi = off >> 5, sh = off & 31u;
            atomicOr(buf + wi, (uint32_t)(v << sh));
            const uint64_t rem = v >> (32 - sh);
            if ((uint32_t)rem) atomicOr(buf + wi + 1, (uint32_t)rem);
            if ((uint32_t)(rem >> 32)) atomicOr(buf + wi + 2, (uint32_t)(rem >> 32));
        }
        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncwarp();
    uint32_t *__restrict__ out = E.out_words + word0;
    for (uint32_t w = lane; w < nwords; w += 32) {
        const uint32_t val = buf[w];
        if (w == 0 || w == nwords - 1) { if (val) atomicOr(out + w, val); }
        else out[w] = val;
    }
}

// =============================================================================== symbol compaction (b2f_lz77_default)
__global__ void __launch_bounds__(1024) k_scan_nsym_single(EncDev E, uint64_t *tile_symoff, uint64_t *total) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t carry_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < E.n_tiles; base += 1024) {
        const uint32_t k = base + tid;
        const uint64_t x = k < E.n_tiles ? E.tile_nsym[k] : 0;
        uint64_t incl = x;
        for (int d = 1; d < 32; d <<= 1) { uint64_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
        if (lane == 31) wsum[wid] = incl;
        __syncthreads();
        uint64_t woff = 0;
        for (uint32_t w = 0; w < wid; w++) woff += wsum[w];
        const uint64_t carry = carry_s;
        if (k < E.n_tiles) tile_symoff[k] = carry + woff + incl - x;
        __syncthreads();
        if (tid == 1023) carry_s = carry + woff + incl;
        __syncthreads();
    }
    if (tid == 0) *total = carry_s;
}
__global__ void __launch_bounds__(256) k_compact_syms(EncDev E, const uint64_t *tile_symoff, uint32_t *dst) {
    const uint32_t tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (tile >= E.n_tiles) return;
    const uint32_t c = find_owner(E.tile0, E.n_chunks, tile);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t ts = (tile - E.tile0[c]) * kTile;
    const uint32_t *__restrict__ so = E.sym + cd.off + ts;
    const uint32_t nsym = E.tile_nsym[tile];
    uint32_t *d = dst + tile_symoff[tile];
    for (uint32_t k = lane; k < nsym; k += 32) d[k] = so[k];
}

// =============================================================================== launchers
#define B2F_LAUNCH_CHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return e__; } while (0)

cudaError_t enc_init_attributes() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_lz_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << kHashBits) * 4));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_lz_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_bitpack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kPackWords * 4));
    return e;
}

// One slice = a contiguous range of chunks [c0, c1): chain -> match -> exits -> stitch -> emit, in order, on one stream.
static cudaError_t enc_launch_lz_slice(const EncDev &E, const uint32_t *h_seg0, const uint32_t *h_pt0, const uint32_t *h_tile0, const uint32_t *h_grp0,
                                       uint32_t c0, uint32_t c1, cudaStream_t st, StageTimer *tm) {
    if (c1 <= c0) return cudaSuccess;
    const uint32_t nseg = h_seg0[c1] - h_seg0[c0], npt = h_pt0[c1] - h_pt0[c0], nt = h_tile0[c1] - h_tile0[c0], ng = h_grp0[c1] - h_grp0[c0];
    if (tm) tm->mark(st, "lz_chain");
    k_lz_chain<<<nseg, 32, (1u << kHashBits) * 4, st>>>(E, h_seg0[c0]); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "lz_match");
    k_lz_match<<<npt, 512, kMatchSmem, st>>>(E, h_pt0[c0]); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "parse_exits");
    k_parse_exits<<<(nt + 63) / 64, 64, 0, st>>>(E, h_tile0[c0], h_tile0[c1]); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "parse_stitch");
    k_parse_stitch<<<(c1 - c0 + 63) / 64, 64, 0, st>>>(E, c0, c1); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "parse_emit");
    k_parse_emit<<<ng, 64, 0, st>>>(E, h_grp0[c0]); B2F_LAUNCH_CHECK();
    return cudaSuccess;
}

// LZ77 stage.  With aux streams the chunks are split into slices that run concurrently: the chain / exits kernels are
// latency bound (one warp per segment / thread per tile) and leave most issue slots free for the match kernel of another slice.
cudaError_t enc_launch_lz(const EncDev &E, const uint32_t *h_seg0, const uint32_t *h_pt0, const uint32_t *h_tile0, const uint32_t *h_grp0,
                          cudaStream_t st, StageTimer *tm, cudaStream_t *aux, cudaEvent_t *ev, uint32_t n_aux, const SliceFeed *feed) {
    if (E.n_chunks == 0) return cudaSuccess;
    if (n_aux < 2 || E.n_chunks < 2 * n_aux) {
        if (feed) { if (tm) tm->mark(st, "h2d"); cudaError_t fe = feed->copy(feed->self, 0, E.n_chunks, st); if (fe != cudaSuccess) return fe; }
        return enc_launch_lz_slice(E, h_seg0, h_pt0, h_tile0, h_grp0, 0, E.n_chunks, st, tm);
    }
    if (tm) tm->mark(st, "lz_pipeline");
    cudaError_t e = cudaEventRecord(ev[0], st); if (e != cudaSuccess) return e;
    // slices balanced by match tiles (proportional to bytes)
    const uint32_t total = h_pt0[E.n_chunks];
    uint32_t c0 = 0;
    for (uint32_t g = 0; g < n_aux; g++) {
        uint32_t c1 = c0;
        const uint32_t want = (uint32_t)((uint64_t)total * (g + 1) / n_aux);
        while (c1 < E.n_chunks && (h_pt0[c1] < want || g + 1 == n_aux)) c1++;
        e = cudaStreamWaitEvent(aux[g], ev[0], 0); if (e != cudaSuccess) return e;
        if (feed) { e = feed->copy(feed->self, c0, c1, aux[g]); if (e != cudaSuccess) return e; }   // this slice's H2D overlaps the previous slices' kernels
        e = enc_launch_lz_slice(E, h_seg0, h_pt0, h_tile0, h_grp0, c0, c1, aux[g], nullptr); if (e != cudaSuccess) return e;
        e = cudaEventRecord(ev[1 + g], aux[g]); if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(st, ev[1 + g], 0); if (e != cudaSuccess) return e;
        c0 = c1;
    }
    return cudaSuccess;
}
cudaError_t enc_launch_entropy(const EncDev &E, cudaStream_t st, StageTimer *tm) {
    if (tm) tm->mark(st, "huff_build");
    k_huff_build<<<E.n_blocks, 32, 0, st>>>(E); B2F_LAUNCH_CHECK();
    if (E.n_tiles) {
        if (tm) tm->mark(st, "tile_bits");
        k_tile_bits<<<(E.n_tiles + 7) / 8, 256, 0, st>>>(E); B2F_LAUNCH_CHECK();
    }
    if (tm) tm->mark(st, "scan");
    k_scan_tiles<<<E.n_blocks, 256, 0, st>>>(E); B2F_LAUNCH_CHECK();
    k_scan_blocks<<<(E.n_streams + 63) / 64, 64, 0, st>>>(E); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "write_headers");
    k_write_headers<<<(E.n_blocks + 3) / 4, 128, 0, st>>>(E); B2F_LAUNCH_CHECK();
    if (E.n_tiles) {
        if (tm) tm->mark(st, "bitpack");
        k_bitpack<<<(E.n_tiles + 7) / 8, 256, 8 * kPackWords * 4, st>>>(E); B2F_LAUNCH_CHECK();
    }
    return cudaSuccess;
}
cudaError_t enc_launch_compact(const EncDev &E, uint64_t *tile_symoff, uint64_t *total, uint32_t *dst, cudaStream_t st) {
    k_scan_nsym_single<<<1, 1024, 0, st>>>(E, tile_symoff, total); B2F_LAUNCH_CHECK();
    k_compact_syms<<<(E.n_tiles + 7) / 8, 256, 0, st>>>(E, tile_symoff, dst); B2F_LAUNCH_CHECK();
    return cudaSuccess;
}
uint32_t enc_launch_count_lz() { return 5; }
uint32_t enc_launch_count_entropy(bool has_tiles) { return has_tiles ? 6 : 4; }

}  // namespace b2f

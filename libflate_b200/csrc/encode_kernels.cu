// encode_kernels.cu -- the DEFLATE encode hot path as hand-written sm_100a kernels.
//
// Pipeline (one launch each, all on the ctx stream; see DESIGN.md for bytes/unit and rooflines):
//   K1 lz_find       CTA per 256 KiB segment: ordered hash-table update handed from warp to warp, chain walks + LCP out of
//                    shared-memory rings; lz_fixup / lz_fixup2 finish the few positions with very long chains    (E2/A2/L1)
//      (lz_chain + lz_match: the earlier two-kernel form of the same step, B2F_LZ_FUSED=0)
//   K3 parse_exits   thread per 2 KiB tile: right-to-left exit DP of the greedy walk            (E2)
//   K4 parse_stitch  thread per chunk: chain the tile entry points
//   K5 parse_emit    thread per tile: walk, emit symbols, per-block histograms                  (S1/S2)
//   K6 huff_build    one warp per DEFLATE block: code lengths, canonical codes, header bits     (H1-H3/S3/S4)
//   K7 tile_bits     warp per tile: coded size of the tile
//   K8 scan_tiles / scan_blocks: bit offsets of every tile / block / stream                     (B2, E1)
//   K9 write_headers warp per block: BFINAL/BTYPE, dynamic header, EOB, sync marker             (B2)
//   K10 bitpack      warp per tile: LSB-first packing in shared memory, coalesced store         (B1)
// Reference behaviour: libflate_lz77/src/default.rs:59-183, src/deflate/encode.rs:261-426,
// src/deflate/symbol.rs:95-183,321-386,486-540, src/huffman.rs:35-55,192-363, src/bit.rs:25-49.
#include <stdlib.h>
#include "common.cuh"
#include "huff_build.cuh"
#include "huff_warp.cuh"
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

// md[p] = the byte itself (literal; length field 0) or length<<16 | distance of the single candidate libflate would take at p:
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
            uint32_t out = sb[pos + sbase];                       // literal: the byte itself (length field 0)
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

// =============================================================================== K1+K2 fused: lz_find
// One CTA per chain segment does what k_lz_chain + k_lz_match do, with every intermediate in shared memory:
//   * the bytes and the hash-chain links of the last 32 KiB (+ kFindCap blocks of slack) live in two rings, so a chain hop is
//     a shared-memory read instead of a dependent L2 load (47 % of k_lz_match's stall samples);
//   * one loader warp streams the segment into the byte ring, 128-byte blocks, several loads in flight;
//   * the worker warps claim blocks dynamically (atomic counter).  Only the hash-table update is ordered: a worker runs its
//     four atomicMax steps when the turn (an mbarrier per block slot) reaches its block -- then every earlier position has been
//     inserted and no later one, exactly the single-warp invariant of k_lz_chain -- publishes the block's links and passes the
//     turn on.  Walking the chains, the LCP and the md[] store run outside the ordered section, concurrently on all workers;
//     because blocks are claimed by whichever warp is free, a slow block delays nobody else until the ring slack is used up.
//   * the tails are cut so that no block is slow: a chain walk stops after kFindHops hops (0.3 % of the positions of
//     titles-shaped text, but they were half of the walk time) and the position goes to a queue that k_lz_fixup finishes from
//     HBM (link[] is written through for it); a match longer than 19 bytes is extended by the whole warp, 128 bytes per round,
//     once per run of consecutive positions that share the same distance (their lengths differ by one each).
// Ring safety: a worker publishes the block it is working on; it only reads positions >= 128 b - 32768.  The loader does not
// overwrite a ring slot before every published block (and hence every future one) is past the slot's old content:
// block x may be staged once min(cur_blk) >= x + 1 - kFindCap.  The link writes of block b reuse older slots than the bytes
// of block b + 3 staged before it, so the same test covers them.
#ifndef B2F_FIND_WARPS
#define B2F_FIND_WARPS 31
#endif
#ifndef B2F_FIND_TURN
#define B2F_FIND_TURN 1
#endif
#ifndef B2F_FIND_LOADDEPTH
#define B2F_FIND_LOADDEPTH 8
#endif
#ifndef B2F_FIND_CAP
#define B2F_FIND_CAP (B2F_FIND_TURN > 1 ? 72 : 64)
#endif
#ifndef B2F_FIND_HOPS
#define B2F_FIND_HOPS 16
#endif
#ifndef B2F_FIND_UNROLL
#define B2F_FIND_UNROLL 4
#endif
#ifndef B2F_FIND_WAIT_NS
#define B2F_FIND_WAIT_NS 20000u
#endif
constexpr int kFindUnroll = B2F_FIND_UNROLL;                          // 1: one copy of the match code, 4: one per step of a block
constexpr uint32_t kFindWarps = B2F_FIND_WARPS;                    // worker warps; one more warp loads
constexpr uint32_t kFindCap = B2F_FIND_CAP;
constexpr uint32_t kFindTurn = B2F_FIND_TURN;                      // 128-byte blocks per claim = per turn of the ordered table update
constexpr uint32_t kFindSteps = 4 * kFindTurn;                     // warp steps (32 positions) per turn
constexpr uint32_t kFindHops = B2F_FIND_HOPS;                      // chain hops per position before it is deferred to k_lz_fixup
constexpr uint32_t kFindSlots = 32;                                // turn mbarriers (> kFindWarps: at most kFindWarps blocks wait for their turn)
constexpr uint32_t kFindRing = kLookback + 128 * kFindCap;
constexpr uint32_t kFindRingBlocks = kFindRing / 128;
constexpr uint32_t kFindMirror = 384;                              // the first bytes of the ring repeated after its end: forward reads never wrap
constexpr uint32_t kFindLoadDepth = B2F_FIND_LOADDEPTH;                             // blocks per loader step (two steps in flight)
constexpr uint32_t kFindLane = 16;                                 // bytes of a match compared by its own lane before the warp takes over
constexpr uint32_t kFindCtrl = (4u << kHashBits) + 2 * kFindRing + kFindRing + kFindMirror;   // offset of the control words
constexpr uint32_t kFindQueue = kFindCtrl + 16 + 4 * 32 + 8 * 2 * kFindSlots + 4 * 32;   // ctrl | cur_blk | turn + publication barriers | dummies
constexpr uint32_t kFindSmem = kFindQueue + 4 * 128 * kFindTurn * kFindWarps;         // | per-warp queue of the positions that need more than one hop
static_assert(kFindWarps >= 1 && kFindWarps < kFindSlots && kFindCap >= 16 && kFindMirror >= 3 + 258 + 31 + 8 && kFindMirror % 128 == 0, "lz_find geometry");
static_assert(kFindRingBlocks % kFindTurn == 0 && kFindCap >= kFindTurn * kFindWarps + kFindTurn + 3 && kFindSmem <= 232448 && kFindTurn <= 2, "lz_find geometry");

__device__ __forceinline__ uint32_t lds_acquire(uint32_t saddr) {
    uint32_t v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory"); return v;
}
__device__ __forceinline__ void sts_release(uint32_t saddr, uint32_t v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
// explicit shared-space loads with 32-bit addresses (generic pointers make the compiler rebuild the window base in the loops)
__device__ __forceinline__ uint32_t s_ld32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t s_ld16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t s_ld32u(uint32_t base_a, uint32_t off) {     // unaligned 4-byte read at base + off
    const uint32_t a = base_a + (off & ~3u);
    return __funnelshift_r(s_ld32(a), s_ld32(a + 4u), (off & 3u) * 8u);
}
// mbarrier hand-off of the turn: slot b mod kFindSlots completes one phase per block; the warp that finished the ordered section
// of block b-1 arrives on it (release), the owner of block b waits for it in hardware (acquire) instead of polling.
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        // the suspend-time hint parks the warp in hardware; without it the waiters poll and take the issue slots the turn holder needs
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity), "r"(B2F_FIND_WAIT_NS) : "memory");
    } while (!ok);
}

#ifdef B2F_FIND_MAXREG
__global__ void __maxnreg__(B2F_FIND_MAXREG) k_lz_find(EncDev E, uint32_t off, uint32_t slice) {
#else
__global__ void __launch_bounds__((kFindWarps + 1) * 32) k_lz_find(EncDev E, uint32_t off, uint32_t slice) {
#endif
    extern __shared__ __align__(16) uint8_t fsm[];
    uint32_t *head = reinterpret_cast<uint32_t *>(fsm);                            // last position + 1 per hash bucket
    uint16_t *lring = reinterpret_cast<uint16_t *>(fsm + (4u << kHashBits));       // link of position q at q mod kFindRing
    uint8_t *bring = fsm + (4u << kHashBits) + 2 * kFindRing;                      // byte of position q at q mod kFindRing (+ mirror)
    uint32_t *ctrl = reinterpret_cast<uint32_t *>(fsm + kFindCtrl);                // [0] next block to claim, [1] blocks staged, [4..36) cur_blk
    const uint32_t ctrl_a = (uint32_t)__cvta_generic_to_shared(ctrl);
    const uint32_t staged_a = ctrl_a + 4, cur_a = ctrl_a + 16, mb_a = ctrl_a + 16 + 4 * 32, dummy_a = mb_a + 8 * 2 * kFindSlots;
    const uint32_t head_a = (uint32_t)__cvta_generic_to_shared(head), queue_a = head_a + kFindQueue;
    const uint32_t lring_a = (uint32_t)__cvta_generic_to_shared(lring), bring_a = (uint32_t)__cvta_generic_to_shared(bring);
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t seg = blockIdx.x + off;
    const uint32_t c = find_owner(E.seg0, E.n_chunks, seg);
    const ChunkDesc cd = E.chunks[c];
    const uint32_t n = cd.len;
    const uint32_t end = (n > 3 ? n : 3) - 3;
    const uint32_t s_start = (seg - E.seg0[c]) * kSeg;
    const uint32_t s_end = min(s_start + kSeg, n);
    const uint32_t lim = min(s_end, end);
    const uint32_t ws = s_start > kLookback ? s_start - kLookback : 0;
    // ring coordinates: q = position - a0, a0 = warm-up start rounded down so that block loads are 4-byte aligned words
    const int64_t a0 = (int64_t)ws - (int64_t)((cd.off + ws) & 3u);
    const uint64_t g0 = (uint64_t)((int64_t)cd.off + a0);
    const uint32_t nblk = (uint32_t)(((int64_t)s_end - a0 + 127) >> 7);
    const uint32_t n_staged = (nblk + 3 + kFindLoadDepth - 1) / kFindLoadDepth * kFindLoadDepth;   // what the loader stages in all
    for (uint32_t i = threadIdx.x; i < (1u << kHashBits); i += (kFindWarps + 1) * 32) head[i] = 0;
    if (threadIdx.x < 36) ctrl[threadIdx.x] = 0;
    if (threadIdx.x < 32) ctrl[36 + 4 * kFindSlots + threadIdx.x] = 0;             // dummy words
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < 2 * kFindSlots; i++) mbar_init(mb_a + 8u * i, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) { mbar_arrive(mb_a); mbar_arrive(mb_a + 8u * kFindSlots); }   // block 0 may start / publish

    if (w == kFindWarps) {
        // ------------------------------------------------------------------ loader warp
        const uint32_t nstage = nblk + 3;                                          // the LCP of the last block reads up to 300 bytes ahead
        uint32_t cur[kFindLoadDepth], nxt[kFindLoadDepth];
#pragma unroll
        for (uint32_t k = 0; k < kFindLoadDepth; k++) cur[k] = ld_in32(E.in, g0 + 128ull * k + 4ull * lane, E.in_size);
        uint32_t rb = 0;                                                           // ring index of block x
        for (uint32_t x = 0; x < nstage; x += kFindLoadDepth) {
#pragma unroll
            for (uint32_t k = 0; k < kFindLoadDepth; k++) nxt[k] = ld_in32(E.in, g0 + 128ull * (x + kFindLoadDepth + k) + 4ull * lane, E.in_size);
            const uint32_t last = x + kFindLoadDepth;                              // blocks [x, last) are staged by this step
            if (last > kFindCap) {                                                 // wait until nobody needs the slots about to be reused
                const uint32_t need = last - kFindCap;
                for (;;) {
                    const uint32_t v = lane < kFindWarps ? lds_acquire(cur_a + 4u * lane) : 0xFFFFFFFFu;
                    if (__reduce_min_sync(0xFFFFFFFFu, v) >= need) break;
                    __nanosleep(200);
                }
            }
#pragma unroll
            for (uint32_t k = 0; k < kFindLoadDepth; k++) {
                *reinterpret_cast<uint32_t *>(bring + rb + 4u * lane) = cur[k];
                if (rb < kFindMirror) *reinterpret_cast<uint32_t *>(bring + kFindRing + rb + 4u * lane) = cur[k];
                rb += 128u; if (rb >= kFindRing) rb = 0;
            }
            __syncwarp();
            if (lane == 0) sts_release(staged_a, last);
#pragma unroll
            for (uint32_t k = 0; k < kFindLoadDepth; k++) cur[k] = nxt[k];
        }
        return;
    }

    // ---------------------------------------------------------------------- worker warps
    uint32_t *__restrict__ md = E.md + cd.off;
    uint16_t *__restrict__ glk = E.link + cd.off;
    for (;;) {
        uint32_t tc = 0;                                                           // turn = kFindTurn consecutive blocks
        if (lane == 0) tc = atomicAdd(&ctrl[0], 1u);
        tc = __shfl_sync(0xFFFFFFFFu, tc, 0);
        const uint32_t b = tc * kFindTurn;                                         // first block of the turn
        if (b >= nblk) break;
        if (lane == 0) sts_release(cur_a + 4u * w, b);                             // from now on this warp reads positions >= 128 b - 32768 only
        const int32_t bpos = (int32_t)a0 + (int32_t)(128u * b);                    // chunk position of the turn's first byte (>= -3)
        const uint32_t rb = (b % kFindRingBlocks) * 128u;                          // ring index of the turn (a turn never wraps)
        const uint32_t need = min(b + kFindTurn + 3u, n_staged);                   // the turn's blocks and the three after them are in the ring
        while (lds_acquire(staged_a) < need) { }
        // Everything the ordered section needs is prepared first: bucket address and pos+1 per position.  Positions that are not
        // inserted (alignment lead-in, the last three bytes of the chunk, beyond the segment) go to a per-lane dummy word with
        // value 0, so the section is four unconditional shared-memory atomics.
        uint32_t tg[kFindSteps], hh[kFindSteps], ha[kFindSteps], p1[kFindSteps], oo[kFindSteps], dd[kFindSteps];
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
            const int32_t pos = bpos + (int32_t)(32 * s4 + lane);
            const uint32_t t = s_ld32u(bring_a, rb + 32 * s4 + lane) & 0xFFFFFFu;
            const bool v = pos >= (int32_t)ws && pos < (int32_t)lim;
            tg[s4] = t;
            hh[s4] = (t * 0x9E3779B1u) >> (32 - kHashBits);
            ha[s4] = v ? head_a + 4u * hh[s4] : dummy_a + 4u * lane;
            p1[s4] = v ? (uint32_t)pos + 1u : 0u;
        }
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) asm volatile("" :: "r"(ha[s4]), "r"(p1[s4]));
        // ---- stage A, ordered: the table update.  The turn is passed on as soon as the atomics are issued (the arrive is a release,
        // so they are performed before the next owner's); everything that only consumes their results comes after.
#ifndef B2F_FIND_NOCHAIN
        mbar_wait(mb_a + 8u * (tc & (kFindSlots - 1)), (tc / kFindSlots) & 1u);
#endif
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) asm volatile("atom.shared.max.u32 %0, [%1], %2;" : "=r"(oo[s4]) : "r"(ha[s4]), "r"(p1[s4]) : "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(mb_a + 8u * ((tc + 1) & (kFindSlots - 1)));
        // ---- stage B: links from the returned heads
        bool bad = false;
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
#ifdef B2F_CHAIN_FORCE_REPAIR
            bad = true;
#else
            bad |= p1[s4] != 0u && oo[s4] >= p1[s4];                               // saw a later position of its own step: out-of-order serialisation
#endif
        }
        if (__any_sync(0xFFFFFFFFu, bad)) {                                        // exact repair, see k_lz_chain
#pragma unroll
            for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
                const uint32_t key = p1[s4] ? hh[s4] : (0x80000000u | lane);
                const uint32_t m = __match_any_sync(0xFFFFFFFFu, key);
                const uint32_t lower = m & ((1u << lane) - 1u);
                uint32_t mn = oo[s4], mm = m & ~(1u << lane);
                while (__any_sync(0xFFFFFFFFu, mm != 0)) {
                    const uint32_t src = mm ? (uint32_t)__ffs((int)mm) - 1u : lane;
                    const uint32_t v = __shfl_sync(0xFFFFFFFFu, oo[s4], src);
                    if (mm) { mn = min(mn, v); mm &= mm - 1u; }
                }
                oo[s4] = lower ? p1[s4] - (lane - (31u - (uint32_t)__clz((int)lower))) : mn;
            }
        }
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
            uint32_t d = (p1[s4] && oo[s4]) ? p1[s4] - oo[s4] : 0u;
            if (d > kLookback) d = 0;
            dd[s4] = d;
            lring[rb + 32 * s4 + lane] = (uint16_t)d;
        }
        __syncwarp();
        // links are published in block order (second, much shorter hand-off chain): once this warp holds the publication turn, the
        // links of every earlier block are in the ring
#ifndef B2F_FIND_NOCHAIN
        mbar_wait(mb_a + 8u * (kFindSlots + (tc & (kFindSlots - 1))), (tc / kFindSlots) & 1u);
#endif
        if (lane == 0) mbar_arrive(mb_a + 8u * (kFindSlots + ((tc + 1) & (kFindSlots - 1))));
        // ---- skip links.  Where the first node of a position's chain (its candidate c) has the position's own trigram, a walk that
        // arrives at the position looking for ANOTHER trigram may go on from c's link at once: link := d + link(c).  Any link that
        // only skips nodes of the node's own trigram keeps every walk exact, so the upgrade needs no ordering -- a walker reads the
        // plain link or the upgraded one (16-bit stores do not tear); done right after publication, nearly every node a later walk
        // meets is already upgraded.  Runs of a frequent trigram that share a bucket with rare ones -- chains of thousands of nodes,
        // the reason k_lz_fixup exists -- collapse to one hop (deferred positions on titles-shaped text: 0.32 % -> 0.0x %).
#pragma unroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
            const uint32_t ri = rb + 32 * s4 + lane, d = dd[s4];
            const uint32_t jx = ri >= d ? ri - d : ri + kFindRing - d;
            const uint32_t tj = s_ld32u(bring_a, jx) & 0xFFFFFFu;
            const uint32_t dn = s_ld16(lring_a + 2u * jx);
            const uint32_t up = d + dn;
            if (d != 0u && tj == tg[s4]) asm volatile("st.shared.u16 [%0], %1;" :: "r"(lring_a + 2u * ri), "h"((uint16_t)((dn != 0u && up <= kLookback) ? up : 0u)) : "memory");
        }
        // ---- unordered part: the matches of this block's positions that belong to the segment (warp-uniform control flow)
        if (bpos + (int32_t)(128u * kFindTurn) - 1 < (int32_t)s_start) continue;
#ifdef B2F_FIND_NOMATCH
        continue;                                                                  // timing experiment: the ordered chain alone
#endif
        // Phase 1, straight-line per step: the first chain hop and the first kFindLane bytes of the comparison are ONE pass over
        // aligned words of the position's and the candidate's bytes (five loads + four funnel shifts each): the candidate is the
        // match iff the first three bytes agree (93 % of the positions that have a link at all), and the same XOR words give the
        // length.  Positions whose first hop lands on another trigram of the bucket and has a successor (6 %) go to this warp's
        // queue and are finished in phase 2 as a dense set of lanes -- the divergent walk no longer holds 32 lanes for one.
        uint32_t qn = 0;
        const uint32_t q_a = queue_a + 4u * 128u * kFindTurn * w;
        uint32_t *const md_l = md + ((int64_t)bpos + (int64_t)lane);               // this lane's first position of the turn: the stores of the
        uint16_t *const glk_l = glk + ((int64_t)bpos + (int64_t)lane);             // steps are immediate offsets from these
#pragma unroll kFindUnroll
        for (uint32_t s4 = 0; s4 < kFindSteps; s4++) {
            const int32_t posi = bpos + (int32_t)(32 * s4 + lane);
            const bool valid = posi >= (int32_t)s_start && posi < (int32_t)s_end;
            if (!__any_sync(0xFFFFFFFFu, valid)) continue;
            const uint32_t pos = (uint32_t)posi;
            const uint32_t ri = rb + 32 * s4 + lane;
            const uint32_t d0 = dd[s4];                                            // the plain link: the candidate
            const bool linked = valid && d0 != 0u && d0 <= E.window;
            const uint32_t jx = ri >= d0 ? ri - d0 : ri + kFindRing - d0;          // d0 == 0: the position itself (unused)
            const uint32_t oa = bring_a + (ri & ~3u), osh = (ri & 3u) * 8u;        // forward reads run into the mirror instead of wrapping
            const uint32_t ca = bring_a + (jx & ~3u), csh = (jx & 3u) * 8u;
            const uint32_t ow0 = s_ld32(oa), ow1 = s_ld32(oa + 4u), ow2 = s_ld32(oa + 8u), ow3 = s_ld32(oa + 12u), ow4 = s_ld32(oa + 16u);
            const uint32_t cw0 = s_ld32(ca), cw1 = s_ld32(ca + 4u), cw2 = s_ld32(ca + 8u), cw3 = s_ld32(ca + 12u), cw4 = s_ld32(ca + 16u);
            const uint32_t dn = s_ld16(lring_a + 2u * jx);
            const uint32_t o0 = __funnelshift_r(ow0, ow1, osh);
            const uint32_t x0 = o0 ^ __funnelshift_r(cw0, cw1, csh);
            const uint32_t x1 = __funnelshift_r(ow1, ow2, osh) ^ __funnelshift_r(cw1, cw2, csh);
            const uint32_t x2 = __funnelshift_r(ow2, ow3, osh) ^ __funnelshift_r(cw2, cw3, csh);
            const uint32_t x3 = __funnelshift_r(ow3, ow4, osh) ^ __funnelshift_r(cw3, cw4, csh);
            const bool found = linked && (x0 & 0xFFFFFFu) == 0u;
            const uint32_t sk = s_ld16(lring_a + 2u * ri);                         // own link after the upgrade
            if (valid) glk_l[32 * s4] = (uint16_t)sk;                              // written through for k_lz_fixup
            const bool more = linked && !found && dn != 0u && kFindHops > 1;
            uint32_t xf = x0, mb = 0;                                              // equal bytes from the position itself, 0..kFindLane
            if (!x0) { xf = x1; mb = 4; if (!x1) { xf = x2; mb = 8; if (!x2) { xf = x3; mb = 12; } } }
            uint32_t k = xf ? mb + ((uint32_t)(__ffs((int)xf) - 1) >> 3) : kFindLane;
            const uint32_t limit = min(E.max_len, n - pos);                        // longest match the position may take
            const bool open = found && k == kFindLane && limit > kFindLane;
            // longer matches: one cooperative extension per run of consecutive lanes with the same distance
            const uint32_t U = __ballot_sync(0xFFFFFFFFu, open);
            if (U) {
                const uint32_t pd = __shfl_up_sync(0xFFFFFFFFu, d0, 1);
                const bool follower = open && lane > 0 && ((U >> (lane - 1)) & 1u) && pd == d0;
                const uint32_t F = __ballot_sync(0xFFFFFFFFu, follower);
                const uint32_t H = U & ~F;
                uint32_t ext = 0;                                                  // head lanes: matching bytes from their position (not capped by max_len)
                for (uint32_t hm = H; hm; hm &= hm - 1u) {
                    const uint32_t h = (uint32_t)__ffs((int)hm) - 1u;
                    const uint32_t ah = __shfl_sync(0xFFFFFFFFu, ri, h), sh = __shfl_sync(0xFFFFFFFFu, jx, h), ph = __shfl_sync(0xFFFFFFFFu, pos, h);
                    const uint32_t fr = ~((F >> h) >> 1);                          // followers directly after h
                    const uint32_t run = h == 31 ? 0u : (fr ? (uint32_t)__ffs((int)fr) - 1u : 31u - h);
                    const uint32_t lim_ext = min(E.max_len + run, n - ph);
                    uint32_t e = kFindLane, eh;
                    for (;;) {
                        const uint32_t o = e + 4u * lane;
                        const uint32_t x = o < lim_ext ? (s_ld32u(bring_a, ah + o) ^ s_ld32u(bring_a, sh + o)) : 1u;
                        const uint32_t mm = __ballot_sync(0xFFFFFFFFu, x != 0);
                        if (mm) {
                            const uint32_t l0 = (uint32_t)__ffs((int)mm) - 1u;
                            const uint32_t xl = __shfl_sync(0xFFFFFFFFu, x, l0);
                            const uint32_t ol = e + 4u * l0;
                            eh = ol >= lim_ext ? lim_ext : min(lim_ext, ol + ((uint32_t)(__ffs((int)xl) - 1) >> 3));
                            break;
                        }
                        e += 128u;
                    }
                    if (lane == h) ext = eh;
                }
                const uint32_t below = H & (0xFFFFFFFFu >> (31u - lane));
                const uint32_t hl = (open && below) ? 31u - (uint32_t)__clz((int)below) : lane;
                const uint32_t eh = __shfl_sync(0xFFFFFFFFu, ext, hl);
                if (open) k = eh - (lane - hl);
            }
            uint32_t out = o0 & 0xFFu;                                             // literal: the byte itself (length field 0)
            if (found) out = (min(k, limit) << 16) | d0;
            if (valid && !more) md_l[32 * s4] = out;
            const uint32_t qm = __ballot_sync(0xFFFFFFFFu, more);
            if (more) asm volatile("st.shared.u32 [%0], %1;" :: "r"(q_a + 4u * (qn + (uint32_t)__popc(qm & ((1u << lane) - 1u)))), "r"((32u * s4 + lane) | (d0 << 8)) : "memory");
            qn += (uint32_t)__popc(qm);
        }
        // Phase 2: the queued positions, one per lane: the rest of the walk (at most kFindHops - 1 more hops), the comparison,
        // or the hand-over to k_lz_fixup.
        __syncwarp();
#ifdef B2F_FIND_NOQUEUE
        qn = 0;                                                                    // timing experiment: phase 1 alone
#endif
        for (uint32_t q0 = 0; q0 < qn; q0 += 32) {
            const bool has = q0 + lane < qn;
            const uint32_t ent = has ? s_ld32(q_a + 4u * (q0 + lane)) : 0u;
            const uint32_t ix = ent & 255u;
            const uint32_t pos = (uint32_t)(bpos + (int32_t)ix), ri = rb + ix;
            const uint32_t tt = s_ld32u(bring_a, ri) & 0xFFFFFFu;
            uint32_t total = ent >> 8, hops = kFindHops - 1, res = 0;
            uint32_t jx = ri >= total ? ri - total : ri + kFindRing - total;
            uint32_t d = has ? s_ld16(lring_a + 2u * jx) : 0u;
            // res: 0 no match, 1 found at ring index jx (distance total), 2 deferred
            while (d) {
                total += d;
                if (total > E.window) break;
                jx = jx >= d ? jx - d : jx + kFindRing - d;
                const uint32_t dn = s_ld16(lring_a + 2u * jx);
                const uint32_t tj = s_ld32u(bring_a, jx) & 0xFFFFFFu;
                if (tj == tt) { res = 1; break; }
                d = dn;
                if (--hops == 0) { if (d) res = 2; break; }
            }
            // deferred positions go to the fix-up queue of this slice (or, should it be full, are finished here)
            bool deferred = res == 2;
            const uint32_t dm = __ballot_sync(0xFFFFFFFFu, deferred);
            if (dm) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(E.fix_count + slice, (uint32_t)__popc(dm));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                const uint32_t slot = base + (uint32_t)__popc(dm & ((1u << lane) - 1u));
                if (deferred) {
                    if (slot < E.fix_cap) E.fix_pos[(uint64_t)slice * E.fix_cap + slot] = cd.off + pos;
                    else {
                        while (d) {
                            total += d;
                            if (total > E.window) break;
                            jx = jx >= d ? jx - d : jx + kFindRing - d;
                            const uint32_t dn = s_ld16(lring_a + 2u * jx);
                            const uint32_t tj = s_ld32u(bring_a, jx) & 0xFFFFFFu;
                            if (tj == tt) { res = 1; break; }
                            d = dn;
                        }
                        deferred = false;
                    }
                }
            }
            uint32_t out = tt & 0xFFu;
            if (res == 1) {
                const uint32_t limit = min(E.max_len, n - pos);
                uint32_t k = 3;
                while (k < limit) {
                    const uint32_t x = s_ld32u(bring_a, ri + k) ^ s_ld32u(bring_a, jx + k);
                    if (x) { k += (uint32_t)(__ffs((int)x) - 1) >> 3; break; }
                    k += 4;
                }
                out = (min(k, limit) << 16) | total;
            }
            if (has && !deferred) md[pos] = out;
        }
        __syncwarp();
    }
    if (lane == 0) sts_release(cur_a + 4u * w, 0xFFFFFFFFu);
}

// Finishes the positions whose chain walk k_lz_find cut short.  Level 1: one thread per position repeats the walk from HBM
// (link[], input bytes) for at most kFixHops hops.  What is still open then sits behind a very long chain -- thousands of
// entries when a run of identical lines filled a bucket -- and goes to level 2, which no longer follows the chain: a warp scans
// the rest of the window for the trigram itself, 128 positions per step.
#ifndef B2F_FIX_HOPS
#define B2F_FIX_HOPS 160
#endif
constexpr uint32_t kFixHops = B2F_FIX_HOPS;
__device__ __forceinline__ uint32_t fix_chunk(const EncDev &E, uint64_t g) {       // chunk containing g (chunks lie in increasing offset order)
    uint32_t lo = 0, hi = E.n_chunks;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (E.chunks[mid].off <= g) lo = mid; else hi = mid; }
    return lo;
}
__global__ void __launch_bounds__(256) k_lz_fixup(EncDev E, uint32_t slice) {
    const uint32_t cnt = min(E.fix_count[slice], E.fix_cap);
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < cnt; i += gridDim.x * 256) {
        const uint64_t g = E.fix_pos[(uint64_t)slice * E.fix_cap + i];
        const ChunkDesc cd = E.chunks[fix_chunk(E, g)];
        const uint8_t *__restrict__ p = E.in + cd.off;
        const uint16_t *__restrict__ lk = E.link + cd.off;
        const uint32_t n = cd.len, pos = (uint32_t)(g - cd.off);
        const uint32_t t = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16);
        uint32_t d = lk[pos], total = 0, j = pos, out = t & 0xFFu, hops = 0;     // literal unless a match is found
        bool found = false, pushed = false;
        while (d) {
            total += d;
            if (total > E.window) break;
            j -= d;
            const uint32_t tj = (uint32_t)p[j] | ((uint32_t)p[j + 1] << 8) | ((uint32_t)p[j + 2] << 16);
            if (tj == t) { found = true; break; }
            d = lk[j];
            if (++hops == kFixHops && d) {
                const uint32_t slot = atomicAdd(E.fix2_count + slice, 1u);
                if (slot < E.fix_cap) { E.fix2_pos[(uint64_t)slice * E.fix_cap + slot] = g; E.fix2_j[(uint64_t)slice * E.fix_cap + slot] = j; pushed = true; break; }
            }
        }
        if (pushed) continue;
        if (found) {
            const uint32_t limit = min(E.max_len - 3, n - (pos + 3));
            uint32_t k = 0;
            while (k < limit && p[pos + 3 + k] == p[j + 3 + k]) k++;
            out = ((3 + k) << 16) | total;
        }
        E.md[g] = out;
    }
}
// 32-bit little-endian load at any byte offset, aligned words only, zero fill at and beyond in_size
__device__ __forceinline__ uint32_t ld_in32_any(const uint8_t *__restrict__ in, uint64_t off, uint64_t in_size) {
    const uint64_t a = off & ~3ull;
    const uint32_t lo = ld_in32(in, a, in_size), hi = ld_in32(in, a + 4, in_size);
    return __funnelshift_r(lo, hi, ((uint32_t)off & 3u) * 8u);
}
__global__ void __launch_bounds__(256) k_lz_fixup2(EncDev E, uint32_t slice) {
    const uint32_t cnt = min(E.fix2_count[slice], E.fix_cap);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t i = (blockIdx.x * 256 + threadIdx.x) >> 5; i < cnt; i += gridDim.x * 8) {
        const uint64_t g = E.fix2_pos[(uint64_t)slice * E.fix_cap + i];
        const ChunkDesc cd = E.chunks[fix_chunk(E, g)];
        const uint8_t *__restrict__ p = E.in + cd.off;
        const uint32_t n = cd.len, pos = (uint32_t)(g - cd.off);
        const uint32_t t = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16);
        const uint32_t lowest = pos > E.window ? pos - E.window : 0u;             // candidates: lowest <= q < hi
        const uint32_t hi = E.fix2_j[(uint64_t)slice * E.fix_cap + i];            // the chain showed that (hi, pos) holds no occurrence; hi itself is another trigram
        // Scan downwards, 512 aligned bytes per step: lane L compares the trigram at each of the 16 positions of its 16 bytes
        // (byte-wise SIMD compares of the words shifted by 0, 1 and 2 bytes), the highest hit inside [lowest, hi) wins.
        const uint32_t T0 = (t & 0xFFu) * 0x01010101u, T1 = ((t >> 8) & 0xFFu) * 0x01010101u, T2 = (t >> 16) * 0x01010101u;
        const int64_t glow = (int64_t)cd.off + lowest, ghi = (int64_t)cd.off + hi;
        int32_t best = -1;
        for (int64_t S = (ghi - 1) & ~511ll; hi > lowest && S + 512 > glow; S -= 512) {
            const int64_t g0 = S + 16 * (int64_t)lane;
            uint32_t wv[5];
#pragma unroll
            for (uint32_t u = 0; u < 5; u++) wv[u] = g0 + 4 * u >= 0 ? ld_in32(E.in, (uint64_t)(g0 + 4 * u), E.in_size) : 0u;
            uint32_t m16 = 0;
#pragma unroll
            for (uint32_t u = 0; u < 4; u++) {
                const uint32_t x1 = __funnelshift_r(wv[u], wv[u + 1], 8), x2 = __funnelshift_r(wv[u], wv[u + 1], 16);
                const uint32_t m = __vcmpeq4(wv[u], T0) & __vcmpeq4(x1, T1) & __vcmpeq4(x2, T2);      // 0xFF per matching position
                m16 |= (((m & 0x80808080u) * 0x00204081u) >> 28) << (4 * u);
            }
            const int64_t lo_cut = glow - g0, hi_cut = ghi - g0;                  // valid bits: lo_cut <= bit < hi_cut
            if (lo_cut > 0) m16 &= lo_cut >= 16 ? 0u : ~((1u << (uint32_t)lo_cut) - 1u);
            if (hi_cut < 16) m16 &= hi_cut <= 0 ? 0u : (1u << (uint32_t)hi_cut) - 1u;
            const int32_t mine = m16 ? (int32_t)(g0 - (int64_t)cd.off) + 31 - __clz((int)m16) : -1;
            best = __reduce_max_sync(0xFFFFFFFFu, mine);
            if (best >= 0) break;
        }
        uint32_t out = t & 0xFFu;                                                  // literal unless a match is found
        if (best >= 0) {                                                           // uniform
            const uint32_t q = (uint32_t)best;
            const uint32_t limit = min(E.max_len - 3, n - (pos + 3));
            uint32_t k = limit;
            for (uint32_t r = 0; r < limit; r += 32) {
                const uint32_t kk = r + lane;
                const bool ne = kk < limit ? p[pos + 3 + kk] != p[q + 3 + kk] : true;
                const uint32_t mm = __ballot_sync(0xFFFFFFFFu, ne);
                if (mm) { k = min(limit, r + (uint32_t)__ffs((int)mm) - 1u); break; }
            }
            out = ((3 + k) << 16) | (pos - q);
        }
        if (lane == 0) E.md[g] = out;
    }
}

// =============================================================================== K3 parse_exits
// Greedy walk i -> i + step(i), step = match length or 1.  For a tile [ts,te) and each of the <= 258
// positions a previous tile can jump into, exit = (first position >= te reached) - te  (SURVEY App. C).
// A warp handles 32 consecutive tiles, lane = tile: the right-to-left DP of a tile is serial, the 32 tiles are the parallelism.
// md[] is read in chunks of 32 positions x 32 tiles: one fully coalesced 128-byte load per tile row (the next chunk is already in
// flight while the current one is processed), transposed through shared memory (row stride 33 words: conflict free both ways).
// The DP's ring of the last 260 exits lives in shared memory as [index][lane]; when the DP is done it holds the exits of positions
// 0..259, i.e. the tile's exit table, which is written out tile by tile with contiguous stores.
constexpr uint32_t kRing = 260;
#ifndef B2F_PX_WARPS
#define B2F_PX_WARPS 1
#endif
constexpr uint32_t kPxWarps = B2F_PX_WARPS;
constexpr uint32_t kPxWarpSmem = 32 * 33 * 4 + kRing * 32 * 2 + 32 * 8 + 32 * 4;      // md chunk | ring | row pointers | row lengths
constexpr uint32_t kPxSmem = kPxWarps * kPxWarpSmem;
__global__ void __launch_bounds__(kPxWarps * 32) k_parse_exits(EncDev E, uint32_t off, uint32_t lim) {
    extern __shared__ __align__(16) uint8_t pxs[];
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t tile_base = off + (blockIdx.x * kPxWarps + w) * 32u;
    if (tile_base >= lim) return;
    uint8_t *ws = pxs + w * kPxWarpSmem;
    uint32_t *mds = reinterpret_cast<uint32_t *>(ws);                                  // [32][33]
    uint16_t *ring = reinterpret_cast<uint16_t *>(ws + 32 * 33 * 4);                   // [kRing][32]
    const uint32_t **rowp = reinterpret_cast<const uint32_t **>(ws + 32 * 33 * 4 + kRing * 32 * 2);
    uint32_t *rowlen = reinterpret_cast<uint32_t *>(ws + 32 * 33 * 4 + kRing * 32 * 2 + 32 * 8);
    const uint32_t tile = tile_base + lane;
    uint32_t len = 0;                                                                  // positions of this lane's tile (0: no tile)
    const uint32_t *mrow = E.md;
    if (tile < lim) {
        const uint32_t c = find_owner(E.tile0, E.n_chunks, tile);
        const ChunkDesc cd = E.chunks[c];
        const uint32_t ts = (tile - E.tile0[c]) * kTile;
        len = min(kTile, cd.len - ts);
        mrow = E.md + cd.off + ts;
    }
    rowp[lane] = mrow; rowlen[lane] = len;
    __syncwarp();
    const uint32_t maxlen = __reduce_max_sync(0xFFFFFFFFu, len);
    uint32_t ridx = len ? (len - 1) % kRing : 0u;                                      // ring index of the position the DP handles next
    uint32_t prev_ex = 0;                                                              // exit of position p + 1 (literal steps need no ring read)
    uint32_t cur[32], nxt[32];
    int32_t j = (int32_t)((maxlen + 31) / 32) - 1;                                     // chunk = positions [32 j, 32 j + 32) of every tile
#pragma unroll
    for (uint32_t r = 0; r < 32; r++) cur[r] = (uint32_t)(32 * j) + lane < rowlen[r] ? __ldg(rowp[r] + 32 * j + lane) : 0u;
    for (; j >= 0; j--) {
        if (j > 0) {
#pragma unroll
            for (uint32_t r = 0; r < 32; r++) nxt[r] = (uint32_t)(32 * (j - 1)) + lane < rowlen[r] ? __ldg(rowp[r] + 32 * (j - 1) + lane) : 0u;
        }
#pragma unroll
        for (uint32_t r = 0; r < 32; r++) mds[r * 33 + lane] = cur[r];
        __syncwarp();
        // the lane's 32 values of the chunk go to registers first (conflict-free reads), so that the serial chain below is only
        // ring read -> select -> ring write
        uint32_t mv[32];
#pragma unroll
        for (uint32_t k = 0; k < 32; k++) mv[k] = mds[lane * 33 + k];
        if (__all_sync(0xFFFFFFFFu, 32u * (uint32_t)(j + 1) <= len)) {
            // full chunk in every lane (all but the last chunks of a chunk's last tile): no branches, so that everything that does not
            // depend on the previous position (step, ring indices) is computed ahead and the serial chain is read -> select -> write
            const uint32_t pbase = 32u * (uint32_t)j;
#pragma unroll
            for (int k = 31; k >= 0; k--) {
                const uint32_t step = max(mv[k] >> 16, 1u);
                uint32_t rk = ridx + kRing - (31u - (uint32_t)k); rk -= rk >= kRing ? kRing : 0u;      // ring index of position pbase + k
                uint32_t ix = rk + step; ix -= ix >= kRing ? kRing : 0u;
                const uint32_t nx = pbase + (uint32_t)k + step;
                const uint32_t rv = ring[ix * 32 + lane];
                const uint32_t ex = nx >= len ? nx - len : step == 1 ? prev_ex : rv;
                ring[rk * 32 + lane] = (uint16_t)ex;
                prev_ex = ex;
            }
            ridx = ridx + kRing - 32u; ridx -= ridx >= kRing ? kRing : 0u;
        } else if ((uint32_t)(32 * j) < len) {
            const uint32_t kmax = min(32u, len - 32u * (uint32_t)j);
#pragma unroll
            for (int k = 31; k >= 0; k--) {
                if ((uint32_t)k < kmax) {
                    const uint32_t step = max(mv[k] >> 16, 1u);
                    const uint32_t nx = 32u * (uint32_t)j + (uint32_t)k + step;
                    uint32_t ix = ridx + step; ix -= ix >= kRing ? kRing : 0u;
                    const uint32_t rv = ring[ix * 32 + lane];                       // (always a valid index; unused when the step leaves the tile)
                    const uint32_t ex = nx >= len ? nx - len : step == 1 ? prev_ex : rv;
                    ring[ridx * 32 + lane] = (uint16_t)ex;
                    prev_ex = ex;
                    ridx = ridx ? ridx - 1 : kRing - 1;
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (uint32_t r = 0; r < 32; r++) cur[r] = nxt[r];
    }
    // ring[i][lane] = exit of position i for i < min(len, 260): the exit tables, one tile after the other
    for (uint32_t r = 0; r < 32; r++) {
        const uint32_t lr = rowlen[r];
        if (!lr) break;
        uint16_t *__restrict__ xt = E.exit_tab + (uint64_t)(tile_base + r) * kExitW;
        for (uint32_t i = lane; i < min(kExitW, lr); i += 32) xt[i] = ring[i * 32 + r];
    }
}

// =============================================================================== K4 parse_stitch
// entry[t + 1] = exit_tab[t][entry[t]]: a serial chain over the tiles of a chunk.  One warp per chunk: the exit tables of 32
// tiles at a time are fetched with coalesced loads into shared memory, then the 32 chain steps are shared-memory reads
// (they were dependent L2 loads, ~0.7 us each, 256 per 256 KiB chunk).
constexpr uint32_t kStWarps = 4;
constexpr uint32_t kStRowWords = kExitW / 2;                                  // 129 words per exit table
constexpr uint32_t kStSmem = kStWarps * 32 * kStRowWords * 4;
static_assert(kExitW % 2 == 0, "exit tables are copied as 32-bit words");
__global__ void __launch_bounds__(kStWarps * 32) k_parse_stitch(EncDev E, uint32_t off, uint32_t lim) {
    extern __shared__ __align__(16) uint32_t sts[];
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t c = off + blockIdx.x * kStWarps + w;
    if (c >= lim) return;
    uint32_t *rows = sts + w * 32 * kStRowWords;
    const uint16_t *rows16 = reinterpret_cast<const uint16_t *>(rows);
    const uint32_t t0 = E.tile0[c], t1 = E.tile0[c + 1];
    uint32_t p = 0;
    for (uint32_t tb = t0; tb < t1; tb += 32) {
        const uint32_t nr = min(32u, t1 - tb);
        // the tables of tiles [tb, tb + nr) are contiguous: nr * 129 words (516-byte rows keep 4-byte alignment)
        const uint32_t *__restrict__ src = reinterpret_cast<const uint32_t *>(E.exit_tab + (uint64_t)tb * kExitW);
#pragma unroll 16
        for (uint32_t i = lane; i < nr * kStRowWords; i += 32) rows[i] = __ldg(src + i);
        __syncwarp();
        uint32_t mine = 0;
        for (uint32_t r = 0; r < nr; r++) {                                   // uniform: every lane follows the chain, lane r keeps entry r
            if (r == lane) mine = p;
            p = rows16[r * kExitW + p];
        }
        if (lane < nr) E.tile_entry[tb + lane] = (uint16_t)mine;
        __syncwarp();
    }
}

// =============================================================================== K5 parse_emit
// CTA = 64 consecutive tiles of one chunk (one DEFLATE block => one histogram), warp = 32 tiles, lane = tile.  md[] is streamed
// like in k_parse_exits (coalesced 128-byte rows, transposed through shared memory, next chunk in flight); each lane walks its
// tile from its entry point, emits the symbols and counts them in the CTA's shared-memory histogram.
constexpr uint32_t kPeWarpSmem = 32 * 33 * 4;
__global__ void __launch_bounds__(64) k_parse_emit(EncDev E, uint32_t off) {
    __shared__ uint32_t sh[kHistStride], shl[256];                            // lit/len + distance codes | raw match lengths (len - 3)
    __shared__ __align__(16) uint32_t pes[2 * kPeWarpSmem / 4];
    const uint32_t grp = blockIdx.x + off;
    const uint32_t c = find_owner(E.grp0, E.n_chunks, grp);
    const ChunkDesc cd = E.chunks[c];
    for (uint32_t i = threadIdx.x; i < kHistStride; i += 64) sh[i] = 0;
    for (uint32_t i = threadIdx.x; i < 256; i += 64) shl[i] = 0;
    __syncthreads();
    const uint32_t w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t *mds = pes + w * (kPeWarpSmem / 4);                               // [32][33]
    const uint32_t n = cd.len;
    const uint32_t ntile_c = E.tile0[c + 1] - E.tile0[c];
    const uint32_t tkb = (grp - E.grp0[c]) * kGrpTiles + w * 32u;             // first tile (index inside the chunk) of this warp
    const uint32_t tk = tkb + lane;
    const bool valid = tk < ntile_c;
    const uint32_t tile = E.tile0[c] + tk;
    const uint32_t ts = tk * kTile;
    const uint32_t len = valid ? min(kTile, n - ts) : 0u;
    const uint32_t *__restrict__ mbase = E.md + cd.off + (uint64_t)tkb * kTile;          // row r starts at mbase + r * kTile
    uint32_t *__restrict__ so = E.sym + cd.off + ts;
    uint32_t i = valid ? E.tile_entry[tile] : 0u, cnt = 0;
    const uint32_t wlen = tkb < ntile_c ? min(32u * kTile, n - tkb * kTile) : 0u;        // positions of the warp's 32 tiles
    auto rowlen = [&](uint32_t r) { return wlen > r * kTile ? min(kTile, wlen - r * kTile) : 0u; };
    const uint32_t nchunk = (min(kTile, wlen) + 31) / 32;
    uint32_t cur[32], nxt[32];
#pragma unroll
    for (uint32_t r = 0; r < 32; r++) cur[r] = lane < rowlen(r) ? __ldg(mbase + r * kTile + lane) : 0u;
    for (uint32_t j = 0; j < nchunk; j++) {
        if (j + 1 < nchunk) {
#pragma unroll
            for (uint32_t r = 0; r < 32; r++) nxt[r] = 32 * (j + 1) + lane < rowlen(r) ? __ldg(mbase + r * kTile + 32 * (j + 1) + lane) : 0u;
        }
#pragma unroll
        for (uint32_t r = 0; r < 32; r++) mds[r * 33 + lane] = cur[r];
        __syncwarp();
        const uint32_t cend = min(len, 32 * (j + 1));
        while (i < cend) {
            const uint32_t v = mds[lane * 33 + (i - 32 * j)];
            const uint32_t L = v >> 16;
            if (L) {                                                          // lengths are counted raw (folded into codes at the end)
                uint32_t dc, de, dx;
                dist_code(v & 0xFFFFu, dc, de, dx);
                atomicAdd(&shl[L - 3], 1u);
                atomicAdd(&sh[286 + dc], 1u);
                so[cnt++] = kSymPtr | v;
                i += L;
            } else {
                atomicAdd(&sh[v & 0xFFu], 1u);
                so[cnt++] = v & 0xFFu;
                i += 1;
            }
        }
        __syncwarp();
#pragma unroll
        for (uint32_t r = 0; r < 32; r++) cur[r] = nxt[r];
    }
    if (valid) E.tile_nsym[tile] = cnt;
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < 256; k += 64) {                        // raw lengths -> length codes 257..285
        const uint32_t nl = shl[k];
        if (nl) { uint32_t lc, le, lx; length_code(k + 3, lc, le, lx); atomicAdd(&sh[lc], nl); }
    }
    __syncthreads();
    uint32_t *__restrict__ gh = E.hist + (uint64_t)cd.block * kHistStride;
    for (uint32_t k = threadIdx.x; k < 316; k += 64) if (sh[k]) atomicAdd(&gh[k], sh[k]);
}

// =============================================================================== K6 huff_build
// skip_built: leave out the blocks whose codes a slice already built (hdr_bits is preset to 0xFFFFFFFF by the host)
__global__ void __launch_bounds__(32) k_huff_build(EncDev E, uint32_t b_off, uint32_t skip_built) {
    __shared__ HuffWarp W;
    const uint32_t b = blockIdx.x + b_off;
    if (skip_built && E.hdr_bits[b] != 0xFFFFFFFFu) return;
    uint32_t *lit = E.litcode + (uint64_t)b * kLitStride;
    uint32_t *dist = E.distcode + (uint64_t)b * kDistStride;
    if (E.blocks[b].fixed) {
        if (threadIdx.x == 0) { build_fixed_codes(lit, dist); E.hdr_bits[b] = 0; }
    } else {
        // one warp per block (huff_warp.cuh); huff_build.cuh is the serial statement of the same construction
        const uint32_t nbits = hw_build_block_codes(E.hist + (uint64_t)b * kHistStride, lit, dist, E.hdr_words + (uint64_t)b * kHdrWords, W);
        if (threadIdx.x == 0) E.hdr_bits[b] = nbits;
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
__global__ void __launch_bounds__(256) k_tile_bits(EncDev E, uint32_t t_off, uint32_t t_lim) {
    const uint32_t tile = t_off + blockIdx.x * 8 + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (tile >= t_lim) return;
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
__global__ void __launch_bounds__(256) k_scan_tiles(EncDev E, uint32_t b_off) {
    __shared__ uint64_t wsum[8];
    __shared__ uint64_t carry_s;
    const uint32_t b = blockIdx.x + b_off, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
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
// Blocks [b0, b1) only: a stream's running bit position is carried in stream_end_bits from one range (slice) to the next, so the
// ranges must be scanned in order (the slices' streams are chained by events).
__global__ void __launch_bounds__(64) k_scan_blocks(EncDev E, uint32_t b0, uint32_t b1) {
    const uint32_t s = blockIdx.x * 64 + threadIdx.x;
    if (s >= E.n_streams) return;
    const uint32_t lo = max(b0, E.stream_blk0[s]), hi = min(b1, E.stream_blk0[s + 1]);
    if (lo >= hi) return;
    uint64_t pos = lo == E.stream_blk0[s] ? (uint64_t)E.hdr_len[s] * 8 : E.stream_end_bits[s];
    for (uint32_t b = lo; b < hi; b++) {
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
__global__ void __launch_bounds__(128) k_write_headers(EncDev E, uint32_t b_off, uint32_t b_lim) {
    const uint32_t b = b_off + blockIdx.x * 4 + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (b >= b_lim) return;
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
__global__ void __launch_bounds__(256) k_bitpack(EncDev E, uint32_t t_off, uint32_t t_lim) {
    extern __shared__ uint32_t pk[];
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tile = t_off + blockIdx.x * 8 + wid;
    if (tile >= t_lim) return;
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

static bool g_lz_fused = true;
cudaError_t enc_init_attributes() {
    cudaError_t e;
    e = cudaFuncSetAttribute(k_lz_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << kHashBits) * 4));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_lz_match, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatchSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_lz_find, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFindSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_parse_exits, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPxSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_parse_stitch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStSmem);
    if (e != cudaSuccess) return e;
    if (const char *v = getenv("B2F_LZ_FUSED")) g_lz_fused = atoi(v) != 0;          // 0 = k_lz_chain + k_lz_match (A/B and fallback)
    e = cudaFuncSetAttribute(k_bitpack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(8 * kPackWords * 4));
    return e;
}

// One slice = a contiguous range of chunks [c0, c1): chain -> match -> exits -> stitch -> emit, in order, on one stream.
static cudaError_t enc_launch_lz_slice(const EncDev &E, const uint32_t *h_seg0, const uint32_t *h_pt0, const uint32_t *h_tile0, const uint32_t *h_grp0,
                                       uint32_t c0, uint32_t c1, cudaStream_t st, StageTimer *tm, uint32_t slice) {
    if (c1 <= c0) return cudaSuccess;
    const uint32_t nseg = h_seg0[c1] - h_seg0[c0], npt = h_pt0[c1] - h_pt0[c0], nt = h_tile0[c1] - h_tile0[c0], ng = h_grp0[c1] - h_grp0[c0];
    if (g_lz_fused) {
        if (tm) tm->mark(st, "lz_find");
        k_lz_find<<<nseg, (kFindWarps + 1) * 32, kFindSmem, st>>>(E, h_seg0[c0], slice); B2F_LAUNCH_CHECK();
        if (tm) tm->mark(st, "lz_fixup");
        // grids sized for the usual share of deferred positions (0.3 % / 0.03 %); both kernels stride over whatever was queued
        const uint32_t gfix = min(4736u, max(148u, nseg * 4u));
        k_lz_fixup<<<gfix, 256, 0, st>>>(E, slice); B2F_LAUNCH_CHECK();
        k_lz_fixup2<<<gfix, 256, 0, st>>>(E, slice); B2F_LAUNCH_CHECK();
    } else {
        if (tm) tm->mark(st, "lz_chain");
        k_lz_chain<<<nseg, 32, (1u << kHashBits) * 4, st>>>(E, h_seg0[c0]); B2F_LAUNCH_CHECK();
        if (tm) tm->mark(st, "lz_match");
        k_lz_match<<<npt, 512, kMatchSmem, st>>>(E, h_pt0[c0]); B2F_LAUNCH_CHECK();
    }
    if (tm) tm->mark(st, "parse_exits");
    k_parse_exits<<<(nt + 32 * kPxWarps - 1) / (32 * kPxWarps), kPxWarps * 32, kPxSmem, st>>>(E, h_tile0[c0], h_tile0[c1]); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "parse_stitch");
    k_parse_stitch<<<(c1 - c0 + kStWarps - 1) / kStWarps, kStWarps * 32, kStSmem, st>>>(E, c0, c1); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "parse_emit");
    k_parse_emit<<<ng, 64, 0, st>>>(E, h_grp0[c0]); B2F_LAUNCH_CHECK();
    return cudaSuccess;
}

// LZ77 stage.  With aux streams the chunks are split into slices that run concurrently: the chain / exits kernels are
// latency bound (one warp per segment / thread per tile) and leave most issue slots free for the match kernel of another slice.
cudaError_t enc_launch_lz(const EncDev &E, const uint32_t *h_seg0, const uint32_t *h_pt0, const uint32_t *h_tile0, const uint32_t *h_grp0,
                          cudaStream_t st, StageTimer *tm, cudaStream_t *aux, cudaEvent_t *ev, uint32_t n_aux, const SliceFeed *feed,
                          const ChunkDesc *h_chunks, bool *sliced, SlicePipe *pipe) {
    if (sliced) *sliced = false;
    if (E.n_chunks == 0) return cudaSuccess;
    if (n_aux < 2 || E.n_chunks < 2 * n_aux) {
        if (feed) { if (tm) tm->mark(st, "h2d"); cudaError_t fe = feed->copy(feed->self, 0, E.n_chunks, st); if (fe != cudaSuccess) return fe; }
        return enc_launch_lz_slice(E, h_seg0, h_pt0, h_tile0, h_grp0, 0, E.n_chunks, st, tm, 0);
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
        if (h_chunks) while (c1 > c0 && c1 < E.n_chunks && h_chunks[c1].block == h_chunks[c1 - 1].block) c1++;   // whole DEFLATE blocks per slice
        e = cudaStreamWaitEvent(aux[g % kAuxStreams], ev[0], 0); if (e != cudaSuccess) return e;
        if (feed) { e = feed->copy(feed->self, c0, c1, aux[g % kAuxStreams]); if (e != cudaSuccess) return e; }   // this slice's H2D overlaps the previous slices' kernels
        e = enc_launch_lz_slice(E, h_seg0, h_pt0, h_tile0, h_grp0, c0, c1, aux[g % kAuxStreams], nullptr, g); if (e != cudaSuccess) return e;
        if (h_chunks && c1 > c0) {
            // the histograms of this slice's blocks are complete: their code construction (one thread per block, pure latency) runs
            // here, under the other slices' LZ77 kernels, instead of on the critical path after the join
            // (with the pipelined entropy stage the slices own ALL blocks: empty ones before the first / after the last chunk too)
            const uint32_t b0 = (pipe && pipe->n_slices == 0) ? 0u : h_chunks[c0].block;
            const uint32_t b1 = c1 < E.n_chunks ? h_chunks[c1].block : (pipe ? E.n_blocks : h_chunks[c1 - 1].block + 1);
            if (b1 > b0) { k_huff_build<<<b1 - b0, 32, 0, aux[g % kAuxStreams]>>>(E, b0, 0u); B2F_LAUNCH_CHECK(); }
            const uint32_t t0 = h_tile0[c0], t1 = h_tile0[c1];
            if (t1 > t0) { k_tile_bits<<<(t1 - t0 + 7) / 8, 256, 0, aux[g % kAuxStreams]>>>(E, t0, t1); B2F_LAUNCH_CHECK(); }
            if (sliced) *sliced = true;
            if (pipe && b1 > b0) {
                // The whole entropy stage of the slice's blocks runs here as well, under the other slices' LZ77 kernels: only the bit
                // position of the slice's first block depends on the previous slice (event chain), and the packed bytes of a slice
                // can leave for the host while the next slices are still being matched (pipe->h_pos = that position, for the copy).
                const uint32_t k = pipe->n_slices;                       // (a slice without chunks does not take part)
                k_scan_tiles<<<b1 - b0, 256, 0, aux[g % kAuxStreams]>>>(E, b0); B2F_LAUNCH_CHECK();
                if (k > 0) { e = cudaStreamWaitEvent(aux[g % kAuxStreams], pipe->ev_scan[k - 1], 0); if (e != cudaSuccess) return e; }
                k_scan_blocks<<<(E.n_streams + 63) / 64, 64, 0, aux[g % kAuxStreams]>>>(E, b0, b1); B2F_LAUNCH_CHECK();
                if (pipe->h_pos) { e = cudaMemcpyAsync(pipe->h_pos + k, E.stream_end_bits, 8, cudaMemcpyDeviceToHost, aux[g % kAuxStreams]); if (e != cudaSuccess) return e; }
                e = cudaEventRecord(pipe->ev_scan[k], aux[g % kAuxStreams]); if (e != cudaSuccess) return e;
                k_write_headers<<<(b1 - b0 + 3) / 4, 128, 0, aux[g % kAuxStreams]>>>(E, b0, b1); B2F_LAUNCH_CHECK();
                if (t1 > t0) { k_bitpack<<<(t1 - t0 + 7) / 8, 256, 8 * kPackWords * 4, aux[g % kAuxStreams]>>>(E, t0, t1); B2F_LAUNCH_CHECK(); }
                e = cudaEventRecord(pipe->ev_pack[k], aux[g % kAuxStreams]); if (e != cudaSuccess) return e;
                pipe->n_slices = k + 1;
            }
        }
        e = cudaEventRecord(ev[1 + g], aux[g % kAuxStreams]); if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(st, ev[1 + g], 0); if (e != cudaSuccess) return e;
        c0 = c1;
    }
    return cudaSuccess;
}
cudaError_t enc_launch_entropy(const EncDev &E, cudaStream_t st, StageTimer *tm, bool sliced) {
    if (tm) tm->mark(st, "huff_build");
    k_huff_build<<<E.n_blocks, 32, 0, st>>>(E, 0u, 1u); B2F_LAUNCH_CHECK();          // whatever the slices did not build (all blocks when not sliced)
    if (E.n_tiles && !sliced) {
        if (tm) tm->mark(st, "tile_bits");
        k_tile_bits<<<(E.n_tiles + 7) / 8, 256, 0, st>>>(E, 0u, E.n_tiles); B2F_LAUNCH_CHECK();
    }
    if (tm) tm->mark(st, "scan");
    k_scan_tiles<<<E.n_blocks, 256, 0, st>>>(E, 0u); B2F_LAUNCH_CHECK();
    k_scan_blocks<<<(E.n_streams + 63) / 64, 64, 0, st>>>(E, 0u, E.n_blocks); B2F_LAUNCH_CHECK();
    if (tm) tm->mark(st, "write_headers");
    k_write_headers<<<(E.n_blocks + 3) / 4, 128, 0, st>>>(E, 0u, E.n_blocks); B2F_LAUNCH_CHECK();
    if (E.n_tiles) {
        if (tm) tm->mark(st, "bitpack");
        k_bitpack<<<(E.n_tiles + 7) / 8, 256, 8 * kPackWords * 4, st>>>(E, 0u, E.n_tiles); B2F_LAUNCH_CHECK();
    }
    return cudaSuccess;
}
cudaError_t enc_launch_compact(const EncDev &E, uint64_t *tile_symoff, uint64_t *total, uint32_t *dst, cudaStream_t st) {
    k_scan_nsym_single<<<1, 1024, 0, st>>>(E, tile_symoff, total); B2F_LAUNCH_CHECK();
    k_compact_syms<<<(E.n_tiles + 7) / 8, 256, 0, st>>>(E, tile_symoff, dst); B2F_LAUNCH_CHECK();
    return cudaSuccess;
}
uint32_t enc_launch_count_lz(bool sliced) { return (g_lz_fused ? 6u : 5u) + (sliced ? 2u : 0u); }
uint32_t enc_launch_count_entropy(bool has_tiles, bool sliced) { return has_tiles ? (sliced ? 5u : 6u) : 4u; }

}  // namespace b2f

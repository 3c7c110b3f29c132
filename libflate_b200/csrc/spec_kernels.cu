// spec_kernels.cu -- sub-block parallel inflate (sm_100a).
//
// A DEFLATE block is one serial bit stream, but Huffman parses self-synchronise: a decoder started at an arbitrary bit
// falls into step with the true parse after a few symbols.  Each candidate block (from the boundary finder) is cut into
// kSpecBits-bit (2048) subsegments, one THREAD each:
//   k_spec_headers   warp per block : parse the dynamic header, build the decode tables (kept in HBM, 12 KiB per block)
//   k_spec_round     thread per subsegment, Jacobi iteration: round 0 decodes from the subsegment's first bit (the first
//                    subsegment from the true first symbol); round r restarts from the exit of the left neighbour of
//                    round r-1 if that differs from the start used so far.  Fixed point == the true parse.
//   k_spec_verify    CTA per block  : checks that every start equals its left neighbour's exit (=> by induction the true
//                    parse), finds EndOfBlock, exclusive scans of symbol/byte counts
//   k_spec_tokens    thread per subsegment: final decode from the true start, writes one token per symbol
//   k_seg_plan/resolve/cuts/subst : LZ77 resolution of the token stream by ~4 KiB segments, one warp each, with 16-bit
//                    markers for bytes copied from before the segment; one substitution pass per chain of dependent segments
// Nothing here is trusted blindly: the host accepts a block only if its verified EndOfBlock lands exactly on the next
// block of the chain; otherwise the stream goes to the exact in-order kernel (decode_kernels.cu).
// Reference behaviour being reproduced: src/deflate/decode.rs:112-130, symbol.rs:193-243, libflate_lz77/src/lib.rs:164-194.
#include <algorithm>
#include "common.cuh"
#include "inflate_core.cuh"
#include "spec_dev.cuh"

namespace b2f {

struct WarpSyncS { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

// ---------------------------------------------------------------------------------- headers -> tables in HBM
__global__ void __launch_bounds__(128) k_spec_headers(SpecDev S) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    InflateTables *tabs = reinterpret_cast<InflateTables *>(smem_raw);
    const uint32_t wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b = blockIdx.x * 4 + wid;
    if (b >= S.n_blocks) return;
    InflateTables &T = tabs[wid];
    const uint32_t m = S.blk_member[b];
    const uint8_t *p0 = S.in + S.in_off[m];
    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3);
    BitIn bi;
    bi_init(bi, p0 - lead, S.in_len[m] + lead, S.blk_bit[b] + 8ull * lead);
    uint32_t data_rel = 0xFFFFFFFFu, flags = 0;
    const uint32_t bfinal = bi_read(bi, 1);
    const uint32_t btype = bi_read(bi, 2);
    if (!bi_check(bi) && btype == 2) {
        if (load_dynamic(bi, T, (int)lane, 32, WarpSyncS()) == kInfOk) {
            const uint64_t rel = bi.pos - 8ull * lead - S.blk_bit[b];
            if (rel < 0xFFFFFF00ull && T.lit_maxbw) { data_rel = (uint32_t)rel; flags = bfinal; }
        }
    }
    __syncwarp();
    if (data_rel != 0xFFFFFFFFu) {                 // publish the tables (coalesced 16-byte copies)
        const uint4 *src = reinterpret_cast<const uint4 *>(&T);
        uint4 *dst = reinterpret_cast<uint4 *>(S.tabs + b);
        for (uint32_t i = lane; i < sizeof(InflateTables) / 16; i += 32) dst[i] = src[i];
    }
    if (lane == 0) { S.blk_data_rel[b] = data_rel; S.blk_flags[b] = flags; }
}

// ---------------------------------------------------------------------------------- per-thread bit reader over HBM
struct TBits {
    const uint32_t *wp;      // aligned word pointer of the member (covers `lead` bytes before it)
    uint64_t nwords;         // readable words (zero fill beyond)
    uint64_t widx;           // next word to load
    uint64_t bb; uint32_t bc;
    uint64_t pos;            // absolute bit position inside the (lead-shifted) member
};
__device__ __forceinline__ uint32_t tb_word(const TBits &t, uint64_t i) { return i < t.nwords ? __ldg(t.wp + i) : 0u; }
__device__ __forceinline__ void tb_seek(TBits &t, uint64_t bit) {
    t.pos = bit; t.widx = bit >> 5;
    const uint32_t drop = (uint32_t)bit & 31u;
    t.bb = (uint64_t)tb_word(t, t.widx) >> drop; t.bc = 32 - drop; t.widx++;
}
__device__ __forceinline__ void tb_refill(TBits &t) {
    if (t.bc < 32) { t.bb |= (uint64_t)tb_word(t, t.widx) << t.bc; t.bc += 32; t.widx++; }
}
__device__ __forceinline__ void tb_skip(TBits &t, uint32_t n) { t.bb >>= n; t.bc -= n; t.pos += n; }

constexpr uint32_t kExitEob = 0xFFFFFFFFu, kExitBad = 0xFFFFFFFEu, kExitDead = 0xFFFFFFFDu;
constexpr uint32_t kTokSkip = 0x7FFFFFFFu;      // padding token (produces no output)

// Decodes symbols from t.pos until t.pos >= stop_abs, EndOfBlock or an undecodable pattern.
// kEmit: writes one token per symbol to tok[].  Returns the exit code (kExit* or 0 = ran to stop).
// Token sink of the final pass: 8 tokens (one 32-byte sector) are staged in shared memory per thread and written with two
// 16-byte stores, so HBM sees full sectors instead of 4-byte read-modify-writes.
struct TokSink {
    uint32_t *stage;            // 8 words in shared memory (per thread)
    uint4 *gout;                // 32-byte aligned destination
    uint32_t n;
    __device__ __forceinline__ void put(uint32_t v) {
        stage[n & 7u] = v; n++;
        if ((n & 7u) == 0) { const uint4 *s4 = reinterpret_cast<const uint4 *>(stage); gout[0] = s4[0]; gout[1] = s4[1]; gout += 2; }
    }
    __device__ __forceinline__ void finish() { while (n & 7u) put(kTokSkip); }
};

// Decodes symbols from t.pos until t.pos >= stop_abs, EndOfBlock or an undecodable pattern.
// The loop is WARP-UNIFORM: every lane of `mask` iterates until all of them are done, one symbol per iteration, so the
// lanes re-converge after every symbol (independent per-lane loops drift apart and end up fully serialised).
// `run` = this lane has work.  kEmit: writes one token per symbol.  Returns the exit code (kExit* or 0 = ran to stop).
template <bool kEmit>
__device__ __forceinline__ uint32_t spec_decode(uint32_t mask, bool run, TBits &t, const InflateTables &T, uint64_t stop_abs,
                                                uint32_t &nsym, uint32_t &nbytes, TokSink *tok, int32_t &min_rel) {
    uint32_t result = 0;
    bool active = run && t.pos < stop_abs;
    while (__any_sync(mask, active)) {
        if (active) {
            tb_refill(t);
            uint32_t e = T.lit[(uint32_t)t.bb & ((1u << kLitBits) - 1u)];
            if ((e & 15u) == 0) e = lookup_code(T, true, (uint32_t)t.bb & 0x7FFFu);     // long code or unassigned
            const uint32_t w = e & 15u, kind = (e >> 4) & 3u;
            if (w == 0 || kind == kKindSpecial) { result = kExitBad; active = false; }
            else if (kind == kKindLit) {
                tb_skip(t, w);
                if (kEmit) tok->put(e >> 8);
                nsym++; nbytes++;
            } else if (kind == kKindEob) { tb_skip(t, w); result = kExitEob; active = false; }
            else {
                const uint32_t eb = (e >> 20) & 15u;
                const uint32_t len = ((e >> 8) & 0x1FFu) + (((uint32_t)(t.bb >> w)) & ((1u << eb) - 1u));
                tb_skip(t, w + eb);
                tb_refill(t);
                uint32_t d = T.dist[(uint32_t)t.bb & ((1u << kDistBits) - 1u)];
                if ((d & 15u) == 0) d = lookup_code(T, false, (uint32_t)t.bb & 0x7FFFu);
                const uint32_t wd = d & 15u;
                if (wd == 0) { result = kExitBad; active = false; }
                else {
                    const uint32_t deb = (d >> 24) & 15u;
                    const uint32_t dist = ((d >> 8) & 0xFFFFu) + (((uint32_t)(t.bb >> wd)) & ((1u << deb) - 1u));
                    tb_skip(t, wd + deb);
                    if (kEmit) {
                        tok->put(kSymPtr | (len << 16) | dist);
                        const int32_t src = (int32_t)nbytes - (int32_t)dist;          // relative to the subsegment's first output byte
                        if (src < min_rel) min_rel = src;
                    }
                    nsym++; nbytes += len;
                }
            }
            if (active && t.pos >= stop_abs) active = false;
        }
    }
    return result;
}

__device__ __forceinline__ uint32_t owner_u32(const uint32_t *__restrict__ prefix, uint32_t n, uint32_t idx) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= idx) lo = mid; else hi = mid; }
    return lo;
}

struct SegCtx { uint32_t b, k, sg; uint64_t lead8, blk_abs, blk_len; uint32_t data_rel; bool usable; };

__device__ __forceinline__ void load_tables_smem(InflateTables &Ts, const InflateTables *g) {
    const uint4 *src = reinterpret_cast<const uint4 *>(g);
    uint4 *dst = reinterpret_cast<uint4 *>(&Ts);
    for (uint32_t i = threadIdx.x; i < sizeof(InflateTables) / 16; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

// ---------------------------------------------------------------------------------- speculative rounds
// CTA = 128 consecutive subsegments of ONE block (tables staged in shared memory).
__global__ void __launch_bounds__(kSpecCta) k_spec_round(SpecDev S, uint32_t round) {
    __shared__ __align__(16) InflateTables Ts;
    const uint32_t b = owner_u32(S.blk_cta0, S.n_blocks, blockIdx.x);
    const uint32_t data_rel = S.blk_data_rel[b];
    if (data_rel == 0xFFFFFFFFu) return;                         // unusable header: the block never enters the chain
    const uint32_t nseg = S.blk_seg0[b + 1] - S.blk_seg0[b];
    const uint32_t k = (blockIdx.x - S.blk_cta0[b]) * kSpecCta + threadIdx.x;
    const uint32_t sg = S.blk_seg0[b] + min(k, nseg - 1);
    const uint32_t k0 = data_rel / kSpecBits;
    const uint32_t m = S.blk_member[b];
    const uint8_t *p0 = S.in + S.in_off[m];
    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3);
    const uint64_t blk_abs = S.blk_bit[b] + 8ull * lead, blk_len = S.blk_end[b] - S.blk_bit[b];
    const uint64_t seg_end = min((uint64_t)(k + 1) * kSpecBits, blk_len);
    // decide what this lane does: nothing (out of range), bookkeeping only, or a (re-)decode from `start`
    bool decode = false, store = false;
    uint32_t start = 0, ex = 0;
    if (k < nseg) {
        if (k < k0) { S.s_start[sg] = kExitDead; S.s_exit[sg] = kExitDead; }                    // header bits only
        else if (round == 0) { start = k == k0 ? data_rel : k * kSpecBits; store = true; }
        else if (k == k0) S.s_exit[sg] = S.s_exit_prev[sg];                                      // starts at the true first symbol: carry over
        else {
            const uint32_t pe = S.s_exit_prev[sg - 1];
            const uint32_t cur = S.s_start[sg];
            // An unusable neighbour exit (its speculative parse ran into EndOfBlock / an unassigned code, or the block really
            // ends there) carries no information about this subsegment: keep the current parse.
            if (pe >= kExitDead || pe == cur) S.s_exit[sg] = S.s_exit_prev[sg];
            else { start = pe; store = true; }
        }
        if (store) { if (start >= seg_end) ex = start; else decode = true; }     // neighbour's last symbol may already cover this subsegment
    }
    // confirmation rounds mostly carry results over: a CTA without a subsegment to decode skips the 11.6 KiB table load
    if (!__syncthreads_or(decode ? 1 : 0)) {
        if (store) { S.s_start[sg] = start; S.s_exit[sg] = ex; S.s_nsym[sg] = 0; S.s_nbytes[sg] = 0; if (round) atomicOr(S.changed + round, 1u); }
        return;
    }
    load_tables_smem(Ts, S.tabs + b);
    TBits t;
    t.wp = reinterpret_cast<const uint32_t *>(p0 - lead);
    t.nwords = (S.in_len[m] + lead + 3) >> 2;
    t.pos = 0; t.bb = 0; t.bc = 0; t.widx = 0;
    if (decode) tb_seek(t, blk_abs + start);
    uint32_t nsym = 0, nbytes = 0;
    int32_t dummy = 0;
    const uint32_t mask = __activemask();
    const uint32_t r = spec_decode<false>(mask, decode, t, Ts, blk_abs + seg_end, nsym, nbytes, nullptr, dummy);
    if (decode) {
        if (r == kExitEob) { ex = kExitEob; S.s_eob_end[sg] = (uint32_t)(t.pos - blk_abs); }
        else if (r == kExitBad) ex = kExitBad;
        else ex = (uint32_t)(t.pos - blk_abs);
    }
    if (store) {
        S.s_start[sg] = start; S.s_exit[sg] = ex; S.s_nsym[sg] = nsym; S.s_nbytes[sg] = nbytes;
        if (round) atomicOr(S.changed + round, 1u);
    }
}

// ---------------------------------------------------------------------------------- verify + scan (CTA per block)
__global__ void __launch_bounds__(256) k_spec_verify(SpecDev S) {
    __shared__ uint64_t wsum_b[8], wsum_s[8];
    __shared__ uint64_t carry_b, carry_s;
    __shared__ uint32_t bad, eob_seg;
    const uint32_t b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t data_rel = S.blk_data_rel[b];
    if (tid == 0) { carry_b = 0; carry_s = 0; bad = data_rel == 0xFFFFFFFFu ? 1u : 0u; eob_seg = 0xFFFFFFFFu; }
    __syncthreads();
    const uint32_t s0 = S.blk_seg0[b], nseg = S.blk_seg0[b + 1] - s0;
    const uint32_t k0 = data_rel == 0xFFFFFFFFu ? 0 : data_rel / kSpecBits;
    const uint32_t *ex_final = S.s_exit;                         // results of the last round
    if (!bad) {
        // pass 1: first EndOfBlock on the chain + consistency of every link before it
        for (uint32_t base = k0; base < nseg; base += 256) {
            const uint32_t k = base + tid;
            if (k < nseg) {
                const uint32_t ex = ex_final[s0 + k];
                if (ex == kExitEob) atomicMin(&eob_seg, k);
            }
        }
        __syncthreads();
        uint32_t e = eob_seg;
        bool truncated = false;
        if (e == 0xFFFFFFFFu) {
            // no EndOfBlock before the end of the block's extent: if the chain is consistent all the way, the extent was cut
            // short by a false-positive candidate -> status 2 tells the host to drop the candidate that follows this block
            truncated = true; e = nseg - 1;
        }
        __syncthreads();
        {
            for (uint32_t base = k0; base <= e; base += 256) {
                const uint32_t k = base + tid;
                if (k <= e) {
                    const uint32_t st = S.s_start[s0 + k], ex = ex_final[s0 + k];
                    const uint32_t want = k == k0 ? data_rel : ex_final[s0 + k - 1];
                    if (st != want || st >= kExitDead || ((k < e || truncated) && ex >= kExitDead)) atomicOr(&bad, 1u);
                }
            }
        }
        __syncthreads();
        if (truncated) { if (tid == 0) { S.blk_status[b] = bad ? 1u : 2u; S.blk_nout[b] = 0; S.blk_ntok[b] = 0; S.blk_eob_end[b] = 0; } return; }
    }
    if (bad) { if (tid == 0) { S.blk_status[b] = 1; S.blk_nout[b] = 0; S.blk_ntok[b] = 0; S.blk_eob_end[b] = 0; } return; }
    const uint32_t e = eob_seg;
    // pass 2: exclusive scans of bytes / symbols over subsegments k0..e
    for (uint32_t base = k0; base <= e; base += 256) {
        const uint32_t k = base + tid;
        const bool in = k <= e;
        const uint64_t xb = in ? S.s_nbytes[s0 + k] : 0, xs = in ? ((uint64_t)S.s_nsym[s0 + k] + 7) & ~7ull : 0;   // token regions are padded to 8
        uint64_t ib = xb, is = xs;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t tb = __shfl_up_sync(0xFFFFFFFFu, ib, d), ts = __shfl_up_sync(0xFFFFFFFFu, is, d);
            if ((int)lane >= d) { ib += tb; is += ts; }
        }
        if (lane == 31) { wsum_b[wid] = ib; wsum_s[wid] = is; }
        __syncthreads();
        uint64_t ob = 0, os = 0;
        for (uint32_t w = 0; w < wid; w++) { ob += wsum_b[w]; os += wsum_s[w]; }
        const uint64_t cb = carry_b, cs = carry_s;
        if (in) { S.s_out_rel[s0 + k] = cb + ob + ib - xb; S.s_tok_rel[s0 + k] = cs + os + is - xs; }
        __syncthreads();
        if (tid == 255) { carry_b = cb + ob + ib; carry_s = cs + os + is; }
        __syncthreads();
    }
    if (tid == 0) { S.blk_status[b] = 0; S.blk_nout[b] = carry_b; S.blk_ntok[b] = carry_s; S.blk_eob_end[b] = S.s_eob_end[s0 + e]; S.blk_eob_seg[b] = e; }
}

// ---------------------------------------------------------------------------------- final decode -> tokens
__global__ void __launch_bounds__(kSpecCta) k_spec_tokens(SpecDev S, uint32_t cta_off) {
    __shared__ __align__(16) InflateTables Ts;
    __shared__ __align__(16) uint32_t stage[kSpecCta * 8];
    const uint32_t cta = blockIdx.x + cta_off;
    const uint32_t b = owner_u32(S.blk_cta0, S.n_blocks, cta);
    if (S.blk_sel[b] == 0) return;                               // not on the verified chain
    load_tables_smem(Ts, S.tabs + b);
    const uint32_t data_rel = S.blk_data_rel[b];
    const uint32_t k = (cta - S.blk_cta0[b]) * kSpecCta + threadIdx.x;
    const uint32_t k0 = data_rel / kSpecBits, e = S.blk_eob_seg[b];
    const bool mine = k >= k0 && k <= e;
    const uint32_t sg = S.blk_seg0[b] + (mine ? k : k0);
    const uint32_t m = S.blk_member[b];
    const uint8_t *p0 = S.in + S.in_off[m];
    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3);
    const uint64_t blk_abs = S.blk_bit[b] + 8ull * lead, blk_len = S.blk_end[b] - S.blk_bit[b];
    const uint32_t start = S.s_start[sg];
    const uint64_t seg_end = min((uint64_t)(k + 1) * kSpecBits, blk_len);
    const bool decode = mine && start < seg_end;
    if (mine && !decode) S.s_min_src[sg] = INT64_MAX;                     // no symbol starts in this subsegment
    TBits t;
    t.wp = reinterpret_cast<const uint32_t *>(p0 - lead);
    t.nwords = (S.in_len[m] + lead + 3) >> 2;
    t.pos = 0; t.bb = 0; t.bc = 0; t.widx = 0;
    if (decode) tb_seek(t, blk_abs + start);
    uint32_t nsym = 0, nbytes = 0;
    int32_t min_rel = INT32_MAX;
    TokSink sink = { stage + threadIdx.x * 8, reinterpret_cast<uint4 *>(S.tokens + S.blk_tok0[b] + (decode ? S.s_tok_rel[sg] : 0)), 0 };
    const uint32_t mask = __activemask();
    spec_decode<true>(mask, decode, t, Ts, blk_abs + seg_end, nsym, nbytes, &sink, min_rel);
    if (decode) {
        sink.finish();
        S.s_min_src[sg] = min_rel == INT32_MAX ? INT64_MAX : (int64_t)S.s_out_rel[sg] + min_rel;
    }
}

// ---------------------------------------------------------------------------------- LZ77 resolution by segments + markers
// The reference resolves a stream's back-references strictly in order into one growing buffer (libflate_lz77/src/lib.rs:164-194).
// Here the token stream of the verified chain is cut into SEGMENTS of about kSegBytes output bytes and every segment is resolved
// by its own warp AS IF nothing were known about the bytes before it: a byte copied from before the segment's first byte becomes
// a 16-bit marker (kMarker | distance before the segment start - 1); copies of markers stay markers.  Then
//   k_seg_cuts   finds the segments no later segment reaches across (libflate's 256 KiB LZ77 chunks never reference each other,
//                libflate_lz77/src/default.rs:73,108; foreign streams have few or no such cuts), and
//   k_seg_subst  runs one CTA per chain of segments between two cuts: segment after segment, every marker is replaced by the final
//                byte it points at (all targets lie before the segment, i.e. are final) out of a shared-memory window.
// The serial dependency of the reference is thereby reduced to one barrier per segment inside a chain; everything else -- the
// whole token walk -- runs on thousands of independent warps.  Cross-block references (zlib / gzip output) are ordinary markers.
__global__ void __launch_bounds__(128) k_seg_plan(SpecDev S, uint32_t slot_lo, uint32_t slot_hi) {
    const uint32_t slot = slot_lo + blockIdx.x * 128 + threadIdx.x;
    if (slot >= slot_hi) return;
    const uint32_t k = owner_u32(S.sel_slot0, S.n_sel, slot);
    const uint32_t b = S.sel_blocks[k], j = slot - S.sel_slot0[k];
    const uint32_t s0 = S.blk_seg0[b], k0 = S.blk_data_rel[b] / kSpecBits, e = S.blk_eob_seg[b];
    const uint64_t nout = S.blk_nout[b], ntok = S.blk_ntok[b];
    // the block's subsegments k0..e have non-decreasing output offsets: segment j = those that start in [j, j+1) * kSegBytes
    auto first = [&](uint64_t target) {
        uint32_t lo = k0, hi = e + 1;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (S.s_out_rel[s0 + mid] >= target) hi = mid; else lo = mid + 1; }
        return lo;
    };
    const uint32_t f0 = j == 0 ? k0 : first((uint64_t)j * kSegBytes), f1 = first((uint64_t)(j + 1) * kSegBytes);
    uint64_t t0 = 0, t1 = 0, o0 = 0, o1 = 0;
    if (f0 < f1) {
        t0 = S.s_tok_rel[s0 + f0]; o0 = S.s_out_rel[s0 + f0];
        t1 = f1 <= e ? S.s_tok_rel[s0 + f1] : ntok; o1 = f1 <= e ? S.s_out_rel[s0 + f1] : nout;
    }
    S.seg_tok[slot] = S.blk_tok0[b] + t0; S.seg_ntok[slot] = (uint32_t)(t1 - t0);
    S.seg_out[slot] = S.blk_out0[b] + o0; S.seg_nout[slot] = (uint32_t)(o1 - o0);
    S.seg_member[slot] = S.blk_member[b]; S.seg_reach[slot] = 0; S.seg_cut[slot] = 0;
}

constexpr uint32_t kSegMask = kSegRing - 1;
constexpr uint32_t kSegFreeQ = 2 * kSegRing;                     // byte offset of the queue of order-free copies (<= 64 entries + 4 read-ahead)
constexpr uint32_t kSegOrdQ = kSegFreeQ + 68 * 8;                // byte offset of the queue of in-order copies (<= 32 entries + 2 read-ahead)
constexpr uint32_t kSegSmem = kSegOrdQ + 34 * 8;
static_assert((kSegRing & kSegMask) == 0 && kSegStepMax + kSegLazy + 258 + 8 < kSegRing, "segment ring geometry");

// ceil(65536 / p): k mod p == k - p * ((k * inv) >> 16) for p < 32, k < 400 (run-length matches; avoids a division per symbol)
__constant__ uint32_t kInvPeriod[32] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115};

__global__ void __launch_bounds__(32) k_seg_resolve(SpecDev S, uint32_t slot_lo) {
    extern __shared__ __align__(16) uint8_t seg_smem[];
    uint16_t *ring = reinterpret_cast<uint16_t *>(seg_smem);          // symbol of segment position p at (a0 + p) & kSegMask
    uint2 *fq = reinterpret_cast<uint2 *>(seg_smem + kSegFreeQ), *oq = reinterpret_cast<uint2 *>(seg_smem + kSegOrdQ);
    const uint32_t lane = threadIdx.x, slot = blockIdx.x + slot_lo;
    const uint32_t nout = S.seg_nout[slot];
    if (!nout) return;
    const uint32_t ntok = S.seg_ntok[slot];
    const uint32_t *__restrict__ tok = S.tokens + S.seg_tok[slot];
    const uint64_t A = S.seg_out[slot];                               // offset in out / sym16 of the segment's first byte
    const uint32_t mem = S.seg_member[slot];
    const uint64_t avail = A - S.mem_out_off[mem];                    // bytes of the member that precede the segment
    uint16_t *__restrict__ gsym = S.sym16 + A;
    const uint32_t a0 = (uint32_t)A & kSegMask;                       // ring index == absolute offset mod kSegRing: 8-symbol groups of
                                                                      // the ring are 16-byte aligned groups of sym16
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t sub = lane >> 3, l8 = lane & 7u;                   // order-free copies: four at a time, eight lanes each
    const uint32_t lead = (8u - ((uint32_t)A & 7u)) & 7u;             // symbols before the first 16-byte aligned group of sym16
    uint32_t pos = 0, flushed = 0, err = 0, reach = 0;                // segment-relative
    uint32_t tnext = lane < ntok ? __ldg(tok + lane) : kTokSkip;
    for (uint32_t i0 = 0; i0 < ntok;) {
        const uint32_t tk = tnext;
        tnext = i0 + 32 + lane < ntok ? __ldg(tok + i0 + 32 + lane) : kTokSkip;    // the next step's tokens (a step nearly always takes all 32)
        bool live = tk != kTokSkip;
        bool is_m = live && (tk & kSymPtr);
        uint32_t len = !live ? 0u : is_m ? (tk >> 16) & 0x1FFu : 1u;
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += v; }
        uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31), take = 32;
        if (total > kSegStepMax) {                                    // (long runs only) keep the step's output well inside the ring
            take = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, incl <= kSegStepMax));
            if (lane >= take) { live = false; is_m = false; len = 0; }
            total = __shfl_sync(0xFFFFFFFFu, incl, take - 1);
            tnext = i0 + take + lane < ntok ? __ldg(tok + i0 + take + lane) : kTokSkip;
        }
        const uint32_t off = incl - len;                              // (garbage for dropped lanes, which do nothing)
        const uint32_t dst = pos + off;
        const uint32_t d0 = (a0 + dst) & kSegMask;
        if (live && !is_m) ring[d0] = (uint16_t)(tk & 0xFFu);
        const uint32_t dist = tk & 0xFFFFu;
        if (is_m && (uint64_t)dist > (uint64_t)dst + avail) err |= 2u;             // "Too long backword reference": the in-order kernel reports it
        const int32_t srel = (int32_t)dst - (int32_t)dist;            // segment-relative position of the first source byte
        const uint32_t before = (is_m && srel < 0) ? (uint32_t)(-srel) : 0u;
        reach = max(reach, before);
        const uint32_t nmark = min(len, before);                      // leading bytes that come from before the segment
        const uint32_t len2 = len - nmark, off2 = off + nmark;        // the rest copies segment bytes
        const bool rest = is_m && len2 != 0;
        const bool is_far = rest && dist + (total - off2) > kSegRing; // source older than what the ring still holds at the end of the step
        const bool isfree = rest && !is_far && dist >= off + len;     // source ends before the step: needs no ordering
        const bool inord = rest && !is_far && !isfree;
        const uint32_t fm = __ballot_sync(0xFFFFFFFFu, nmark != 0), fc = __ballot_sync(0xFFFFFFFFu, isfree), om = __ballot_sync(0xFFFFFFFFu, inord);
        const uint32_t nfm = (uint32_t)__popc(fm), nfree = nfm + (uint32_t)__popc(fc), nord = (uint32_t)__popc(om);
        const uint32_t d2 = (a0 + pos + off2) & kSegMask;
        if (nmark) fq[__popc(fm & lt)] = make_uint2(d0 | (nmark << 16), kMarker | (before - 1u));                // value of byte k: second word - k
        if (isfree) fq[nfm + __popc(fc & lt)] = make_uint2(d2 | (len2 << 16), ((d2 - dist) & kSegMask) | (dist << 16));
        if (inord) oq[__popc(om & lt)] = make_uint2(d2 | (len2 << 16), ((d2 - dist) & kSegMask) | (dist << 16));
        if (is_far) {
            // (segments much longer than the ring only) the source left the ring: it was written through to sym16 in an earlier step
            const uint16_t *gs = gsym + (pos + off2 - dist);
            for (uint32_t k = 0; k < len2; k++) ring[(d2 + k) & kSegMask] = __ldcg(gs + k);
        }
        __syncwarp();
        // ---- order-free work: markers (no reads at all) and copies whose source is older than the step; these never overlap
        for (uint32_t i = 0; i < nfree; i += 4) {
            const uint2 c = fq[min(i + sub, nfree - 1u)];
            const uint32_t mlen = i + sub < nfree ? c.x >> 16 : 0u, md = c.x & 0xFFFFu, ms = c.y & 0xFFFFu;
            const bool iscopy = (c.y >> 16) != 0;
            for (uint32_t k = l8; __any_sync(0xFFFFFFFFu, k < mlen); k += 8) {
                const uint32_t v = iscopy ? (uint32_t)ring[(ms + k) & kSegMask] : ms - k;
                if (k < mlen) ring[(md + k) & kSegMask] = (uint16_t)v;
            }
        }
        __syncwarp();
        // ---- copies that read bytes produced in this step, in token order, each spread over the lanes
        uint2 nx = oq[0];
        for (uint32_t i = 0; i < nord; i++) {
            const uint2 c = nx;
            nx = oq[i + 1];
            const uint32_t mlen = c.x >> 16, mdist = c.y >> 16, md = c.x & 0xFFFFu, ms = c.y & 0xFFFFu;
            if (mdist >= mlen && mlen <= 32) {
                const uint16_t v = ring[(ms + lane) & kSegMask];
                if (lane < mlen) ring[(md + lane) & kSegMask] = v;
            } else if (mdist >= 32) {                                 // every 32-symbol slice reads only earlier slices
                for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
                    const uint16_t v = ring[(ms + k0 + lane) & kSegMask];
                    if (k0 + lane < mlen) ring[(md + k0 + lane) & kSegMask] = v;
                    __syncwarp();
                }
            } else {                                                  // short period: symbol k repeats symbol k mod dist of the source
                const uint32_t inv = kInvPeriod[mdist];
                for (uint32_t k = lane; k < mlen; k += 32) ring[(md + k) & kSegMask] = ring[(ms + k - mdist * ((k * inv) >> 16)) & kSegMask];
            }
            __syncwarp();
        }
        // ---- write the finished 8-symbol groups through to sym16 (16-byte stores), once enough of them have piled up: at the
        // start of a step less than kSegLazy symbols are unwritten, so the ring (step output <= kSegStepMax) never overwrites them
        // and a far source (older than pos + total - kSegRing) is always in sym16
        pos += total;
        if (pos >= lead && pos - flushed >= kSegLazy) {
            if (flushed < lead) {                                     // the unaligned head, once
                if (lane < lead) gsym[lane] = ring[(a0 + lane) & kSegMask];
                flushed = lead;
            }
            const uint32_t target = pos - ((pos - lead) & 7u);        // last group boundary at or below pos
            if (target > flushed) {
                const uint32_t ng = (target - flushed) >> 3;
                for (uint32_t g = lane; g < ng; g += 32)
                    *reinterpret_cast<uint4 *>(gsym + flushed + 8u * g) = *reinterpret_cast<const uint4 *>(ring + ((a0 + flushed + 8u * g) & kSegMask));
                flushed = target;
            }
        }
        i0 += take;
        __syncwarp();
    }
    if (pos >= lead) {                                                // what is left: head, whole groups, then the tail symbol by symbol
        if (flushed < lead) { if (lane < lead) gsym[lane] = ring[(a0 + lane) & kSegMask]; flushed = lead; }
        const uint32_t target = pos - ((pos - lead) & 7u);
        for (uint32_t g = lane; g < ((target - flushed) >> 3); g += 32)
            *reinterpret_cast<uint4 *>(gsym + flushed + 8u * g) = *reinterpret_cast<const uint4 *>(ring + ((a0 + flushed + 8u * g) & kSegMask));
        flushed = target;
    }
    for (uint32_t k = flushed + lane; k < pos; k += 32) gsym[k] = ring[(a0 + k) & kSegMask];
    if (pos != nout) err |= 1u;
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    reach = __reduce_max_sync(0xFFFFFFFFu, reach);
    if (lane == 0) { S.seg_reach[slot] = reach; if (err) atomicOr(S.mem_err + mem, err); }
}

// seg_cut[c] = 1 when no segment at or after c (inside c's part) reads a byte before c's first one -- only segments that start
// less than 32 KiB after it can.  A member's first segment is a cut by construction (a reach before it is an error, flagged by
// k_seg_resolve).  The first segment of a part starts a chain in any case: the parts are processed one after the other, so
// everything before it is final and its CTA preloads the window from out ("warm" start).
// Also packs what k_seg_subst needs of a slot into one 16-byte record (out offset, bytes | warm << 31, start | member << 1) and
// appends every chain start to the list of its part.
__global__ void __launch_bounds__(128) k_seg_cuts(SpecDev S, uint32_t part) {
    const uint32_t lo = S.part_slot0[part], hi = S.part_slot0[part + 1];
    const uint32_t c = lo + blockIdx.x * 128 + threadIdx.x;
    if (c >= hi) return;
    const uint64_t A = S.seg_out[c];
    const uint32_t m = S.seg_member[c], n = S.seg_nout[c];
    bool cut = n != 0;
    if (n) {
        for (uint32_t s = c; s < hi && S.seg_member[s] == m; s++) {
            if (!S.seg_nout[s]) continue;
            const uint64_t o = S.seg_out[s];
            if (o >= A + 32768u) break;
            if ((uint64_t)S.seg_reach[s] > o - A) { cut = false; break; }
        }
    }
    bool first = n != 0;                                              // first non-empty slot of the part?
    for (uint32_t s = c; first && s > lo; s--) if (S.seg_nout[s - 1]) first = false;
    const bool warm = first && !cut;
    S.seg_cut[c] = (cut || first) ? 1 : 0;
    S.seg_rec[c] = make_uint4((uint32_t)A, (uint32_t)(A >> 32), n | (warm ? 0x80000000u : 0u), ((cut || first) ? 1u : 0u) | (m << 1));
    if (cut || first) S.chain_list[lo + atomicAdd(S.chain_count + part, 1u)] = c;
}

// Soft chain starts (see kSuperBytes): a non-empty slot that is not a chain start already and whose first byte lies in another
// kSuperBytes-aligned window of out than its predecessor's.  Every chain start (hard or soft) also learns where its chain ends and
// whether a soft chain continues it; hard starts with a soft successor go to the tail list (k_soft_tails walks from them).
__device__ __forceinline__ bool seg_is_soft(uint64_t A, uint64_t A_prev) { return A / kSuperBytes != A_prev / kSuperBytes; }
__global__ void __launch_bounds__(128) k_seg_soft(SpecDev S, uint32_t part) {
    const uint32_t lo = S.part_slot0[part], hi = S.part_slot0[part + 1];
    const uint32_t c = lo + blockIdx.x * 128 + threadIdx.x;
    if (c >= hi || !S.seg_nout[c]) return;
    const uint32_t m = S.seg_member[c];
    const uint64_t A = S.seg_out[c];
    bool hard = S.seg_cut[c] != 0, soft = false;
    if (!hard) {
        uint32_t p = c;                                               // previous non-empty slot (exists: the part's first one is a hard start)
        while (p > lo && !S.seg_nout[p - 1]) p--;
        if (p > lo) soft = seg_is_soft(A, S.seg_out[p - 1]);
    }
    if (!hard && !soft) return;
    // walk forward to the next chain start of any kind
    uint64_t prevA = A, end = A + S.seg_nout[c];
    uint32_t next_soft = kNoSlot;
    for (uint32_t s = c + 1; s < hi && S.seg_member[s] == m; s++) {
        const uint32_t n = S.seg_nout[s];
        if (!n) continue;
        const uint64_t a = S.seg_out[s];
        if (S.seg_cut[s]) break;
        if (seg_is_soft(a, prevA)) { next_soft = s; break; }
        prevA = a; end = a + n;
    }
    S.chain_end[c] = end; S.chain_next_soft[c] = next_soft;
    if (soft) {
        uint4 r = S.seg_rec[c];
        r.z |= 0x40000000u; r.w |= 1u;                                // soft | chain start
        S.seg_rec[c] = r;
        S.soft_list[lo + atomicAdd(S.chain_count + 2 * kMaxParts + part, 1u)] = c;
    } else if (next_soft != kNoSlot) S.tail_list[lo + atomicAdd(S.chain_count + 4 * kMaxParts + part, 1u)] = c;
}

// Persistent CTAs take chains (the slots from a cut up to the next cut of the same member) from the part's list.  The last
// kSubRing final bytes of the chain are kept in shared memory at index (offset - base) mod kSubRing, base = chain start rounded
// down to 16, so that aligned 16-byte groups of out are aligned groups of the window.  The chain is walked slot by slot, one
// 16-byte group per thread; everything the next step needs is requested one step ahead (the slot records two steps ahead, the
// symbols one), and the sixteen window reads of a group are issued unconditionally, so a step costs little more than its barrier.
constexpr uint32_t kSubThreads = 512, kSubRing = 49152, kSubGroups = kSubThreads;
__device__ __forceinline__ uint32_t sub_cap(uint64_t B) { return kSubGroups * 16u - ((uint32_t)B & 15u); }   // bytes of one step from offset B

// one step: bytes [B, B + bl) of the segment that starts at A.  kFirst: B == A (every marker target is still in the window).
template <bool kFirst>
__device__ __forceinline__ void sub_step(uint8_t *out, uint8_t *win, bool vec, uint64_t A, uint64_t B, uint32_t bl, uint32_t rA, uint32_t rB,
                                         const uint4 xa, const uint4 xb, uint32_t tid) {
    const uint32_t lead16 = (uint32_t)B & 15u;
    const uint32_t ng = (lead16 + bl + 15u) >> 4;
    if (tid >= ng) return;
    const uint64_t p = (B - lead16) + 16ull * tid;                    // offset in out of this thread's 16-byte group
    uint32_t rp = rB + 16u * tid + kSubRing - lead16;                 // window index of p (p may lie below B: + kSubRing first)
    rp -= rp >= 2 * kSubRing ? 2 * kSubRing : rp >= kSubRing ? kSubRing : 0u;
    const int32_t thr = kFirst ? 0x7FFF : (int32_t)kSubRing - 1 - (int32_t)(B - A) - (int32_t)bl;   // markers above it have left the window
    const uint32_t w[8] = { xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w };
    uint32_t v[16], t[16];
#pragma unroll
    for (uint32_t q = 0; q < 16; q++) {
        v[q] = (w[q >> 1] >> (16u * (q & 1u))) & 0xFFFFu;
        int32_t r = (int32_t)rA - 1 - (int32_t)(v[q] & 0x7FFFu);
        r += (r >> 31) & (int32_t)kSubRing;
        t[q] = win[r];                                                // unconditional: the index is always inside the window
    }
    const uint32_t lo = tid == 0 ? lead16 : 0u;                       // symbols [lo, hi) of the group belong to the step
    const uint32_t hi = min(16u, lead16 + bl - 16u * tid);
    if (!kFirst) {
#pragma unroll
        for (uint32_t q = 0; q < 16; q++)                             // (steps after the first of a long segment only)
            if (q >= lo && q < hi && (v[q] & kMarker) && (int32_t)(v[q] & 0x7FFFu) > thr) t[q] = out[A - 1 - (uint64_t)(v[q] & 0x7FFFu)];
    }
    uint32_t o[4] = { 0, 0, 0, 0 };
#pragma unroll
    for (uint32_t q = 0; q < 16; q++) {
        v[q] = ((v[q] & kMarker) ? t[q] : v[q]) & 0xFFu;
        o[q >> 2] |= v[q] << (8u * (q & 3u));
    }
    if (lo == 0 && hi == 16) {
        const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<uint4 *>(win + rp) = ov;
        if (vec) *reinterpret_cast<uint4 *>(out + p) = ov;
        else {
#pragma unroll
            for (uint32_t q = 0; q < 16; q++) out[p + q] = (uint8_t)v[q];
        }
    } else if (vec && (lo == 0 || hi == 16)) {
        // first / last group of the step: only bytes [lo, hi) are this segment's.  The window is private to the CTA: merge and store 16
        // bytes.  out is shared with the neighbouring chains' CTAs: at most four naturally aligned stores cover exactly [lo, hi).
        uint4 old = *reinterpret_cast<const uint4 *>(win + rp);
        uint32_t ow[4] = { old.x, old.y, old.z, old.w };
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
            const uint32_t b0 = 4 * k, b1 = 4 * k + 4;                // bytes of word k inside [lo, hi) -> byte mask
            const uint32_t a = max(lo, b0), e = min(hi, b1);
            const uint32_t m = a < e ? (e - a == 4 ? 0xFFFFFFFFu : (((1u << (8 * (e - a))) - 1u) << (8 * (a - b0)))) : 0u;
            ow[k] = (ow[k] & ~m) | (o[k] & m);
        }
        *reinterpret_cast<uint4 *>(win + rp) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        uint8_t *g = out + p;
        if (lo == 0) {                                                // bytes [0, hi): 8 + 4 + 2 + 1
            uint32_t at = 0;
            if (hi & 8) { *reinterpret_cast<uint2 *>(g) = make_uint2(o[0], o[1]); at = 8; }
            if (hi & 4) { *reinterpret_cast<uint32_t *>(g + at) = o[at >> 2]; at += 4; }
            if (hi & 2) { *reinterpret_cast<uint16_t *>(g + at) = (uint16_t)(o[at >> 2] >> (8 * (at & 3))); at += 2; }
            if (hi & 1) g[at] = (uint8_t)(o[at >> 2] >> (8 * (at & 3)));
        } else {                                                      // bytes [lo, 16): 1 + 2 + 4 + 8
            uint32_t at = lo;
            if (at & 1) { g[at] = (uint8_t)(o[at >> 2] >> (8 * (at & 3))); at += 1; }
            if (at & 2) { *reinterpret_cast<uint16_t *>(g + at) = (uint16_t)(o[at >> 2] >> (8 * (at & 3))); at += 2; }
            if (at & 4) { *reinterpret_cast<uint32_t *>(g + at) = o[at >> 2]; at += 4; }
            if (at & 8) *reinterpret_cast<uint2 *>(g + at) = make_uint2(o[2], o[3]);
        }
    } else {
#pragma unroll
        for (uint32_t q = 0; q < 16; q++)
            if (q >= lo && q < hi) { out[p + q] = (uint8_t)v[q]; uint32_t r = rp + q; if (r >= kSubRing) r -= kSubRing; win[r] = (uint8_t)v[q]; }
    }
}
// The same step for a SOFT chain (history before S0 unknown): the window holds 16-bit symbols (a byte, or a marker relative to the
// chain's own start S0), and the step's result goes back into sym16 in place; k_soft_tails / k_soft_rest turn it into bytes.
template <bool kFirst>
__device__ __forceinline__ void sub_step16(uint16_t *sym, uint16_t *win, uint64_t S0, uint64_t A, uint64_t B, uint32_t bl, uint32_t rA, uint32_t rB,
                                           const uint4 xa, const uint4 xb, uint32_t tid) {
    const uint32_t lead16 = (uint32_t)B & 15u;
    const uint32_t ng = (lead16 + bl + 15u) >> 4;
    if (tid >= ng) return;
    const uint64_t p = (B - lead16) + 16ull * tid;
    uint32_t rp = rB + 16u * tid + kSubRing - lead16;
    rp -= rp >= 2 * kSubRing ? 2 * kSubRing : rp >= kSubRing ? kSubRing : 0u;
    const int32_t thr = kFirst ? 0x7FFF : (int32_t)kSubRing - 1 - (int32_t)(B - A) - (int32_t)bl;
    const uint32_t w[8] = { xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w };
    const uint32_t lo = tid == 0 ? lead16 : 0u, hi = min(16u, lead16 + bl - 16u * tid);
    uint32_t o[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
#pragma unroll
    for (uint32_t q = 0; q < 16; q++) {
        uint32_t v = (w[q >> 1] >> (16u * (q & 1u))) & 0xFFFFu;
        int32_t r = (int32_t)rA - 1 - (int32_t)(v & 0x7FFFu);
        r += (r >> 31) & (int32_t)kSubRing;
        uint32_t t = win[r];
        if (!kFirst && (v & kMarker) && (int32_t)(v & 0x7FFFu) > thr && q >= lo && q < hi) {      // the target left the window
            const uint64_t tg = A - 1 - (uint64_t)(v & 0x7FFFu);
            t = tg >= S0 ? (uint32_t)sym[tg] : kMarker | (uint32_t)(S0 - 1 - tg);
        }
        v = (v & kMarker) ? t : v;
        o[q >> 1] |= v << (16u * (q & 1u));
    }
    if (lo == 0 && hi == 16) {
        reinterpret_cast<uint4 *>(win + rp)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4 *>(win + rp)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        reinterpret_cast<uint4 *>(sym + p)[0] = make_uint4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<uint4 *>(sym + p)[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (uint32_t q = 0; q < 16; q++)
            if (q >= lo && q < hi) {
                const uint16_t v = (uint16_t)(o[q >> 1] >> (16u * (q & 1u)));
                uint32_t r = rp + q; if (r >= kSubRing) r -= kSubRing;
                win[r] = v; sym[p + q] = v;
            }
    }
}
__device__ __forceinline__ void sub_load(const uint16_t *__restrict__ sym, uint64_t B, uint32_t bl, uint32_t tid, uint4 &xa, uint4 &xb) {
    const uint32_t lead16 = (uint32_t)B & 15u;
    if (tid < ((lead16 + bl + 15u) >> 4)) {
        const uint4 *p = reinterpret_cast<const uint4 *>(sym + (B - lead16) + 16ull * tid);
        xa = __ldg(p); xb = __ldg(p + 1);
    }
}
// mbarrier + bulk-copy (TMA engine, 1-D) helpers: the symbols of the next segments are fetched into shared memory by the copy
// engine while the CTA works on the current one; no thread waits for HBM on the critical path
__device__ __forceinline__ void sb_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void sb_expect(uint32_t a, uint32_t bytes) { asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" :: "r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void sb_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void sb_wait(uint32_t a, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 1000; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}

constexpr uint32_t kSubStages = 3;                                    // segments in flight (symbols prefetched by the copy engine)
constexpr uint32_t kSubStageBytes = kSubGroups * 32;                  // 16 symbols (32 bytes) per consumer thread
constexpr uint32_t kSubTail = kSubStages * kSubStageBytes + kSubStages * 16 + 2 * kSubStages * 8;   // stages | descriptors | full + empty barriers
constexpr uint32_t kSubSmem = kSubRing + kSubTail, kSubSmemSoft = 2 * kSubRing + kSubTail;              // window of bytes / of 16-bit symbols
constexpr uint32_t kSubCtaThreads = kSubThreads + 32;                 // 16 consumer warps + 1 producer warp
constexpr uint32_t kDescStart = 1u, kDescWarm = 2u, kDescExit = 4u;   // descriptor flags (word w; the member index sits above bit 3)
__device__ __forceinline__ void sb_arrive(uint32_t a) { asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(a) : "memory"); }
__device__ __forceinline__ void sub_bar() { asm volatile("bar.sync 1, %0;" :: "n"(kSubThreads) : "memory"); }      // the consumer warps only

// Persistent CTAs: ONE producer thread walks the chains of the part (taken from the part's list with an atomic counter), and for
// every segment posts a descriptor and starts a bulk copy (TMA engine) of its symbols into one of kSubStages shared-memory stages;
// 512 consumer threads wait for the stage (mbarrier, complete_tx), substitute the markers out of the window and hand the stage back
// (second set of mbarriers).  The producer runs up to four segments ahead -- across chain ends too -- so neither the slot records
// (dependent L2 reads) nor the symbols (HBM) are ever waited for on the consumers' critical path: a step is ~100 instructions per
// warp plus one barrier.
template <bool kSoft>
__global__ void __launch_bounds__(kSubCtaThreads, kSoft ? 1 : 2) k_seg_subst(SpecDev S, uint32_t part) {
    extern __shared__ __align__(128) uint8_t ssm[];
    constexpr uint32_t kWinBytes = kSoft ? 2 * kSubRing : kSubRing;
    uint8_t *win = ssm;
    uint8_t *stage = ssm + kWinBytes;
    uint4 *desc = reinterpret_cast<uint4 *>(ssm + kWinBytes + kSubStages * kSubStageBytes);          // what each stage holds
    const uint32_t full_a = (uint32_t)__cvta_generic_to_shared(ssm + kWinBytes + kSubStages * kSubStageBytes + kSubStages * 16);
    const uint32_t empty_a = full_a + kSubStages * 8;
    const uint32_t stage_a = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
        for (uint32_t i = 0; i < kSubStages; i++) { sb_init(full_a + 8u * i, 1u); sb_init(empty_a + 8u * i, 1u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= kSubThreads) {
        // ------------------------------------------------------------------ producer (one thread)
        if (tid != kSubThreads) return;
        const uint16_t *__restrict__ sym = S.sym16;
        const uint4 *__restrict__ rec = S.seg_rec;                   // n_slots + 2 entries
        const uint32_t *__restrict__ list = (kSoft ? S.soft_list : S.chain_list) + S.part_slot0[part];
        const uint32_t nchains = S.chain_count[(kSoft ? 2 : 0) * kMaxParts + part];
        const uint32_t part_hi = S.part_slot0[part + 1];
        uint32_t use = 0;
        auto post = [&](uint32_t x, uint32_t y, uint32_t n, uint32_t flags) {
            const uint32_t k = use % kSubStages;
            if (use >= kSubStages) sb_wait(empty_a + 8u * k, ((use / kSubStages) - 1u) & 1u);      // the consumers are done with this stage's previous use
            desc[k] = make_uint4(x, y, n, flags);
            uint32_t bytes = 0;
            if (n) {
                const uint64_t A = x | ((uint64_t)y << 32);
                const uint32_t lead16 = (uint32_t)A & 15u;
                bytes = ((lead16 + min(n, sub_cap(A)) + 15u) >> 4) * 32u;
            }
            sb_expect(full_a + 8u * k, bytes);                        // release: the descriptor is visible to whoever sees the phase complete
            if (bytes) sb_bulk_g2s(stage_a + k * kSubStageBytes, sym + ((x | ((uint64_t)y << 32)) & ~15ull), bytes, full_a + 8u * k);
            use++;
        };
        for (;;) {
            const uint32_t ci = atomicAdd(S.chain_count + (kSoft ? 3 : 1) * kMaxParts + part, 1u);
            if (ci >= nchains) break;
            uint32_t s = list[ci];
            uint4 r0 = rec[s], r1 = rec[s + 1], r2 = rec[s + 2];
            const uint32_t mem = r0.w >> 1;
            if (S.mem_err[mem]) continue;                             // the member goes to the in-order kernel: markers may point anywhere
            post(r0.x, r0.y, r0.z & 0x3FFFFFFFu, kDescStart | ((r0.z & 0x80000000u) ? kDescWarm : 0u) | (mem << 3));
            for (;;) {
                s++;
                r0 = r1; r1 = r2; r2 = rec[s + 2];                   // (two records are always on their way)
                if (s >= part_hi || (r0.w & 1u) || (r0.w >> 1) != mem) break;
                if (r0.z) post(r0.x, r0.y, r0.z, mem << 3);           // (empty slots carry no start flag: walked through)
            }
        }
        post(0, 0, 0, kDescExit);
        return;
    }
    // ---------------------------------------------------------------------- consumers
    uint8_t *out = S.out;                                             // read back by later segments of the chain: no __restrict__
    const bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    uint16_t *sym = S.sym16;
    uint16_t *win16 = reinterpret_cast<uint16_t *>(win);
    uint32_t rA = 0;                                                  // window index of the current segment's first byte
    uint64_t S0 = 0;                                                  // soft chains: the chain's first byte
    for (uint32_t use = 0;; use++) {
        const uint32_t k = use % kSubStages;
        sb_wait(full_a + 8u * k, (use / kSubStages) & 1u);
        const uint4 d = desc[k];
        if (d.w & kDescExit) break;
        const uint64_t A = d.x | ((uint64_t)d.y << 32);
        const uint32_t n = d.z;
        if (d.w & kDescStart) {
            rA = d.x & 15u;                                           // base of the window = chain start rounded down to 16
            if (kSoft) {
                // nothing is known about the bytes before the chain: the window answers "marker relative to the chain start"
                S0 = A;
                for (uint32_t q = tid; q < 32768u; q += kSubThreads) {
                    int32_t r = (int32_t)rA - 1 - (int32_t)q; r += (r >> 31) & (int32_t)kSubRing;
                    win16[r] = (uint16_t)(kMarker | q);
                }
                sub_bar();
            } else if (d.w & kDescWarm) {
                // warm start (first chain of a later part): the 32 KiB before the segment are final in out -- preload them
                const uint64_t m0 = S.mem_out_off[d.w >> 3];
                const uint32_t back = (uint32_t)min((uint64_t)32768u, A - m0);
                for (uint32_t q = tid; q < back; q += kSubThreads) {  // byte q+1 before A -> window index rA - 1 - q (mod kSubRing)
                    int32_t r = (int32_t)rA - 1 - (int32_t)q; r += (r >> 31) & (int32_t)kSubRing;
                    win[r] = out[A - 1 - q];
                }
                sub_bar();
            }
        }
        const uint32_t bl0 = min(n, sub_cap(A));
        {
            const uint4 *sp = reinterpret_cast<const uint4 *>(stage + k * kSubStageBytes) + 2 * tid;
            const bool mine = tid < ((((uint32_t)A & 15u) + bl0 + 15u) >> 4);
            const uint4 xa = mine ? sp[0] : make_uint4(0, 0, 0, 0), xb = mine ? sp[1] : make_uint4(0, 0, 0, 0);
            if (kSoft) sub_step16<true>(sym, win16, S0, A, A, bl0, rA, rA, xa, xb, tid);
            else sub_step<true>(out, win, vec, A, A, bl0, rA, rA, xa, xb, tid);
        }
        sub_bar();                                                    // the window is complete; every consumer is done with the stage
        if (tid == 0) sb_arrive(empty_a + 8u * k);
        for (uint32_t b0 = bl0; b0 < n;) {                            // (segments longer than one step: highly compressible data)
            const uint64_t B = A + b0;
            const uint32_t bl = min(n - b0, sub_cap(B));
            uint32_t rB = rA + b0; rB -= (rB / kSubRing) * kSubRing;
            uint4 za = make_uint4(0, 0, 0, 0), zb = za;
            sub_load(sym, B, bl, tid, za, zb);
            if (kSoft) sub_step16<false>(sym, win16, S0, A, B, bl, rA, rB, za, zb, tid);
            else sub_step<false>(out, win, vec, A, B, bl, rA, rB, za, zb, tid);
            sub_bar();
            b0 += bl;
        }
        rA += n - (n / kSubRing) * kSubRing; if (rA >= kSubRing) rA -= kSubRing;
    }
}

// sym16 of a soft chain after k_seg_subst<true> -> bytes: a byte as it is, a marker d = the byte d+1 before the chain's start S0.
// The markers of a chain can only point into the last 32 KiB before S0, i.e. into the TAIL of the chain before it.
__device__ __forceinline__ void soft_finish_range(const SpecDev &S, uint64_t S0, uint64_t lo, uint64_t hi, uint32_t tid, uint32_t nthreads) {
    uint8_t *out = S.out;
    const uint16_t *__restrict__ sym = S.sym16;
    const bool vec = (reinterpret_cast<uintptr_t>(out) & 15u) == 0;
    const uint64_t g0 = lo & ~15ull;
    const uint64_t ng = (hi + 15 - g0) >> 4;
    for (uint64_t g = tid; g < ng; g += nthreads) {
        const uint64_t p = g0 + 16 * g;
        const uint4 xa = __ldg(reinterpret_cast<const uint4 *>(sym + p)), xb = __ldg(reinterpret_cast<const uint4 *>(sym + p) + 1);
        const uint32_t w[8] = { xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w };
        uint32_t o[4] = { 0, 0, 0, 0 }, v[16];
#pragma unroll
        for (uint32_t q = 0; q < 16; q++) {
            v[q] = (w[q >> 1] >> (16u * (q & 1u))) & 0xFFFFu;
            if ((v[q] & kMarker) && p + q >= lo && p + q < hi) v[q] = out[S0 - 1 - (uint64_t)(v[q] & 0x7FFFu)];
            v[q] &= 0xFFu;
            o[q >> 2] |= v[q] << (8u * (q & 3u));
        }
        if (vec && p >= lo && p + 16 <= hi) *reinterpret_cast<uint4 *>(out + p) = make_uint4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (uint32_t q = 0; q < 16; q++) if (p + q >= lo && p + q < hi) out[p + q] = (uint8_t)v[q];
        }
    }
}
// Serial part of the soft chains: one CTA per hard chain that is continued by soft chains walks them in order and finishes the
// last 32 KiB of each (whose markers point into the previous chain's last 32 KiB, final by then).
__global__ void __launch_bounds__(512) k_soft_tails(SpecDev S, uint32_t part) {
    __shared__ uint32_t item;
    const uint32_t n = S.chain_count[4 * kMaxParts + part];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) item = atomicAdd(S.chain_count + 5 * kMaxParts + part, 1u);
        __syncthreads();
        if (item >= n) return;
        const uint32_t c0 = S.tail_list[S.part_slot0[part] + item];
        if (S.mem_err[S.seg_member[c0]]) continue;
        for (uint32_t c = S.chain_next_soft[c0]; c != kNoSlot; c = S.chain_next_soft[c]) {
            const uint64_t S0 = S.seg_out[c], S1 = S.chain_end[c];
            soft_finish_range(S, S0, S1 - S0 > 32768u ? S1 - 32768u : S0, S1, threadIdx.x, 512);
            __syncthreads();                                          // this tail is what the next chain's markers read
        }
    }
}
// Parallel part: everything of a soft chain before its last 32 KiB, 64 KiB per work item.
__global__ void __launch_bounds__(256) k_soft_rest(SpecDev S, uint32_t part) {
    const uint32_t ns = S.chain_count[2 * kMaxParts + part];
    constexpr uint32_t kItem = 65536, kItems = (uint32_t)(kSuperBytes / kItem) + 4;     // a soft chain is at most ~kSuperBytes + one segment long
    for (uint64_t it = blockIdx.x; it < (uint64_t)ns * kItems; it += gridDim.x) {
        const uint32_t c = S.soft_list[S.part_slot0[part] + (uint32_t)(it / kItems)], j = (uint32_t)(it % kItems);
        if (S.mem_err[S.seg_member[c]]) continue;
        const uint64_t S0 = S.seg_out[c], S1 = S.chain_end[c];
        const uint64_t end = S1 - S0 > 32768u ? S1 - 32768u : S0;     // the tail belongs to k_soft_tails
        uint64_t lo = S0 + (uint64_t)j * kItem;
        if (lo >= end) continue;
        // (a chain much longer than kSuperBytes -- one giant segment -- is finished by its last item)
        const uint64_t hi = j + 1 == kItems ? end : min(end, lo + kItem);
        soft_finish_range(S, S0, lo, hi, threadIdx.x, 256);
    }
}

// ---------------------------------------------------------------------------------- launchers
cudaError_t spec_init_attributes() {
    cudaError_t e = cudaFuncSetAttribute(k_spec_headers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(InflateTables)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_seg_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSegSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_seg_subst<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSubSmemSoft);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_seg_subst<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSubSmem);
}
// Rounds [r0, r1) of the Jacobi iteration, then the verification.  r0 = 0 also parses the block headers; a later call continues
// an iteration that had not converged (regions whose codes do not self-synchronise settle one subsegment per round).
cudaError_t spec_launch_parse(const SpecDev &S, uint32_t r0, uint32_t r1, cudaStream_t st) {
    if (!S.n_blocks || r1 <= r0) return cudaSuccess;
    cudaError_t e;
    if (r0 == 0) {
        k_spec_headers<<<(S.n_blocks + 3) / 4, 128, 4 * sizeof(InflateTables), st>>>(S);
        e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    SpecDev R = S;
    for (uint32_t r = r0; r < r1; r++) {
        // round r reads the exits of round r-1 (s_exit_prev) and writes s_exit
        if (r & 1) { R.s_exit = S.s_exit_prev; R.s_exit_prev = S.s_exit; } else { R.s_exit = S.s_exit; R.s_exit_prev = S.s_exit_prev; }
        k_spec_round<<<S.n_ctas, kSpecCta, 0, st>>>(R, r);
        e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    if (((r1 - 1) & 1)) { R.s_exit = S.s_exit_prev; R.s_exit_prev = S.s_exit; } else { R.s_exit = S.s_exit; R.s_exit_prev = S.s_exit_prev; }
    k_spec_verify<<<S.n_blocks, 256, 0, st>>>(R);
    return cudaGetLastError();
}
cudaError_t spec_launch_tokens(const SpecDev &S, uint32_t cta_lo, uint32_t cta_hi, cudaStream_t st) {
    if (cta_hi <= cta_lo) return cudaSuccess;
    k_spec_tokens<<<cta_hi - cta_lo, kSpecCta, 0, st>>>(S, cta_lo);
    return cudaGetLastError();
}
cudaError_t spec_launch_segments(const SpecDev &S, uint32_t part, cudaStream_t st) {
    const uint32_t lo = S.part_slot0[part], hi = S.part_slot0[part + 1];
    if (hi <= lo) return cudaSuccess;
    k_seg_plan<<<(hi - lo + 127) / 128, 128, 0, st>>>(S, lo, hi);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_seg_resolve<<<hi - lo, 32, kSegSmem, st>>>(S, lo);
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_seg_cuts<<<(hi - lo + 127) / 128, 128, 0, st>>>(S, part);
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_seg_soft<<<(hi - lo + 127) / 128, 128, 0, st>>>(S, part);
    return cudaGetLastError();
}
cudaError_t spec_launch_subst(const SpecDev &S, uint32_t part, cudaStream_t st) {
    if (part >= S.n_parts || S.part_slot0[part + 1] <= S.part_slot0[part]) return cudaSuccess;
    const uint32_t slots = S.part_slot0[part + 1] - S.part_slot0[part];
    k_seg_subst<false><<<slots < 148u * 2u ? slots : 148u * 2u, kSubCtaThreads, kSubSmem, st>>>(S, part);      // persistent: one chain at a time per CTA
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    // soft chains (streams without natural cuts only; the three kernels return at once when there are none)
    const uint32_t nsoft_max = (uint32_t)std::min<uint64_t>(148, (uint64_t)slots * kSegBytes / kSuperBytes + 1);
    k_seg_subst<true><<<nsoft_max, kSubCtaThreads, kSubSmemSoft, st>>>(S, part);
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_soft_tails<<<std::min<uint32_t>(148u, nsoft_max), 512, 0, st>>>(S, part);
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_soft_rest<<<148 * 4, 256, 0, st>>>(S, part);
    return cudaGetLastError();
}

}  // namespace b2f

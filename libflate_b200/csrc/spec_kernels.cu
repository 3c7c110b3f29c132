// spec_kernels.cu -- sub-block parallel inflate (sm_100a).
//
// A DEFLATE block is one serial bit stream, but Huffman parses self-synchronise: a decoder started at an arbitrary bit
// falls into step with the true parse after a few symbols.  Each candidate block (from the boundary finder) is cut into
// 4096-bit subsegments, one THREAD each:
//   k_spec_headers   warp per block : parse the dynamic header, build the decode tables (kept in HBM, 12 KiB per block)
//   k_spec_round     thread per subsegment, Jacobi iteration: round 0 decodes from the subsegment's first bit (the first
//                    subsegment from the true first symbol); round r restarts from the exit of the left neighbour of
//                    round r-1 if that differs from the start used so far.  Fixed point == the true parse.
//   k_spec_verify    CTA per block  : checks that every start equals its left neighbour's exit (=> by induction the true
//                    parse), finds EndOfBlock, exclusive scans of symbol/byte counts
//   k_spec_tokens    thread per subsegment: final decode from the true start, writes one token per symbol
//   k_spec_resolve   warp per block : LZ77 resolution of the token stream, 32 tokens per step, 64 KiB output ring in
//                    shared memory (history window), multi-round resolution of matches that depend on each other,
//                    coalesced 32 KiB flushes to HBM
// Nothing here is trusted blindly: the host accepts a block only if its verified EndOfBlock lands exactly on the next
// block of the chain; otherwise the stream goes to the exact in-order kernel (decode_kernels.cu).
// Reference behaviour being reproduced: src/deflate/decode.rs:112-130, symbol.rs:193-243, libflate_lz77/src/lib.rs:164-194.
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
__global__ void __launch_bounds__(kSpecCta) k_spec_tokens(SpecDev S) {
    __shared__ __align__(16) InflateTables Ts;
    __shared__ __align__(16) uint32_t stage[kSpecCta * 8];
    const uint32_t b = owner_u32(S.blk_cta0, S.n_blocks, blockIdx.x);
    if (S.blk_sel[b] == 0) return;                               // not on the verified chain
    load_tables_smem(Ts, S.tabs + b);
    const uint32_t data_rel = S.blk_data_rel[b];
    const uint32_t k = (blockIdx.x - S.blk_cta0[b]) * kSpecCta + threadIdx.x;
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

// ---------------------------------------------------------------------------------- independent LZ77 units inside a block
// A subsegment boundary is a cut point when no later match of the block reaches back across it (libflate's 256 KiB LZ77
// chunks never reference each other: libflate_lz77/src/default.rs:73,108).  Units between cut points resolve in parallel.
__global__ void __launch_bounds__(32) k_spec_units(SpecDev S) {
    const uint32_t b = S.sel_blocks[blockIdx.x];
    if (threadIdx.x != 0) return;
    const uint32_t s0 = S.blk_seg0[b];
    const uint32_t k0 = S.blk_data_rel[b] / kSpecBits, e = S.blk_eob_seg[b];
    const uint32_t u0 = S.sel_unit0[blockIdx.x], umax = S.sel_unit0[blockIdx.x + 1] - u0;
    const uint64_t nout = S.blk_nout[b], ntok = S.blk_ntok[b];
    const uint32_t *__restrict__ tok = S.tokens + S.blk_tok0[b];
    // Right to left over the subsegments with `after` = lowest position read by any token of LATER subsegments.  A cut can
    // only lie inside subsegment k when after >= start(k); the exact token is found by walking k's tokens backwards.
    uint32_t nu = 0;
    int64_t after = INT64_MAX;
    uint64_t unit_end_out = nout, unit_end_tok = ntok;
    auto emit = [&](uint64_t o, uint64_t tkn) {
        const uint32_t slot = u0 + umax - 1 - nu;
        S.unit_out[slot] = o; S.unit_tok[slot] = tkn; S.unit_ntok[slot] = unit_end_tok - tkn; S.unit_nout[slot] = unit_end_out - o; S.unit_blk[slot] = b;
        unit_end_out = o; unit_end_tok = tkn; nu++;
    };
    for (uint32_t k = e + 1; k-- > k0;) {
        const uint64_t o = S.s_out_rel[s0 + k], t0 = S.s_tok_rel[s0 + k];
        const uint64_t o_next = k == e ? nout : S.s_out_rel[s0 + k + 1], t_next = k == e ? ntok : S.s_tok_rel[s0 + k + 1];
        if (k > k0 && after >= (int64_t)o && unit_end_out - o >= kUnitMinBytes && nu + 1 < umax && t_next > t0) {
            // walk the tokens of subsegment k backwards: cut before token t iff every token >= t reads at or after dst(t)
            int64_t run = after; uint64_t dst_end = o_next;
            for (uint64_t t = t_next; t-- > t0;) {
                const uint32_t tk = tok[t];
                const uint32_t len = (tk & kSymPtr) ? (tk >> 16) & 0x1FFu : (tk == kTokSkip ? 0u : 1u);
                const uint64_t dst = dst_end - len;
                if (tk & kSymPtr) { const int64_t src = (int64_t)dst - (int64_t)(tk & 0xFFFFu); if (src < run) run = src; }
                if (len && run >= (int64_t)dst && unit_end_out - dst >= kUnitMinBytes && (t > t0 || k > k0)) { emit(dst, t); break; }
                dst_end = dst;
            }
        }
        const int64_t ms = S.s_min_src[s0 + k];
        if (ms < after) after = ms;
    }
    emit(S.s_out_rel[s0 + k0], S.s_tok_rel[s0 + k0]);             // the block's first unit
    // compact to the front of the block's slot range (order does not matter); mark the rest empty
    for (uint32_t i = 0; i < nu; i++) {
        const uint32_t from = u0 + umax - nu + i, to = u0 + i;
        if (from != to) { S.unit_out[to] = S.unit_out[from]; S.unit_tok[to] = S.unit_tok[from]; S.unit_ntok[to] = S.unit_ntok[from]; S.unit_nout[to] = S.unit_nout[from]; S.unit_blk[to] = S.unit_blk[from]; }
    }
    for (uint32_t i = nu; i < umax; i++) S.unit_blk[u0 + i] = 0xFFFFFFFFu;
}

// ---------------------------------------------------------------------------------- LZ77 resolution (warp per block)
__device__ __forceinline__ uint32_t r_lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void r_sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint2 r_lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void r_sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" :: "r"(a), "r"(x), "r"(y) : "memory"); }
// ring[dst] = ring[src] for the lanes with on != 0, as predicated instructions (no divergent branch)
__device__ __forceinline__ void r_copy8_if(uint32_t on, uint32_t src, uint32_t dst) {
    asm volatile("{ .reg .pred p; .reg .u32 v; setp.ne.u32 p, %0, 0; @p ld.shared.u8 v, [%1]; @p st.shared.u8 [%2], v; }" :: "r"(on), "r"(src), "r"(dst) : "memory");
}
__device__ __forceinline__ uint32_t r_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// The ring keeps the most recent kResRing bytes; every step is written through to HBM, so sources older than the ring are
// read back from HBM/L2 (ld.global.cg).  24 KiB per warp => 9 resident warps per SM instead of 3 with a full 64 KiB window.
#ifndef B2F_RESOLVE_FREE_FIRST
#define B2F_RESOLVE_FREE_FIRST 0        // 1: copy the matches whose source ends before the step without ordering (modelled on the CPU against
#endif                                  // sequential LZ77; not yet measured on a GPU -- round-2 candidate)
constexpr uint32_t kResRing = 24576;
// ceil(65536 / p): k mod p == k - p * ((k * inv) >> 16) for p < 32, k < 400 (run-length matches; avoids a division per match)
__constant__ uint32_t kInvPeriod[32] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462, 5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428, 2341, 2260, 2185, 2115};
constexpr uint32_t kResSmem = kResRing + 34 * 8 + 32;            // + the per-step queue of match parameters + 32 write-only dummy bytes

__global__ void __launch_bounds__(32) k_spec_resolve(SpecDev S, uint32_t uoff) {
    extern __shared__ __align__(16) uint8_t ring[];
    const uint32_t lane = threadIdx.x;
    const uint32_t u = blockIdx.x + uoff;
    const uint32_t b = S.unit_blk[u];
    if (b == 0xFFFFFFFFu) return;                                // unused slot
    const uint32_t *__restrict__ tok = S.tokens + S.blk_tok0[b] + S.unit_tok[u];
    const uint64_t ntok = S.unit_ntok[u];
    const uint64_t out0 = S.blk_out0[b] + S.unit_out[u];        // absolute offset in S.out of the unit's first byte
    const uint64_t mem0 = S.mem_out_off[S.blk_member[b]];       // start of the member's output (history before it does not exist)
    uint8_t *__restrict__ g = S.out;
    const uint32_t galign = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 3u);
    uint32_t rb;                                                 // 32-bit shared address of the ring, pinned in a register (the
    {                                                            // compiler would otherwise rebuild it from %cluster_ctaid per access)
        const uint64_t ga = reinterpret_cast<uint64_t>(ring);
        asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(rb) : "l"(ga));
    }
    const uint32_t qb = rb + kResRing;                           // queue of (p1, p2) pairs of the step's in-order matches
    const uint32_t dummy = qb + 34 * 8 + lane;
    uint64_t pos = out0;
    // ring index of `pos`, kept incrementally (no modulo in the loop) and congruent to the global ADDRESS mod 4, so that aligned
    // words of the ring are aligned words of the output
    uint32_t rpos = (uint32_t)((out0 + galign) % kResRing);
    uint64_t flushed = out0;                                     // bytes before this offset are in HBM
    bool words = false;                                          // write-through by aligned words once `flushed` is word aligned
    uint32_t err = 0;
    uint32_t tnext = lane < ntok ? __ldg(tok + lane) : 0u;
    // ring index of the byte `off` bytes after the step start (off < kResRing) / `back` bytes before index i (back <= kResRing)
    #define RFWD(off) ((rpos + (off)) >= kResRing ? (rpos + (off)) - kResRing : (rpos + (off)))
    #define RBACK(i, back) ((i) >= (back) ? (i) - (back) : (i) + kResRing - (back))
    #define RWRAP(i) ((i) >= kResRing ? (i) - kResRing : (i))
    for (uint64_t i0 = 0; i0 < ntok; i0 += 32) {
        const uint32_t tk = tnext;
        const uint64_t in = i0 + 32 + lane;
        tnext = in < ntok ? __ldg(tok + in) : 0u;                // prefetch the next step's tokens
        const bool live = i0 + lane < ntok && tk != kTokSkip;
        const bool is_m = live && (tk & kSymPtr);
        const uint32_t len = !live ? 0u : is_m ? (tk >> 16) & 0x1FFu : 1u;
        uint32_t incl = len;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += v; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t off = incl - len;                          // offset of this token's output inside the step
        const uint32_t d0 = RFWD(off);                            // ring index of this token's first output byte
        if (live && !is_m) ring[d0] = (uint8_t)tk;
        const uint32_t dist = tk & 0xFFFFu;
        const uint64_t dst = pos + off;
        const bool ok = is_m && (uint64_t)dist <= dst - out0;     // source inside the unit; anything else is flagged and never dereferenced
        if (is_m && !ok) err |= (uint64_t)dist > dst - mem0 ? 2u : 1u;   // reaches before the unit (1) / before the stream (2)
        // Matches whose source is older than the ring read final data from HBM and never depend on this step's output: every
        // lane resolves its own one now, all in parallel (their loads overlap instead of queueing up one per match).
        const bool is_far = ok && dist + (total - off) > kResRing;
        if (is_far) {
            const uint8_t *gs = g + dst - dist;
            for (uint32_t k0 = 0; k0 < len; k0 += 16) {           // 16 loads in flight, then the stores (a byte-by-byte loop would
                uint32_t v[16];                                   // serialise on load latency: ring[] and g[] may alias for the compiler)
#pragma unroll
                for (uint32_t t = 0; t < 16; t++) v[t] = k0 + t < len ? (uint32_t)__ldcg(gs + k0 + t) : 0u;
#pragma unroll
                for (uint32_t t = 0; t < 16; t++)
                    if (k0 + t < len) { const uint32_t di = RWRAP(d0 + k0 + t); ring[di] = (uint8_t)v[t]; }
            }
        }
        // remaining matches in token order; every copy is spread over the 32 lanes (sorted text makes most matches depend on
        // the bytes written just before them, so resolving them independently buys nothing).  Each lane packs the parameters of
        // its own match once; the loop only shuffles them (the next match's while the current one is being copied).
        const uint32_t p1 = d0 | (len << 16);                                           // ring index of the first output byte | length
        const bool inorder = ok && !is_far;
        const uint32_t p2 = (inorder ? RBACK(d0, dist) : 0u) | (dist << 16);            // ring index of the first source byte | distance
        const uint32_t mmask = __ballot_sync(0xFFFFFFFFu, inorder);
        const uint32_t nm = __popc(mmask);
#if B2F_RESOLVE_FREE_FIRST
        // Three of four matches of titles-shaped text (tools/analysis_resolve_steps.py: 12.2 of 16.3 per step) read only bytes
        // from before the step (dist >= off + len): their sources are final, so they need no ordering among themselves and no
        // barrier between them -- the copies of successive matches overlap.  Only the others (4.2 per step) run in token order.
        const bool isfree = inorder && dist >= off + len;
        const uint32_t fmask = __ballot_sync(0xFFFFFFFFu, isfree), dmask = mmask & ~fmask;
        const uint32_t nf = __popc(fmask);
        const uint32_t lt = (1u << lane) - 1u;
        if (inorder) r_sts64(qb + 8u * (isfree ? (uint32_t)__popc(fmask & lt) : nf + (uint32_t)__popc(dmask & lt)), p1, p2);
        __syncwarp();
        for (uint32_t i = 0; i < nf; i++) {
            const uint2 c = r_lds64(qb + 8u * i);
            const uint32_t mlen = c.x >> 16;
            const uint32_t md = (c.x & 0xFFFFu) + lane, ms = (c.y & 0xFFFFu) + lane;
            for (uint32_t k0 = 0; k0 < mlen; k0 += 32) r_copy8_if(k0 + lane < mlen, rb + RWRAP(ms + k0), rb + RWRAP(md + k0));
        }
        __syncwarp();
        const uint32_t i_first = nf;
#else
        if (inorder) r_sts64(qb + 8u * __popc(mmask & ((1u << lane) - 1u)), p1, p2);   // compacted, in token order
        __syncwarp();
        const uint32_t i_first = 0;
#endif
        uint2 nx = r_lds64(qb + 8u * i_first), nx2 = r_lds64(qb + 8u * (i_first + 1u));
        for (uint32_t i = i_first; i < nm; i++) {
            const uint2 c = nx;
            nx = nx2;
            nx2 = r_lds64(qb + 8u * (i + 2));                       // parameters two matches ahead (slots >= nm are read but never used)
            const uint32_t mlen = c.x >> 16, mdist = c.y >> 16;
            const uint32_t md = (c.x & 0xFFFFu) + lane, ms = (c.y & 0xFFFFu) + lane;    // this lane's byte of the first 32-byte slice
            if (mdist >= mlen && mlen <= 32) {                      // the common case: one slice, source entirely older than the output
                r_sts8(lane < mlen ? rb + RWRAP(md) : dummy, r_lds8(rb + RWRAP(ms)));   // idle lanes store into their dummy byte: no branch
            } else if (mdist >= 32) {                               // each 32-byte slice only reads bytes of earlier slices
                for (uint32_t k0 = 0; k0 < mlen; k0 += 32) {
                    r_copy8_if(k0 + lane < mlen, rb + RWRAP(ms + k0), rb + RWRAP(md + k0));
                    __syncwarp();
                }
            } else {                                                // short period: byte k repeats byte k mod dist of the source
                const uint32_t inv = kInvPeriod[mdist];
                for (uint32_t k = lane; k < mlen; k += 32) {
                    const uint32_t r = k - mdist * ((k * inv) >> 16);
                    r_sts8(rb + RWRAP(md - lane + k), r_lds8(rb + RWRAP(ms - lane + r)));
                }
            }
            __syncwarp();
        }
        // write the step through to HBM
        const uint64_t new_end = pos + total;
        if (words) {
            const uint64_t fe = new_end - ((new_end + galign) & 3u);        // last word boundary at or below new_end
            if (fe > flushed) {
                const uint32_t nw = (uint32_t)(fe - flushed) >> 2;
                const uint32_t r0 = RBACK(rpos, (uint32_t)(pos - flushed));  // pos - flushed <= 3
                uint32_t *gw = reinterpret_cast<uint32_t *>(g + flushed);
                for (uint32_t k = lane; k < nw; k += 32) gw[k] = *reinterpret_cast<const uint32_t *>(ring + RWRAP(r0 + 4 * k));
                flushed = fe;
            }
        } else {
            for (uint32_t k = lane; k < total; k += 32) g[pos + k] = ring[RFWD(k)];
            if (new_end - out0 >= 8) { words = true; flushed = new_end - ((new_end + galign) & 3u); }
        }
        pos = new_end;
        rpos = RFWD(total);
        __syncwarp();
    }
    if (words) {                                                 // the bytes after the last whole word
        const uint32_t nt = (uint32_t)(pos - flushed), r0 = RBACK(rpos, nt);
        if (lane < nt) g[flushed + lane] = ring[RWRAP(r0 + lane)];
    }
    #undef RFWD
    #undef RBACK
    #undef RWRAP
    err = __reduce_or_sync(0xFFFFFFFFu, err);
    if (lane == 0) { S.res_err[u] = err; S.res_len[u] = pos - out0; }
}

// ---------------------------------------------------------------------------------- launchers
cudaError_t spec_init_attributes() {
    cudaError_t e = cudaFuncSetAttribute(k_spec_headers, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * sizeof(InflateTables)));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_spec_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResSmem);
}
cudaError_t spec_launch_parse(const SpecDev &S, uint32_t rounds, cudaStream_t st) {
    if (!S.n_blocks) return cudaSuccess;
    k_spec_headers<<<(S.n_blocks + 3) / 4, 128, 4 * sizeof(InflateTables), st>>>(S);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    SpecDev R = S;
    for (uint32_t r = 0; r < rounds; r++) {
        // Jacobi iteration: round r reads the exits of round r-1 (s_exit_prev) and writes s_exit
        if (r & 1) { R.s_exit = S.s_exit_prev; R.s_exit_prev = S.s_exit; } else { R.s_exit = S.s_exit; R.s_exit_prev = S.s_exit_prev; }
        k_spec_round<<<S.n_ctas, kSpecCta, 0, st>>>(R, r);
        e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    if (((rounds - 1) & 1)) { R.s_exit = S.s_exit_prev; R.s_exit_prev = S.s_exit; } else { R.s_exit = S.s_exit; R.s_exit_prev = S.s_exit_prev; }
    k_spec_verify<<<S.n_blocks, 256, 0, st>>>(R);
    return cudaGetLastError();
}
cudaError_t spec_launch_tokens(const SpecDev &S, uint32_t n_sel, cudaStream_t st) {
    if (!n_sel) return cudaSuccess;
    k_spec_tokens<<<S.n_ctas, kSpecCta, 0, st>>>(S);
    return cudaGetLastError();
}
cudaError_t spec_launch_units(const SpecDev &S, uint32_t n_sel, cudaStream_t st) {
    if (!n_sel) return cudaSuccess;
    k_spec_units<<<n_sel, 32, 0, st>>>(S);
    return cudaGetLastError();
}
// resolves the unit slots [u0, u1)
cudaError_t spec_launch_resolve(const SpecDev &S, uint32_t u0, uint32_t u1, cudaStream_t st) {
    if (u1 <= u0) return cudaSuccess;
    k_spec_resolve<<<u1 - u0, 32, kResSmem, st>>>(S, u0);
    return cudaGetLastError();
}

}  // namespace b2f

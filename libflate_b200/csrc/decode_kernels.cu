// decode_kernels.cu -- DEFLATE decode hot path (sm_100a).
//   k_find_blocks      every bit offset of a stream is tested for a plausible dynamic-block header
//                      (CTA-local two-level compaction: 17-bit precheck -> code-length-code Kraft test -> full validation)
//   (the candidates feed the sub-block parallel inflate in spec_kernels.cu)
//   k_inflate_streams  warp per stream, blocks in order (small streams, foreign streams with cross-block references,
//                      streams with stored/fixed blocks, and every error path -- it reproduces the reference's error
//                      kinds and partial output exactly)
// Every lane of a warp runs the scalar decode of inflate_core.cuh uniformly (broadcast loads, shared-memory tables);
// LZ77 copies and flushes are spread over the 32 lanes.
// Reference behaviour: src/deflate/decode.rs:81-165, src/deflate/symbol.rs:193-243,387-484,
// src/huffman.rs:96-179, libflate_lz77/src/lib.rs:164-194.
#include "common.cuh"
#include "inflate_core.cuh"
#include "finder_core.cuh"
#include "decode_dev.cuh"

namespace b2f {

struct WarpSync { __device__ __forceinline__ void operator()() const { __syncwarp(); } };

constexpr uint32_t kRingBytes = 65536, kRingMask = kRingBytes - 1;

// Output policy: the last 64 KiB of output live in a shared-memory ring (index = absolute output offset mod 64 Ki);
// back-references are served from it; every time the write position crosses a 32 KiB boundary the completed half is
// flushed to HBM with coalesced 4-byte stores.
struct WindowOut {
    uint8_t *ring; uint8_t *g; uint64_t capacity; uint64_t flushed; uint32_t lane; uint64_t last_block;
    __device__ __forceinline__ uint64_t cap() const { return capacity; }
    __device__ __forceinline__ void flush_to(uint64_t upto) {
        if (upto > capacity) upto = capacity;
        if (upto <= flushed) return;
        __syncwarp();
        const uint64_t a = flushed, b = upto;
        if (((a | b) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 3) == 0) {
            for (uint64_t i = a + 4ull * lane; i < b; i += 128) *reinterpret_cast<uint32_t *>(g + i) = *reinterpret_cast<const uint32_t *>(ring + (i & kRingMask));
        } else {
            for (uint64_t i = a + lane; i < b; i += 32) g[i] = ring[i & kRingMask];
        }
        flushed = upto;
        __syncwarp();
    }
    __device__ __forceinline__ void advance(uint64_t newpos) {
        const uint64_t boundary = newpos & ~32767ull;
        if (boundary > flushed) flush_to(boundary);
    }
    __device__ __forceinline__ void lit(uint64_t pos, uint8_t b) {
        if (lane == 0) ring[pos & kRingMask] = b;
        advance(pos + 1);
    }
    __device__ __forceinline__ void copy(uint64_t pos, uint32_t len, uint32_t dist) {
        __syncwarp();
        const uint64_t src = pos - dist;
        if (dist >= len) { for (uint32_t k = lane; k < len; k += 32) ring[(pos + k) & kRingMask] = ring[(src + k) & kRingMask]; }
        else { for (uint32_t k = lane; k < len; k += 32) ring[(pos + k) & kRingMask] = ring[(src + k % dist) & kRingMask]; }
        advance(pos + len);
    }
    __device__ __forceinline__ void raw(uint64_t pos, const uint8_t *src, uint64_t n) {   // stored block: straight to HBM (+ ring for later matches)
        flush_to(pos);
        __syncwarp();
        for (uint64_t k = lane; k < n; k += 32) g[pos + k] = src[k];
        const uint64_t keep = n > kRingBytes ? kRingBytes : n;
        for (uint64_t k = n - keep + lane; k < n; k += 32) ring[(pos + k) & kRingMask] = src[k];
        if (pos + n > flushed) flushed = pos + n;
        __syncwarp();
    }
    __device__ __forceinline__ void block_end(uint64_t, uint64_t, uint64_t out_pos, bool) { last_block = out_pos; }
    __device__ __forceinline__ int fast(BitIn &b, const InflateTables &T, uint64_t &out_pos, uint64_t hist_base);
};

// ---------------------------------------------------------------------------------- fast symbol loop (device only)
// Register-resident bit buffer with a prefetched input word, explicit 32-bit shared-memory addresses for the decode tables
// and the output ring, 32-bit loop counters.  Handles literals, EndOfBlock and matches whose codes hit the primary tables.
// Anything else (long codes, invalid patterns, end of input/output, too-long references, the probe's stop bit) leaves the
// reader untouched at the symbol boundary and returns 0, so that the exact generic path decides.
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

template <bool kCount>
__device__ __forceinline__ int fast_symbols(BitIn &b, const InflateTables &T, uint64_t &out_pos, uint64_t hist_base, uint64_t cap,
                                            uint8_t *ring, uint64_t &flushed, uint8_t *g, uint32_t lane, uint32_t &far) {
    const uint64_t nbytes = b.limit >> 3;
    if (b.eof || b.err) return 0;
    uint64_t next = b.next;
    if ((reinterpret_cast<uintptr_t>(b.p + next) & 3) != 0 || next + 16 > nbytes || out_pos + 600 > cap || b.pos > b.stop) return 0;
    const uint32_t *__restrict__ ip = reinterpret_cast<const uint32_t *>(b.p + next);
    // budgets (all 32-bit): input words, output bytes, bits before the probe's stop
    uint32_t words_left = (uint32_t)min((nbytes - next - 12) >> 2, (uint64_t)0x7FFFFFFFu);
    const uint32_t out_budget = (uint32_t)min(cap - out_pos - 258, (uint64_t)0x7FFFFFF0u);
    const uint32_t bit_budget = (uint32_t)min(b.stop - b.pos, (uint64_t)0x7FFFFFF0u);
    const uint32_t rel_hist = (uint32_t)min(out_pos + hist_base, (uint64_t)0x7FFFFFF0u);   // bytes of history before this call
    const uint32_t rel_blk = (uint32_t)min(out_pos, (uint64_t)0x7FFFFFF0u);                  // probe: bytes this block has produced so far
    uint64_t bb = b.bb; uint32_t bc = b.bc;
    uint32_t nw = __ldg(ip);
    uint32_t used_words = 0, bits = 0, produced = 0;
    const uint32_t lit_s = (uint32_t)__cvta_generic_to_shared(T.lit), dist_s = (uint32_t)__cvta_generic_to_shared(T.dist);
    const uint32_t ring_s = kCount ? 0u : (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t op0 = (uint32_t)out_pos;                        // low 32 bits of the absolute output offset (ring index = low 16 bits)
    uint32_t region = (uint32_t)flushed & ~32767u;
    int ret = 0;
    for (;;) {
        if (used_words + 2 >= words_left || produced > out_budget || bits > bit_budget) break;
        if (bc < 32) { bb |= (uint64_t)nw << bc; bc += 32; used_words++; nw = __ldg(ip + used_words); }
        const uint32_t lo = (uint32_t)bb;
        const uint32_t e = lds_u32(lit_s + ((lo & ((1u << kLitBits) - 1u)) << 2));
        const uint32_t w = e & 15u, kind = (e >> 4) & 3u;
        if (kind == kKindLit) {
            bb >>= w; bc -= w; bits += w;
            if (!kCount) { if (lane == 0) sts_u8(ring_s + ((op0 + produced) & kRingMask), e >> 8); }
            produced += 1;
        } else if (kind == kKindLen) {
            const uint32_t eb = (e >> 20) & 15u;
            const uint32_t len = ((e >> 8) & 0x1FFu) + ((lo >> w) & ((1u << eb) - 1u));
            const uint32_t used = w + eb;
            uint64_t bb2 = bb >> used; uint32_t bc2 = bc - used; uint32_t uw2 = used_words, nw2 = nw;
            if (bc2 < 32) { bb2 |= (uint64_t)nw2 << bc2; bc2 += 32; uw2++; nw2 = __ldg(ip + uw2); }
            const uint32_t lo2 = (uint32_t)bb2;
            const uint32_t d = lds_u32(dist_s + ((lo2 & ((1u << kDistBits) - 1u)) << 2));
            const uint32_t wd = d & 15u;
            if (wd == 0) break;                                    // long or unassigned distance code
            const uint32_t deb = (d >> 24) & 15u;
            const uint32_t dist = ((d >> 8) & 0xFFFFu) + ((lo2 >> wd) & ((1u << deb) - 1u));
            if (kCount) { if (dist > produced + rel_blk) far = 1; }   // reference reaches before this block's own output
            else if (dist > produced + rel_hist) break;            // "Too long backword reference": reported by the generic path
            const uint32_t used2 = wd + deb;
            bb = bb2 >> used2; bc = bc2 - used2; used_words = uw2; nw = nw2; bits += used + used2;
            if (!kCount) {
                __syncwarp();
                const uint32_t dst = op0 + produced, src = dst - dist;
                if (dist >= len) { for (uint32_t k = lane; k < len; k += 32) sts_u8(ring_s + ((dst + k) & kRingMask), lds_u8(ring_s + ((src + k) & kRingMask))); }
                else { for (uint32_t k = lane; k < len; k += 32) sts_u8(ring_s + ((dst + k) & kRingMask), lds_u8(ring_s + ((src + k % dist) & kRingMask))); }
            }
            produced += len;
        } else {
            if (kind == kKindEob) { bb >>= w; bc -= w; bits += w; ret = 1; }
            break;
        }
        if (!kCount) {
            if (((op0 + produced) & ~32767u) != region) {          // crossed a 32 KiB boundary: flush the completed half of the ring
                const uint64_t abs_now = out_pos + produced, upto = abs_now & ~32767ull;
                __syncwarp();
                const uint64_t a0 = flushed;
                if (((a0 | upto) & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 3) == 0) {
                    for (uint64_t i = a0 + 4ull * lane; i < upto; i += 128) *reinterpret_cast<uint32_t *>(g + i) = lds_u32(ring_s + ((uint32_t)i & kRingMask));
                } else {
                    for (uint64_t i = a0 + lane; i < upto; i += 32) g[i] = (uint8_t)lds_u8(ring_s + ((uint32_t)i & kRingMask));
                }
                flushed = upto; region = (uint32_t)upto & ~32767u;
                __syncwarp();
            }
        }
    }
    b.bb = bb; b.bc = bc; b.next = next + 4ull * used_words; b.pos += bits; out_pos += produced;
    return ret;
}

__device__ __forceinline__ int WindowOut::fast(BitIn &b, const InflateTables &T, uint64_t &out_pos, uint64_t hist_base) { uint32_t f = 0; return fast_symbols<false>(b, T, out_pos, hist_base, capacity, ring, flushed, g, lane, f); }

constexpr uint32_t kWinSmem = kRingBytes + (uint32_t)sizeof(InflateTables) + 64;

// ---------------------------------------------------------------------------------- in-order, warp per stream
__global__ void __launch_bounds__(32) k_inflate_streams(DecDev D) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *ring = smem_raw;
    InflateTables &T = *reinterpret_cast<InflateTables *>(smem_raw + kRingBytes);
    const uint32_t lane = threadIdx.x, s = blockIdx.x;
    if (s >= D.n) return;
    BitIn b;
    const uint8_t *p0 = D.in + D.in_off[s];
    const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3);       // start the reader on an aligned word (the lead bytes are never consumed)
    bi_init(b, p0 - lead, D.in_len[s] + lead, 8ull * lead);
    const uint64_t o0 = D.out_off[s];
    WindowOut out = { ring, D.out, o0 + D.out_cap[s], o0, lane, o0 };
    InflateResult R;
    inflate_blocks(b, T, out, o0, 0ull - o0, 0xFFFFFFFFu, (int)lane, 32, WarpSync(), R);
    out.flush_to(R.out_len);
    if (lane == 0) { D.status[s] = R.status; D.out_len[s] = R.out_len - o0; D.consumed[s] = R.consumed - lead; D.good_len[s] = out.last_block - o0; }
}

// ---------------------------------------------------------------------------------- block-boundary finder
constexpr uint32_t kFindBytes = 1024;       // input bytes (8192 bit offsets) per CTA
__device__ __forceinline__ uint64_t smem_bits64(const uint32_t *w, uint32_t bit) {
    const uint32_t byte = bit >> 3, idx = byte >> 2, sh = (byte & 3u) * 8u + (bit & 7u);
    const uint32_t x0 = w[idx], x1 = w[idx + 1], x2 = w[idx + 2];
    return (uint64_t)__funnelshift_r(x0, x1, sh) | ((uint64_t)__funnelshift_r(x1, x2, sh) << 32);
}
__device__ __forceinline__ uint32_t find_owner_u32(const uint32_t *__restrict__ prefix, uint32_t n, uint32_t idx) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= idx) lo = mid; else hi = mid; }
    return lo;
}
__global__ void __launch_bounds__(256) k_find_blocks(FindDev F, uint32_t seg_off) {
    __shared__ __align__(16) uint32_t sw[(kFindBytes + 64) / 4];
    __shared__ uint16_t qa[kFindBytes * 8], qb[kFindBytes * 8];
    __shared__ uint32_t na, nb;
    const uint32_t tid = threadIdx.x;
    const uint32_t seg = blockIdx.x + seg_off;
    const uint32_t sel = find_owner_u32(F.seg0, F.n_sel, seg);
    const uint32_t m = F.members[sel];
    const uint64_t len = F.in_len[m];
    const uint8_t *__restrict__ p = F.in + F.in_off[m];
    const uint64_t b0 = (uint64_t)(seg - F.seg0[sel]) * kFindBytes;
    if (tid == 0) { na = 0; nb = 0; }
    for (uint32_t i = tid; i < (kFindBytes + 64) / 4; i += 256) {
        uint32_t v = 0;
        const uint64_t byte = b0 + 4ull * i;
        for (uint32_t k = 0; k < 4; k++) if (byte + k < len) v |= (uint32_t)p[byte + k] << (8 * k);
        sw[i] = v;
    }
    __syncthreads();
    const uint64_t limit = len * 8;
    const uint32_t lane = tid & 31;
    {   // step A: thread t tests the 32 bit offsets of word t with one sliding 64-bit window
        const uint64_t win = smem_bits64(sw, tid * 32);
        uint32_t pass = hdr_precheck_mask32(win);
        const uint64_t q0 = b0 * 8 + tid * 32;
        if (q0 + 32 + 17 > limit) {                 // tail of the stream: drop offsets whose 17 header bits do not fit
            for (uint32_t j = 0; j < 32; j++) if (q0 + j + 17 > limit) pass &= ~(1u << j);
        }
        uint32_t cnt = __popc(pass), incl = cnt;
        for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(&na, incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31) + incl - cnt;
        while (pass) { const uint32_t j = __ffs((int)pass) - 1; pass &= pass - 1; qa[base++] = (uint16_t)(tid * 32 + j); }
    }
    __syncthreads();
    for (uint32_t i = tid; i < na; i += 256) {
        const uint32_t o = qa[i];
        if (precode_check(smem_bits64(sw, o), smem_bits64(sw, o + 64))) qb[atomicAdd(&nb, 1u)] = (uint16_t)o;
    }
    __syncthreads();
    for (uint32_t i = tid; i < nb; i += 256) {                   // survivors go to a global queue; validated densely by k_validate_candidates
        const uint32_t slot = atomicAdd(F.q_count, 1u);
        if (slot < F.q_cap) { F.q_member[slot] = m; F.q_bit[slot] = b0 * 8 + qb[i]; }
    }
}
// Validates the queue entries [*q_done, *q_count): the finder may run piece by piece while the input is still arriving.
__global__ void __launch_bounds__(128) k_validate_candidates(FindDev F) {
    const uint32_t nq = min(*F.q_count, F.q_cap);
    for (uint32_t i = *F.q_done + blockIdx.x * 128 + threadIdx.x; i < nq; i += gridDim.x * 128) {
        const uint32_t m = F.q_member[i];
        const uint64_t q = F.q_bit[i];
        if (validate_dynamic_header(F.in + F.in_off[m], F.in_len[m], q)) {
            const uint32_t slot = atomicAdd(F.cand_count, 1u);
            if (slot < F.cand_cap) { F.cand_member[slot] = m; F.cand_bit[slot] = q; }
        }
    }
}
__global__ void k_find_advance(FindDev F) { *F.q_done = min(*F.q_count, F.q_cap); }

cudaError_t dec_init_attributes() {
    return cudaFuncSetAttribute(k_inflate_streams, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWinSmem);
}
cudaError_t dec_launch_serial(const DecDev &D, cudaStream_t st) {
    if (D.n == 0) return cudaSuccess;
    k_inflate_streams<<<D.n, 32, kWinSmem, st>>>(D);
    return cudaGetLastError();
}
cudaError_t dec_launch_find(const FindDev &F, uint32_t seg_lo, uint32_t seg_hi, cudaStream_t st) {
    if (seg_hi <= seg_lo) return cudaSuccess;
    k_find_blocks<<<seg_hi - seg_lo, 256, 0, st>>>(F, seg_lo);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    // grid sized for the usual share of survivors of this range (the kernel strides over whatever was queued)
    const uint64_t want = (uint64_t)(seg_hi - seg_lo) * kFindBytes / 32 + 8192;
    k_validate_candidates<<<(unsigned)((want + 127) / 128), 128, 0, st>>>(F);
    e = cudaGetLastError(); if (e != cudaSuccess) return e;
    k_find_advance<<<1, 1, 0, st>>>(F);
    return cudaGetLastError();
}
}  // namespace b2f

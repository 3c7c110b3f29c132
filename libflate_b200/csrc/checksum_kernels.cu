// checksum_kernels.cu -- CRC-32 (ISO-HDLC, reflected 0xEDB88320) and Adler-32 as coalesced reductions.
// Replaces checksum::Crc32 / checksum::Adler32 (src/checksum.rs:4-33 -> crates crc32fast / adler32).
//
// Both checksums are linear in the message, so every thread hashes one small piece and the pieces are
// folded with the algebraic combine:
//   CRC  : crc(A||B) = crc(A) * x^(8|B|) mod P  xor crc(B)   (GF(2) polynomial product, reflected bit order)
//   Adler: A = 1 + sum d_i,  B = n + sum (n - i) d_i  (mod 65521)
#include "common.cuh"
#include "checksum_dev.cuh"

namespace b2f {

__constant__ uint32_t c_x2n[32];        // x^(2^k) mod P, reflected

static uint32_t h_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t d_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {           // bit 31 of `a` is x^0
        if ((a >> i) & 1u) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// x^(8*len) mod P
__device__ __forceinline__ uint32_t d_xpow8(uint64_t len) {
    uint32_t p = 1u << 31; uint32_t k = 3;
    while (len) {
        if (len & 1) p = d_multmodp(c_x2n[k & 31], p);
        len >>= 1; k++;
    }
    return p;
}

__device__ __forceinline__ uint32_t find_owner64(const uint64_t *__restrict__ prefix, uint32_t n, uint64_t idx) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= idx) lo = mid; else hi = mid; }
    return lo;
}

constexpr uint32_t kPiece = 512;         // bytes per thread
#ifndef B2F_CHECKSUM_VEC16
#define B2F_CHECKSUM_VEC16 0             // 1: 16-byte loads in the piece loop (not yet measured on a GPU -- round-2 candidate)
#endif

// acc_crc[s] ^= crc(piece) * x^(8 * bytes after the piece);  acc_a/acc_b: Adler partial sums (already mod 65521)
template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(256) k_checksum(ChecksumDev C) {
    __shared__ uint32_t T[4][256];
    if (DO_CRC) {
        for (uint32_t i = threadIdx.x; i < 256; i += 256) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            T[0][i] = c;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < 256; i += 256) {
            uint32_t c = T[0][i];
            for (int t = 1; t < 4; t++) { c = (c >> 8) ^ T[0][c & 0xFF]; T[t][i] = c; }
        }
        __syncthreads();
    }
    const uint64_t piece = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    const bool active = piece < C.n_pieces;
    uint32_t s = 0; uint64_t p0 = 0, p1 = 0, n = 0;
    if (active) {
        s = find_owner64(C.piece0, C.n_streams, piece);
        n = C.len[s];
        p0 = (piece - C.piece0[s]) * kPiece;
        p1 = p0 + kPiece < n ? p0 + kPiece : n;
    }
    const uint8_t *__restrict__ p = C.in + (active ? C.off[s] : 0);
    uint32_t crc = 0xFFFFFFFFu, s1 = 0, s2 = 0;
    if (active && p1 > p0) {
        uint64_t i = p0;
        const uint32_t L = (uint32_t)(p1 - p0);
        // head bytes up to 4-byte alignment of the global address
        while (i < p1 && ((reinterpret_cast<uintptr_t>(p + i)) & 3)) {
            const uint32_t d = p[i];
            if (DO_CRC) crc = T[0][(crc ^ d) & 0xFF] ^ (crc >> 8);
            if (DO_ADLER) { s1 += d; s2 += (L - (uint32_t)(i - p0)) * d; }
            i++;
        }
#if B2F_CHECKSUM_VEC16
        // Every lane walks its own 512-byte piece, so a warp-wide load touches 32 different lines whatever its width: with 16-byte
        // loads the L1 handles a quarter of the requests (the kernel is bound by them: 1.0 B of DRAM traffic per byte at 390 GB/s).
        while (i + 4 <= p1 && ((reinterpret_cast<uintptr_t>(p + i)) & 15)) {       // words up to 16-byte alignment
            const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
            if (DO_CRC) { const uint32_t a = crc ^ w; crc = T[3][a & 0xFF] ^ T[2][(a >> 8) & 0xFF] ^ T[1][(a >> 16) & 0xFF] ^ T[0][a >> 24]; }
            if (DO_ADLER) {
                const uint32_t r = L - (uint32_t)(i - p0);
                const uint32_t d0 = w & 0xFF, d1 = (w >> 8) & 0xFF, d2 = (w >> 16) & 0xFF, d3 = w >> 24;
                s1 += d0 + d1 + d2 + d3;
                s2 += r * d0 + (r - 1) * d1 + (r - 2) * d2 + (r - 3) * d3;
            }
            i += 4;
        }
        while (i + 16 <= p1) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + i);
            const uint32_t ww[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
            for (uint32_t q = 0; q < 4; q++) {
                const uint32_t w = ww[q];
                if (DO_CRC) { const uint32_t a = crc ^ w; crc = T[3][a & 0xFF] ^ T[2][(a >> 8) & 0xFF] ^ T[1][(a >> 16) & 0xFF] ^ T[0][a >> 24]; }
                if (DO_ADLER) {
                    const uint32_t r = L - (uint32_t)(i - p0) - 4u * q;
                    const uint32_t d0 = w & 0xFF, d1 = (w >> 8) & 0xFF, d2 = (w >> 16) & 0xFF, d3 = w >> 24;
                    s1 += d0 + d1 + d2 + d3;
                    s2 += r * d0 + (r - 1) * d1 + (r - 2) * d2 + (r - 3) * d3;
                }
            }
            i += 16;
        }
#endif
        while (i + 4 <= p1) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(p + i);
            if (DO_CRC) { const uint32_t a = crc ^ w; crc = T[3][a & 0xFF] ^ T[2][(a >> 8) & 0xFF] ^ T[1][(a >> 16) & 0xFF] ^ T[0][a >> 24]; }
            if (DO_ADLER) {
                const uint32_t r = L - (uint32_t)(i - p0);
                const uint32_t d0 = w & 0xFF, d1 = (w >> 8) & 0xFF, d2 = (w >> 16) & 0xFF, d3 = w >> 24;
                s1 += d0 + d1 + d2 + d3;
                s2 += r * d0 + (r - 1) * d1 + (r - 2) * d2 + (r - 3) * d3;
            }
            i += 4;
        }
        while (i < p1) {
            const uint32_t d = p[i];
            if (DO_CRC) crc = T[0][(crc ^ d) & 0xFF] ^ (crc >> 8);
            if (DO_ADLER) { s1 += d; s2 += (L - (uint32_t)(i - p0)) * d; }
            i++;
        }
    }
    if (active && p1 > p0) {
        const uint64_t after = n - p1;
        if (DO_CRC) {
            crc = ~crc;
            const uint32_t contrib = after ? d_multmodp(d_xpow8(after), crc) : crc;
            atomicXor(C.acc_crc + s, contrib);
        }
        if (DO_ADLER) {
            const uint64_t a = s1 % 65521u;
            const uint64_t b = ((uint64_t)s2 + (after % 65521u) * (uint64_t)(s1 % 65521u)) % 65521u;
            atomicAdd(reinterpret_cast<unsigned long long *>(C.acc_a + s), (unsigned long long)a);
            atomicAdd(reinterpret_cast<unsigned long long *>(C.acc_b + s), (unsigned long long)b);
        }
    }
}

// value[s] from the accumulators and the chained init value
__global__ void __launch_bounds__(64) k_checksum_final(ChecksumDev C, int do_crc, int do_adler) {
    const uint32_t s = blockIdx.x * 64 + threadIdx.x;
    if (s >= C.n_streams) return;
    const uint64_t n = C.len[s];
    if (do_crc) {
        const uint32_t init = C.init_crc ? C.init_crc[s] : 0u;
        uint32_t v = C.acc_crc[s];
        if (init) v ^= n ? d_multmodp(d_xpow8(n), init) : init;
        C.out_crc[s] = v;
    }
    if (do_adler) {
        const uint32_t init = C.init_adler ? C.init_adler[s] : 1u;
        const uint64_t ia = init & 0xFFFF, ib = init >> 16;
        const uint64_t a = (ia + C.acc_a[s]) % 65521u;
        const uint64_t b = (ib + (n % 65521u) * ia + C.acc_b[s]) % 65521u;
        C.out_adler[s] = (uint32_t)((b << 16) | a);
    }
}

cudaError_t checksum_init_tables() {
    uint32_t t[32];
    uint32_t p = 1u << 30;
    t[0] = p;
    for (int n = 1; n < 32; n++) t[n] = p = h_multmodp(p, p);
    return cudaMemcpyToSymbol(c_x2n, t, sizeof t);
}

cudaError_t checksum_launch(const ChecksumDev &C, bool do_crc, bool do_adler, cudaStream_t st) {
    if (C.n_streams == 0) return cudaSuccess;
    if (C.n_pieces) {
        const uint32_t grid = (uint32_t)((C.n_pieces + 255) / 256);
        if (do_crc && do_adler) k_checksum<true, true><<<grid, 256, 0, st>>>(C);
        else if (do_crc) k_checksum<true, false><<<grid, 256, 0, st>>>(C);
        else if (do_adler) k_checksum<false, true><<<grid, 256, 0, st>>>(C);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    k_checksum_final<<<(C.n_streams + 63) / 64, 64, 0, st>>>(C, do_crc ? 1 : 0, do_adler ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace b2f

// checksum_kernels.cu -- CRC-32 (ISO-HDLC, reflected 0xEDB88320) and Adler-32 as coalesced streaming reductions.
// Replaces checksum::Crc32 / checksum::Adler32 (src/checksum.rs:4-33 -> crates crc32fast / adler32).
//
// Both checksums are linear in the message (CRC over GF(2), Adler over Z/65521), so the message is cut into pieces that are
// hashed independently and folded with the algebraic combine.  What makes this version HBM bound is HOW it is cut:
//   * a warp reads 512-byte rows with one 16-byte load per lane (fully coalesced, eight rows in flight per warp);
//   * each lane therefore owns four words of every row, i.e. 128 interleaved sub-streams per warp.  For a sub-stream whose words
//     are 512 bytes apart the CRC recurrence is  R <- Z512(R xor w)  (Z_m = "advance the register over m zero bytes", a linear map),
//     which costs exactly the four table lookups of slice-by-4 -- with tables for Z512 instead of Z4;
//   * the four Z512 tables are replicated per lane ("bank private": entry i of lane l at word 32 i + l, 128 KiB of shared memory),
//     so the data dependent lookups never conflict;
//   * at the end of a span the 128 sub-streams are folded by a 7-level tree (Z4, Z8 inside a lane, Z16 .. Z256 across lanes by
//     shuffles; small ordinary tables), and the span's value is moved to its place in the stream with x^(8 k) mod P.
// Bytes before the stream (alignment lead-in) and after it (row padding) are read as zero: leading zeros do not change a CRC
// register that starts at 0, trailing zeros are undone by the constant x^(-8*512) in k_checksum_final (x is invertible mod P).
// Adler-32: per lane S1 = sum of chunk sums, S2 = sum of row index * chunk sum, S3 = sum of in-chunk weighted sums (dp4a);
// B follows from (bytes after the chunk) * S1 algebra, all reduced mod 65521 at the end of a span.
// The algebra was checked against zlib on the CPU first (tools/crc_interleave_model.py).
#include "common.cuh"
#include "checksum_dev.cuh"
#include "encode_dev.cuh"

namespace b2f {

constexpr uint32_t kPoly = 0xEDB88320u;
// tables (per device): Z512 slices | Z4..Z256 slices | x^(8 v 256^w) for w < 5 | x^(2^k)
__device__ uint32_t g_tabU[4 * 256];
__device__ uint32_t g_tabZ[7 * 4 * 256];
__device__ uint32_t g_tabXP[5 * 256];
__constant__ uint32_t c_x2n[32];        // x^(2^k) mod P, reflected
__constant__ uint32_t c_xinv512;        // x^(-8*512) mod P

static uint32_t h_multmodp(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ kPoly : b >> 1;
    }
    return p;
}
static uint32_t h_xpow(uint64_t bits, const uint32_t *x2n) {      // x^bits mod P
    uint32_t p = 1u << 31; uint32_t k = 0;
    while (bits) { if (bits & 1) p = h_multmodp(x2n[k & 31], p); bits >>= 1; k++; }
    return p;
}
__device__ __forceinline__ uint32_t d_multmodp(uint32_t a, uint32_t b) {
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 31; i >= 0; i--) {           // bit 31 of `a` is x^0
        if ((a >> i) & 1u) p ^= b;
        b = (b & 1u) ? (b >> 1) ^ kPoly : b >> 1;
    }
    return p;
}
// x^(8*len) mod P, generic exponent (k_checksum_final: once per stream)
__device__ __forceinline__ uint32_t d_xpow8(uint64_t len) {
    uint32_t p = 1u << 31; uint32_t k = 3;
    while (len) {
        if (len & 1) p = d_multmodp(c_x2n[k & 31], p);
        len >>= 1; k++;
    }
    return p;
}
// v * x^(8*e) mod P with the byte-window table (e < 2^40): at most five products
__device__ __forceinline__ uint32_t d_shift8(uint32_t v, uint64_t e) {
#pragma unroll 1
    for (uint32_t w = 0; w < 5 && e; w++, e >>= 8) {
        const uint32_t b = (uint32_t)e & 255u;
        if (b) v = d_multmodp(__ldg(&g_tabXP[w * 256 + b]), v);
    }
    return v;
}

__device__ __forceinline__ uint32_t find_owner64(const uint64_t *__restrict__ prefix, uint32_t n, uint64_t idx) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (prefix[mid] <= idx) lo = mid; else hi = mid; }
    return lo;
}

constexpr uint32_t kCkThreads = 1024;
constexpr uint32_t kCkUBytes = 4 * 256 * 32 * 4;                  // bank-private Z512 tables
constexpr uint32_t kCkZBytes = 7 * 4 * 256 * 4;
constexpr uint32_t kCkSmemCrc = kCkUBytes + kCkZBytes + 32768;    // + slack to place the Z512 tables on a 32 KiB boundary

__device__ __forceinline__ uint32_t ck_lds(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
// Z512(x): four conflict-free lookups; ul = shared address of the lane's column of table 0.  Table 0 starts on a 32 KiB boundary,
// so bits 7..14 of ul are zero and (byte << 7) | ul is one LOP3; the table number goes into the load's immediate offset.
template <uint32_t kOff> __device__ __forceinline__ uint32_t ck_lds_off(uint32_t a) { uint32_t v; asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(kOff)); return v; }
__device__ __forceinline__ uint32_t ck_adv512(uint32_t x, uint32_t ul) {
    const uint32_t a = ck_lds_off<0>(((x << 7) & 0x7F80u) | ul);
    const uint32_t b = ck_lds_off<32768>(((x >> 1) & 0x7F80u) | ul);
    const uint32_t c = ck_lds_off<65536>(((x >> 9) & 0x7F80u) | ul);
    const uint32_t d = ck_lds_off<98304>(((x >> 17) & 0x7F80u) | ul);
    return a ^ b ^ c ^ d;
}
// Z_(4 << level)(x) from the ordinary tables (zt = shared address of g_tabZ's copy)
__device__ __forceinline__ uint32_t ck_advz(uint32_t x, uint32_t zt, uint32_t level) {
    const uint32_t t = zt + level * 4096u;
    return ck_lds(t + ((x & 0xFFu) << 2)) ^ ck_lds(t + 1024u + (((x >> 8) & 0xFFu) << 2)) ^ ck_lds(t + 2048u + (((x >> 16) & 0xFFu) << 2)) ^
           ck_lds(t + 3072u + ((x >> 24) << 2));
}

// the lane's 16-byte chunk at virtual position pv of the stream (virtual = counted from the stream start rounded down to 16);
// bytes outside [mis, nv) read as zero and nothing outside the stream's own 16-byte chunks is touched
__device__ __forceinline__ uint4 ck_load_edge(const uint8_t *__restrict__ base, uint64_t pv, uint64_t mis, uint64_t nv) {
    uint32_t w[4] = { 0, 0, 0, 0 };
    for (uint32_t k = 0; k < 16; k++) {
        const uint64_t q = pv + k;
        if (q >= mis && q < nv) w[k >> 2] |= (uint32_t)base[q] << (8u * (k & 3u));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// acc_crc[s] ^= raw(span) * x^(8 * (bytes of the stream after the span's rows + 512));  acc_a / acc_b: Adler partial sums mod 65521
template <bool DO_CRC, bool DO_ADLER>
__global__ void __launch_bounds__(kCkThreads, 1) k_checksum(ChecksumDev C) {
    extern __shared__ __align__(16) uint8_t csm[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t ul = 0, zt = 0;
    if (DO_CRC) {
        const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(csm);
        const uint32_t pad = (32768u - (sbase & 32767u)) & 32767u;             // Z512 tables on a 32 KiB boundary of the shared window
        uint32_t *su = reinterpret_cast<uint32_t *>(csm + pad);
        // replicate: the warp takes 32 entries at a time, entry e goes to words 32 e + lane (conflict free)
        for (uint32_t e0 = warp * 32; e0 < 1024; e0 += 32 * (kCkThreads / 32)) {
            const uint32_t mine = g_tabU[e0 + lane];
#pragma unroll 8
            for (uint32_t e = 0; e < 32; e++) su[(e0 + e) * 32 + lane] = __shfl_sync(0xFFFFFFFFu, mine, e);
        }
        uint32_t *sz = reinterpret_cast<uint32_t *>(csm + pad + kCkUBytes);
        for (uint32_t i = threadIdx.x; i < 7 * 4 * 256; i += kCkThreads) sz[i] = g_tabZ[i];
        ul = (uint32_t)__cvta_generic_to_shared(su) + 4u * lane;
        zt = (uint32_t)__cvta_generic_to_shared(sz);
        __syncthreads();
    }
    for (uint64_t item = (uint64_t)blockIdx.x * (kCkThreads / 32) + warp; item < C.n_spans; item += (uint64_t)gridDim.x * (kCkThreads / 32)) {
        const uint32_t s = find_owner64(C.span0, C.n_streams, item);
        const uint8_t *p0 = C.in + C.off[s];
        const uint64_t mis = reinterpret_cast<uintptr_t>(p0) & 15u;
        const uint8_t *__restrict__ base = p0 - mis;                          // 16-byte aligned; virtual position 0
        const uint64_t n = C.len[s], nv = n + mis;
        const uint64_t rows_total = (nv + kChecksumRow - 1) / kChecksumRow;
        const uint64_t r0 = (item - C.span0[s]) * C.span_rows;
        const uint64_t r1 = min(rows_total, r0 + C.span_rows);
        const uint32_t K = (uint32_t)(r1 - r0);
        uint32_t R0 = 0, R1 = 0, R2 = 0, R3 = 0;                              // CRC registers of the lane's four sub-streams
        uint32_t S1 = 0, S2 = 0, S3 = 0;
        // one 512-byte row: k = row index inside the span; the span's LAST row is not followed by the Z512 advance
        auto adler_row = [&](const uint4 v, uint32_t k) {
            const uint32_t c0 = __dp4a(v.x, 0x01010101u, 0u), c1 = __dp4a(v.y, 0x01010101u, 0u), c2 = __dp4a(v.z, 0x01010101u, 0u), c3 = __dp4a(v.w, 0x01010101u, 0u);
            const uint32_t c = c0 + c1 + c2 + c3;
            S1 += c; S2 += k * c;
            S3 += __dp4a(v.x, 0x03020100u, 0u) + __dp4a(v.y, 0x03020100u, 0u) + __dp4a(v.z, 0x03020100u, 0u) + __dp4a(v.w, 0x03020100u, 0u) + 4u * c1 + 8u * c2 + 12u * c3;
        };
        auto row_mid = [&](const uint4 v, uint32_t k) {
            if (DO_ADLER) adler_row(v, k);
            if (DO_CRC) { R0 = ck_adv512(R0 ^ v.x, ul); R1 = ck_adv512(R1 ^ v.y, ul); R2 = ck_adv512(R2 ^ v.z, ul); R3 = ck_adv512(R3 ^ v.w, ul); }
        };
        auto row_last = [&](const uint4 v, uint32_t k) {
            if (DO_ADLER) adler_row(v, k);
            if (DO_CRC) { R0 ^= v.x; R1 ^= v.y; R2 ^= v.z; R3 ^= v.w; }
        };
        // Rows that lie wholly inside the stream are loaded unconditionally, four at a time with the next four already in flight.
        // The span's last row and the (at most two) rows that contain the stream's ends go through the guarded loader.
        const uint32_t ri0 = mis ? 1u : 0u;                                   // first interior row of the stream
        const uint64_t ri1 = nv / kChecksumRow;                               // interior rows: [ri0, ri1)
        uint32_t k = 0;                                                       // row r0 + k is next
        const uint32_t kl = K - 1;                                            // the span's last row
        if (r0 < ri0 && k < kl) { row_mid(ck_load_edge(base, r0 * kChecksumRow + 16u * lane, mis, nv), 0); k = 1; }
        const uint32_t kmid = ri1 > r0 ? (uint32_t)min((uint64_t)kl, ri1 - r0) : 0u;   // rows [k, kmid) are interior and not last
        if (k < kmid) {
            const uint4 *__restrict__ gp = reinterpret_cast<const uint4 *>(base + (r0 + k) * kChecksumRow) + lane;   // row stride = 32 uint4
            const uint32_t cnt = kmid - k, cnt4 = cnt & ~3u;
            uint4 cur[4], nxt[4];
            if (cnt4) {
#pragma unroll
                for (uint32_t u = 0; u < 4; u++) cur[u] = __ldg(gp + 32u * u);
                for (uint32_t done = 0; done < cnt4; done += 4) {
                    if (done + 4 < cnt4) {
#pragma unroll
                        for (uint32_t u = 0; u < 4; u++) nxt[u] = __ldg(gp + 32u * (done + 4 + u));
                    }
#pragma unroll
                    for (uint32_t u = 0; u < 4; u++) row_mid(cur[u], k + done + u);
#pragma unroll
                    for (uint32_t u = 0; u < 4; u++) cur[u] = nxt[u];
                }
            }
            for (uint32_t d = cnt4; d < cnt; d++) row_mid(__ldg(gp + 32u * d), k + d);
            k = kmid;
        }
        for (; k < kl; k++) row_mid(ck_load_edge(base, (r0 + k) * kChecksumRow + 16u * lane, mis, nv), k);     // (a padded row before the last: never in practice)
        {
            const uint64_t r = r0 + kl;
            const bool interior = r >= ri0 && r < ri1;
            row_last(interior ? __ldg(reinterpret_cast<const uint4 *>(base + r * kChecksumRow) + lane) : ck_load_edge(base, r * kChecksumRow + 16u * lane, mis, nv), kl);
        }
        const uint64_t vend = r1 * kChecksumRow;                              // virtual end of the span's rows (may lie beyond nv in the last row)
        if (DO_CRC) {
            // fold the 128 sub-streams: slot q of lane l sits 4 (4 l + q) bytes into the row
            uint32_t y = ck_advz(ck_advz(R0, zt, 0) ^ R1, zt, 1) ^ (ck_advz(R2, zt, 0) ^ R3);
#pragma unroll
            for (uint32_t lv = 0; lv < 5; lv++) {
                const uint32_t other = __shfl_down_sync(0xFFFFFFFFu, y, 1u << lv);
                if ((lane & ((2u << lv) - 1u)) == 0) y = ck_advz(y, zt, 2 + lv) ^ other;
            }
            // lane 0: raw(rows) = Z4(y); place it: x^(8 (nv - vend + 4 + 512)); the + 512 keeps the exponent positive (undone in the final kernel)
            if (lane == 0) atomicXor(C.acc_crc + s, d_shift8(y, nv + 4u + kChecksumRow - vend));
        }
        if (DO_ADLER) {
            // weight of the byte j of the chunk of row k: nv - (r0 512 + 512 k + 16 lane + j)
            const uint64_t w0 = nv - r0 * kChecksumRow - 16u * lane;          // (wraps for lanes beyond the end: their S1 is 0)
            const uint64_t t1 = (w0 % 65521u) * (uint64_t)(S1 % 65521u) % 65521u;
            const uint64_t t2 = ((uint64_t)kChecksumRow * S2 + S3) % 65521u;
            uint32_t a = S1 % 65521u, b = (uint32_t)((t1 + 65521u - t2) % 65521u);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, d); b += __shfl_xor_sync(0xFFFFFFFFu, b, d); }
            if (lane == 0) {
                atomicAdd(reinterpret_cast<unsigned long long *>(C.acc_a + s), (unsigned long long)(a % 65521u));
                atomicAdd(reinterpret_cast<unsigned long long *>(C.acc_b + s), (unsigned long long)(b % 65521u));
            }
        }
    }
}

// value[s] from the accumulators and the chained init value
__global__ void __launch_bounds__(64) k_checksum_final(ChecksumDev C, int do_crc, int do_adler) {
    const uint32_t s = blockIdx.x * 64 + threadIdx.x;
    if (s >= C.n_streams) return;
    const uint64_t n = C.len[s];
    if (do_crc) {
        // crc(M, init) = Z_n(init ^ ~0) ^ raw(M) ^ ~0 ; raw(M) = acc * x^(-8*512)
        const uint32_t init = C.init_crc ? C.init_crc[s] : 0u;
        const uint32_t raw = d_multmodp(c_xinv512, C.acc_crc[s]);
        C.out_crc[s] = n ? (d_multmodp(d_xpow8(n), ~init) ^ raw ^ 0xFFFFFFFFu) : init;
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
    static uint32_t hU[4 * 256], hZ[7 * 4 * 256], hXP[5 * 256], x2n[32], t0[256];
    static bool built = false;
    static uint32_t xinv = 0;
    if (!built) {
        uint32_t p = 1u << 30;
        x2n[0] = p;
        for (int n = 1; n < 32; n++) x2n[n] = p = h_multmodp(p, p);
        for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? kPoly ^ (c >> 1) : c >> 1; t0[i] = c; }
        auto zadv = [&](uint32_t c, uint32_t m) { for (uint32_t k = 0; k < m; k++) c = t0[c & 0xFF] ^ (c >> 8); return c; };
        for (uint32_t j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) hU[j * 256 + b] = zadv(b << (8 * j), 512);
        for (uint32_t lv = 0; lv < 7; lv++) for (uint32_t j = 0; j < 4; j++) for (uint32_t b = 0; b < 256; b++) hZ[(lv * 4 + j) * 256 + b] = zadv(b << (8 * j), 4u << lv);
        for (uint32_t w = 0; w < 5; w++) for (uint32_t v = 0; v < 256; v++) hXP[w * 256 + v] = h_xpow(8ull * v << (8 * w), x2n);
        xinv = h_xpow(0xFFFFFFFFull - 8ull * kChecksumRow, x2n);              // x has order 2^32 - 1 modulo the (primitive) CRC-32 polynomial
        built = true;
    }
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(g_tabU, hU, sizeof hU)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(g_tabZ, hZ, sizeof hZ)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(g_tabXP, hXP, sizeof hXP)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_x2n, x2n, sizeof x2n)) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_xinv512, &xinv, sizeof xinv)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(k_checksum<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCkSmemCrc)) != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_checksum<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCkSmemCrc);
}

cudaError_t checksum_launch(const ChecksumDev &C, bool do_crc, bool do_adler, cudaStream_t st, StageTimer *tm) {
    if (C.n_streams == 0) return cudaSuccess;
    if (C.n_spans) {
        const uint32_t per = kCkThreads / 32;
        const uint32_t grid = (uint32_t)((C.n_spans + per - 1) / per < 148 ? (C.n_spans + per - 1) / per : 148);
        if (do_crc && do_adler) k_checksum<true, true><<<grid, kCkThreads, kCkSmemCrc, st>>>(C);
        else if (do_crc) k_checksum<true, false><<<grid, kCkThreads, kCkSmemCrc, st>>>(C);
        else if (do_adler) k_checksum<false, true><<<grid, kCkThreads, 0, st>>>(C);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return e;
    }
    if (tm) tm->mark(st, "checksum_final");
    k_checksum_final<<<(C.n_streams + 63) / 64, 64, 0, st>>>(C, do_crc ? 1 : 0, do_adler ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace b2f

// common.cuh -- shared types/constants for the B200 DEFLATE path (libb2f.so).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define B2F_HD __host__ __device__ __forceinline__
#define B2F_D __device__ __forceinline__
#else
#define B2F_HD inline
#define B2F_D inline
#endif

namespace b2f {

// ---- geometry of the encode pipeline (see DESIGN.md "Data layout in HBM") -----------------
constexpr uint32_t kTile = 1024;        // greedy-parse tile (positions); must be >= 259
constexpr uint32_t kExitW = 258;        // entry points per tile that a previous tile can exit into
constexpr uint32_t kPTile = 16384;      // match-kernel tile (positions staged in shared memory)
constexpr uint32_t kSeg = 262144;       // chain-build segment (segments after the first of a chunk re-insert 32 KiB of warm-up)
constexpr uint32_t kHashBits = 14;      // chain-build hash table = 2^14 u32 = 64 KiB per warp
constexpr uint32_t kLookback = 32768;   // libflate_lz77::MAX_DISTANCE
constexpr uint32_t kGrpTiles = 64;      // tiles per emit CTA (one chunk per CTA)
constexpr uint32_t kHdrWords = 160;     // dynamic header bit buffer per block (<= 4495 bits)
constexpr uint32_t kHistStride = 320;   // 286 lit/len + 30 dist (+pad)
constexpr uint32_t kLitStride = 288;
constexpr uint32_t kDistStride = 32;

// symbol word layout == C ABI code word (include/b2f.h): literal byte | 0x80000000|len<<16|dist
constexpr uint32_t kSymPtr = 0x80000000u;

struct ChunkDesc {      // one LZ77 chunk (DefaultLz77Encoder::flush unit, libflate_lz77/src/default.rs:69-109)
    uint64_t off;       // offset of the chunk's first byte in the concatenated device input
    uint32_t len;
    uint32_t block;     // DEFLATE block this chunk is flushed into
};
struct BlockDesc {      // one DEFLATE block (Block::flush, src/deflate/encode.rs:287-295)
    uint32_t stream;
    uint32_t chunk0, nchunks;
    uint32_t tile0, ntiles;
    uint8_t is_final;
    uint8_t sync_after;  // zlib_sync_flush marker after this block (encode.rs:225-234)
    uint8_t fixed;       // FixedHuffmanCodec
    uint8_t pad;
};

// ---- DEFLATE symbol arithmetic (RFC 1951 tables == src/deflate/symbol.rs:22-87, 95-154) ------
B2F_HD int ilog2_u32(uint32_t x) {      // floor(log2(x)), x > 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
// length 3..258 -> (litlen code 257..285, extra bit count, extra value)
B2F_HD void length_code(uint32_t len, uint32_t &code, uint32_t &ebits, uint32_t &extra) {
    uint32_t l = len - 3;
    if (l < 8) { code = 257 + l; ebits = 0; extra = 0; }
    else if (l == 255) { code = 285; ebits = 0; extra = 0; }
    else { int hb = ilog2_u32(l); ebits = (uint32_t)hb - 2; code = 257 + 4 * ((uint32_t)hb - 1) + ((l >> ebits) & 3); extra = l & ((1u << ebits) - 1); }
}
// distance 1..32768 -> (code 0..29, extra bit count, extra value)
B2F_HD void dist_code(uint32_t dist, uint32_t &code, uint32_t &ebits, uint32_t &extra) {
    uint32_t d = dist - 1;
    if (d < 4) { code = d; ebits = 0; extra = 0; }
    else { int hb = ilog2_u32(d); ebits = (uint32_t)hb - 1; code = 2 * (uint32_t)hb + ((d >> ebits) & 1); extra = d & ((1u << ebits) - 1); }
}
B2F_HD uint32_t bitrev(uint32_t v, uint32_t width) {   // reverse the low `width` bits (huffman.rs:19-28)
#if defined(__CUDA_ARCH__)
    return width ? (__brev(v) >> (32 - width)) : 0;
#else
    uint32_t t = 0; for (uint32_t k = 0; k < width; k++) { t = (t << 1) | (v & 1); v >>= 1; } return t;
#endif
}

}  // namespace b2f

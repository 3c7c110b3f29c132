#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
namespace b2f {
// Work decomposition of the checksum kernel (checksum_kernels.cu): a stream is read as 512-byte ROWS of 16-byte aligned chunks
// (row 0 starts at the stream's first byte rounded down to 16; bytes outside the stream count as zero), and a SPAN is
// `span_rows` consecutive rows of one stream -- the unit of work of one warp.
struct ChecksumDev {
    const uint8_t *in;           // device base pointer
    const uint64_t *off;         // [n_streams] byte offset of each buffer
    const uint64_t *len;         // [n_streams]
    const uint64_t *span0;       // [n_streams + 1] prefix of spans
    uint64_t n_spans;
    uint32_t n_streams;
    uint32_t span_rows;          // rows per span (<= kChecksumMaxRows)
    uint32_t *acc_crc;           // [n_streams] zeroed
    uint64_t *acc_a, *acc_b;     // [n_streams] zeroed
    const uint32_t *init_crc, *init_adler;   // optional chaining values
    uint32_t *out_crc, *out_adler;
};
constexpr uint32_t kChecksumRow = 512;
constexpr uint32_t kChecksumMaxRows = 256;     // Adler partial sums stay below 2^32 (checksum_kernels.cu)
constexpr uint32_t kChecksumWarps = 148 * 32;  // resident warps of the persistent grid

inline uint64_t checksum_rows(const uint8_t *base, uint64_t off, uint64_t len) {
    const uint64_t mis = (reinterpret_cast<uintptr_t>(base) + off) & 15u;
    return len ? (mis + len + kChecksumRow - 1) / kChecksumRow : 0;
}
// Fills span0[0..n] and picks span_rows so that a large batch is about one span per resident warp (all warps finish together).
inline void checksum_plan(const uint8_t *base, const uint64_t *off, const uint64_t *len, size_t n, uint64_t *span0, uint32_t *span_rows) {
    uint64_t rows = 0;
    for (size_t s = 0; s < n; s++) rows += checksum_rows(base, off[s], len[s]);
    uint64_t sr = (rows + kChecksumWarps - 1) / kChecksumWarps;
    if (sr < 16) sr = 16;
    if (sr > kChecksumMaxRows) sr = kChecksumMaxRows;
    *span_rows = (uint32_t)sr;
    span0[0] = 0;
    for (size_t s = 0; s < n; s++) span0[s + 1] = span0[s] + (checksum_rows(base, off[s], len[s]) + sr - 1) / sr;
}
cudaError_t checksum_init_tables();
struct StageTimer;
// tm (optional): a "checksum_final" mark is placed between the streaming kernel and the per-stream finalisation, so that the
// stage called "checksum" is the HBM-bound kernel alone
cudaError_t checksum_launch(const ChecksumDev &C, bool do_crc, bool do_adler, cudaStream_t st, StageTimer *tm = nullptr);
}

#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace b2f {
struct ChecksumDev {
    const uint8_t *in;           // device base pointer
    const uint64_t *off;         // [n_streams] byte offset of each buffer
    const uint64_t *len;         // [n_streams]
    const uint64_t *piece0;      // [n_streams + 1] prefix of 512-byte pieces
    uint64_t n_pieces;
    uint32_t n_streams;
    uint32_t *acc_crc;           // [n_streams] zeroed
    uint64_t *acc_a, *acc_b;     // [n_streams] zeroed
    const uint32_t *init_crc, *init_adler;   // optional chaining values
    uint32_t *out_crc, *out_adler;
};
constexpr uint32_t kChecksumPiece = 512;
cudaError_t checksum_init_tables();
cudaError_t checksum_launch(const ChecksumDev &C, bool do_crc, bool do_adler, cudaStream_t st);
}

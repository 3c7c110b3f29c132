#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace b2f {
struct DecDev {
    const uint8_t *in;           // device base of the compressed bytes (+ >= 64 B padding)
    const uint64_t *in_off;      // [n] offset of each raw DEFLATE stream
    const uint64_t *in_len;      // [n] bytes available to the stream (up to the end of its container)
    uint8_t *out;                // device output base
    const uint64_t *out_off;     // [n]
    const uint64_t *out_cap;     // [n]
    uint32_t n;
    // results
    int32_t *status;             // [n] kInf*
    uint64_t *out_len;           // [n]
    uint64_t *consumed;          // [n] bytes pulled from the reader
};
cudaError_t dec_init_attributes();
cudaError_t dec_launch_serial(const DecDev &D, cudaStream_t st);
}

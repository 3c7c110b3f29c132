#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace b2f {
// one raw DEFLATE stream ("member") per entry
struct DecDev {
    const uint8_t *in;           // device base of the compressed bytes (+ >= 64 B padding)
    const uint64_t *in_off;      // [n] offset of each raw DEFLATE stream
    const uint64_t *in_len;      // [n] bytes available to the stream (up to the end of its container)
    uint8_t *out;                // device output base (256-aligned)
    const uint64_t *out_off;     // [n]
    const uint64_t *out_cap;     // [n]
    uint32_t n;
    // results of the in-order kernel
    int32_t *status;             // [n] kInf*
    uint64_t *out_len;           // [n]
    uint64_t *consumed;          // [n] bytes pulled from the reader
    uint64_t *good_len;          // [n] output bytes of the blocks that completed (Decoder::read hands those out before an error surfaces)
};
// block-boundary finder over a subset of members
struct FindDev {
    const uint8_t *in;
    const uint64_t *in_off, *in_len;     // per member
    const uint32_t *members;             // [n_sel] member indices to scan
    const uint32_t *seg0;                // [n_sel + 1] prefix of 1 KiB scan segments
    uint32_t n_sel, n_segs;
    uint32_t *q_member; uint64_t *q_bit; uint32_t *q_count; uint32_t q_cap;            // offsets that passed the cheap tests
    uint32_t *q_done;                                                                  // queue entries already validated
    uint32_t *cand_member; uint64_t *cand_bit; uint32_t *cand_count; uint32_t cand_cap;   // fully validated candidates
};
cudaError_t dec_init_attributes();
cudaError_t dec_launch_serial(const DecDev &D, cudaStream_t st);
cudaError_t dec_launch_find(const FindDev &F, uint32_t seg_lo, uint32_t seg_hi, cudaStream_t st);   // segments [seg_lo, seg_hi) of F.seg0
}

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
};
// block-boundary finder over a subset of members
struct FindDev {
    const uint8_t *in;
    const uint64_t *in_off, *in_len;     // per member
    const uint32_t *members;             // [n_sel] member indices to scan
    const uint32_t *seg0;                // [n_sel + 1] prefix of 1 KiB scan segments
    uint32_t n_sel, n_segs;
    uint32_t *q_member; uint64_t *q_bit; uint32_t *q_count; uint32_t q_cap;            // offsets that passed the cheap tests
    uint32_t *cand_member; uint64_t *cand_bit; uint32_t *cand_count; uint32_t cand_cap;   // fully validated candidates
};
// candidate probe (pass 1) and block decode (pass 2)
struct BlockDev {
    const uint8_t *in;
    const uint64_t *in_off, *in_len;     // per member
    uint32_t n_blocks;
    const uint32_t *blk_member;          // [n_blocks]
    const uint64_t *blk_bit;             // [n_blocks] start bit inside the member
    const uint64_t *blk_stop;            // [n_blocks] probe stop bit (pass 1)
    // pass 1 results
    int32_t *p_status; uint64_t *p_end_bit; uint64_t *p_out_len; uint32_t *p_flags;   // flags: 1 final, 2 needs earlier history
    // pass 2 inputs/results
    uint8_t *out;
    const uint64_t *blk_out;             // [n_blocks] absolute output offset of the block
    const uint64_t *mem_out_off;         // per member: start of its output
    const uint64_t *mem_out_end;         // per member: out_off + cap
    int32_t *d_status; uint64_t *d_out_len;
};
cudaError_t dec_init_attributes();
cudaError_t dec_launch_serial(const DecDev &D, cudaStream_t st);
cudaError_t dec_launch_find(const FindDev &F, cudaStream_t st);
cudaError_t dec_launch_probe(const BlockDev &B, cudaStream_t st);
cudaError_t dec_launch_blocks(const BlockDev &B, cudaStream_t st);
}

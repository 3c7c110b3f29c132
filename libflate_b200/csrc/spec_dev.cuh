#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "inflate_core.cuh"
namespace b2f {
constexpr uint32_t kSpecBits = 4096;      // bits per speculative subsegment (one thread each)
constexpr uint32_t kSpecCta = 128;        // subsegments per CTA (all of one block)
constexpr uint64_t kUnitMinBytes = 65536; // smallest independent LZ77 unit worth its own warp
struct SpecDev {
    const uint8_t *in; const uint64_t *in_off, *in_len;             // members
    uint32_t n_blocks;                                              // candidate blocks, sorted by (member, bit)
    const uint32_t *blk_member; const uint64_t *blk_bit, *blk_end;  // end = next candidate of the member, or the member's end
    const uint32_t *blk_seg0, *blk_cta0;                            // prefixes (n_blocks + 1)
    uint32_t n_segs, n_ctas;
    InflateTables *tabs;                                            // [n_blocks] decode tables in HBM
    uint32_t *blk_data_rel, *blk_flags;                             // first symbol bit (relative), bit0 = BFINAL
    uint32_t *s_start, *s_exit, *s_exit_prev, *s_eob_end, *s_nsym, *s_nbytes;   // per subsegment
    uint32_t *changed;
    uint64_t *s_out_rel, *s_tok_rel;                                // per subsegment: offsets inside the block
    int64_t *s_min_src;                                             // per subsegment: lowest block-relative position its matches read
    uint32_t *blk_status, *blk_eob_end, *blk_eob_seg; uint64_t *blk_nout, *blk_ntok;   // verify results
    // second phase (after the host has walked the chain)
    const uint32_t *blk_sel; const uint64_t *blk_out0, *blk_tok0; const uint32_t *sel_blocks;
    const uint64_t *mem_out_off;
    uint32_t *tokens; uint8_t *out;
    const uint32_t *sel_unit0;                                      // [n_sel + 1] prefix of unit slots per selected block
    uint64_t *unit_out, *unit_tok, *unit_ntok, *unit_nout; uint32_t *unit_blk;   // per unit slot
    uint32_t *res_err; uint64_t *res_len;                           // per unit slot
};
cudaError_t spec_init_attributes();
cudaError_t spec_launch_parse(const SpecDev &S, uint32_t rounds, cudaStream_t st);
cudaError_t spec_launch_tokens(const SpecDev &S, uint32_t n_sel, cudaStream_t st);
cudaError_t spec_launch_units(const SpecDev &S, uint32_t n_sel, cudaStream_t st);
cudaError_t spec_launch_resolve(const SpecDev &S, uint32_t u0, uint32_t u1, cudaStream_t st);
}

#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "inflate_core.cuh"
namespace b2f {
#ifndef B2F_SPEC_BITS
#define B2F_SPEC_BITS 2048
#endif
#ifndef B2F_SPEC_CTA
#define B2F_SPEC_CTA 128
#endif
constexpr uint32_t kSpecBits = B2F_SPEC_BITS;      // bits per speculative subsegment (one thread each)
constexpr uint32_t kSpecCta = B2F_SPEC_CTA;        // subsegments per CTA (all of one block)
// LZ77 resolution works on SEGMENTS: runs of subsegments of one block whose output starts inside the same kSegBytes-aligned
// window of the block's output.  One warp resolves a segment on its own; bytes copied from before the segment's first byte
// become 16-bit MARKERS (0x8000 | distance before the segment start - 1) that a second pass substitutes in stream order.
#ifndef B2F_SEG_BYTES
#define B2F_SEG_BYTES 4096
#endif
constexpr uint32_t kSegBytes = B2F_SEG_BYTES;   // output bytes per segment slot (a segment can be longer: its last subsegment is never split)
#ifndef B2F_SEG_RING
#define B2F_SEG_RING 4096
#endif
#ifndef B2F_SEG_STEPMAX
#define B2F_SEG_STEPMAX (B2F_SEG_RING / 2)
#endif
#ifndef B2F_SEG_LAZY
#define B2F_SEG_LAZY (B2F_SEG_RING / 4)
#endif
constexpr uint32_t kSegRing = B2F_SEG_RING;          // symbols of a segment kept in shared memory by k_seg_resolve
constexpr uint32_t kSegStepMax = B2F_SEG_STEPMAX;    // output symbols of one 32-token step (a step with more output is cut short)
constexpr uint32_t kSegLazy = B2F_SEG_LAZY;          // symbols that may wait in the ring before they are written to sym16
constexpr uint32_t kMarker = 0x8000u;
constexpr uint32_t kChainCounters = 8;
constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
// A chain longer than this is cut into SOFT chains at multiples of it (of the offset in out): a soft chain is resolved with its
// history unknown (markers relative to its own start survive), its last 32 KiB are then fixed up chain after chain and the rest in
// parallel -- a foreign stream with no natural cut is otherwise ONE serial chain.  libflate's own blocks (1 MiB) are never cut.
constexpr uint64_t kSuperBytes = 2u << 20;
constexpr uint32_t kMaxParts = 8;         // pipeline depth of the LZ77 resolution (part p's output is copied out while part p+1 is resolved)
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
    uint32_t n_sel, n_slots;
    const uint32_t *sel_slot0;                                      // [n_sel + 1] prefix of segment slots per selected block (stream order)
    uint64_t *seg_tok, *seg_out;                                    // per slot: first token (absolute index into tokens), first output byte (offset in out)
    uint32_t *seg_ntok, *seg_nout;                                  // per slot: tokens / output bytes (0 = empty slot)
    uint32_t *seg_member, *seg_reach;                               // per slot: member; how far before its first byte its matches reach (0 = self-contained)
    uint8_t *seg_cut;                                               // per slot: 1 = no segment from here on reads anything before this one (chain start)
    uint4 *seg_rec;                                                 // [n_slots + 2] packed copy for k_seg_subst: out offset lo/hi, bytes, cut | member << 1
    uint32_t *chain_list;                                           // [n_slots] chain starts, grouped by part (part p from part_slot0[p])
    uint32_t *chain_count;                                          // [kChainCounters * kMaxParts] zeroed; per part: hard chains | next hard chain | soft chains |
                                                                    // next soft chain | hard chains with soft successors | next of those | rest items taken
    uint32_t *soft_list, *tail_list;                                // [n_slots] each, grouped by part like chain_list: soft chain starts; hard starts followed by soft ones
    uint64_t *chain_end;                                            // [n_slots] per chain start: out offset where its chain ends
    uint32_t *chain_next_soft;                                      // [n_slots] per chain start: slot of the soft chain that continues it (kNoSlot: none)
    uint32_t n_parts, part_slot0[kMaxParts + 1];                    // parts = slot ranges that are resolved, substituted and copied out one after the other
    uint16_t *sym16;                                                // [out span] resolved symbol or marker of every output byte (indexed like out)
    uint32_t *mem_err;                                              // per member: 1 inconsistent size, 2 match reaches before the member's first byte
};
cudaError_t spec_init_attributes();
cudaError_t spec_launch_parse(const SpecDev &S, uint32_t r0, uint32_t r1, cudaStream_t st);   // rounds [r0, r1) + verification
cudaError_t spec_launch_tokens(const SpecDev &S, uint32_t cta_lo, uint32_t cta_hi, cudaStream_t st);   // CTAs (128 subsegments each) [cta_lo, cta_hi)
cudaError_t spec_launch_segments(const SpecDev &S, uint32_t part, cudaStream_t st);   // k_seg_plan + k_seg_resolve + k_seg_cuts over the part's slots
cudaError_t spec_launch_subst(const SpecDev &S, uint32_t part, cudaStream_t st);      // k_seg_subst over the chains that start in the part
}

// encode_dev.cuh -- device-side view of one encode batch + stage timing helper.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace b2f {

struct EncDev {
    // inputs
    const uint8_t *in;            // concatenated stream bytes
    uint64_t in_size;             // bytes readable at `in` (kernels never load beyond it)
    const ChunkDesc *chunks; uint32_t n_chunks;
    const BlockDesc *blocks; uint32_t n_blocks;
    const uint32_t *seg0, *pt0, *tile0, *grp0;       // per-chunk prefix arrays, n_chunks + 1 entries each
    uint32_t n_segs, n_ptiles, n_tiles, n_grps;
    uint32_t window, max_len;
    // LZ77 stage
    uint16_t *link;               // [N]  distance to previous same-hash position
    uint32_t *md;                 // [N]  0 | len<<16 | dist  (candidate at every position)
    uint64_t *fix_pos;            // [kFixSlices * fix_cap] global offsets of the positions k_lz_find left to k_lz_fixup
    uint64_t *fix2_pos;           // [kFixSlices * fix_cap] ... that k_lz_fixup left to k_lz_fixup2
    uint32_t *fix2_j;             // [kFixSlices * fix_cap] chunk position the level-1 walk stopped at
    uint32_t *fix_count, *fix2_count;   // [kFixSlices] each
    uint32_t fix_cap;             // queue entries per slice
    uint16_t *exit_tab;           // [n_tiles * 258]
    uint16_t *tile_entry;         // [n_tiles]
    uint32_t *sym;                // [N]  tile-slotted symbol words
    uint32_t *tile_nsym;          // [n_tiles]
    // entropy stage
    uint32_t *hist;               // [n_blocks * 320]
    uint32_t *litcode;            // [n_blocks * 288]  width<<16 | reversed code
    uint32_t *distcode;           // [n_blocks * 32]
    uint32_t *hdr_words;          // [n_blocks * kHdrWords]
    uint32_t *hdr_bits;           // [n_blocks]
    uint32_t *tile_bits;          // [n_tiles]
    uint64_t *tile_bitrel;        // [n_tiles]  bit offset inside the block's symbol area
    uint64_t *blk_bits, *blk_bitoff, *blk_markpos;   // [n_blocks]
    // streams
    uint32_t n_streams;
    const uint32_t *stream_blk0;  // [n_streams + 1]
    const uint64_t *out_base;     // [n_streams] byte offset of the stream in out (16-aligned)
    const uint32_t *hdr_len;      // [n_streams] container header bytes
    uint64_t *stream_end_bits;    // [n_streams] end of the deflate bits, relative to out_base*8
    uint32_t *out_words;          // output buffer viewed as u32 (zero-filled before the entropy stage)
};

constexpr uint32_t kFixSlices = 8;
// bytes of the link[] scratch buffer for an input span of `span` bytes: the links, then the fix-up counters and queues
inline size_t enc_fix_cap(uint64_t span) { return (size_t)(span / 256 + 4096); }
inline size_t enc_link_bytes(uint64_t span) { return (size_t)(((span * 2 + 511) & ~(uint64_t)255) + 256 + (size_t)kFixSlices * enc_fix_cap(span) * 20); }
inline void enc_set_fix(EncDev &E, uint64_t span) {
    uint8_t *base = reinterpret_cast<uint8_t *>(E.link) + ((span * 2 + 511) & ~(uint64_t)255);
    const size_t q = (size_t)kFixSlices * enc_fix_cap(span);
    E.fix_count = reinterpret_cast<uint32_t *>(base); E.fix2_count = E.fix_count + kFixSlices;
    E.fix_pos = reinterpret_cast<uint64_t *>(base + 256); E.fix2_pos = E.fix_pos + q; E.fix2_j = reinterpret_cast<uint32_t *>(E.fix2_pos + q);
    E.fix_cap = (uint32_t)enc_fix_cap(span);
}

struct StageTimer {
    enum { kMax = 96 };
    cudaEvent_t ev[kMax + 1];
    const char *name[kMax];
    int n = 0;
    bool enabled = true, created = false;
    void create() { if (!created) { for (int i = 0; i <= kMax; i++) cudaEventCreate(&ev[i]); created = true; } }
    void destroy() { if (created) { for (int i = 0; i <= kMax; i++) cudaEventDestroy(ev[i]); created = false; } }
    void reset() { n = 0; }
    void mark(cudaStream_t st, const char *nm) { if (!enabled || !created || n >= kMax) return; cudaEventRecord(ev[n], st); name[n++] = nm; }
    void finish(cudaStream_t st) { if (!enabled || !created) return; cudaEventRecord(ev[n], st); }
    // after a stream synchronize
    float stage_ms(int i) const { float ms = 0; if (i < n) cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); return ms; }
    float total_ms() const { float ms = 0; if (n > 0) cudaEventElapsedTime(&ms, ev[0], ev[n]); return ms; }
};

// optional host->device feed of the input, one call per slice of chunks [c0, c1), on the stream that will process the slice
struct SliceFeed { void *self; cudaError_t (*copy)(void *self, uint32_t c0, uint32_t c1, cudaStream_t st); };

// optional pipelined entropy stage: every slice scans / packs its own blocks as soon as its LZ77 stage is done.  ev_scan[g] fires
// when the bit position after slice g's blocks is known (copied to h_pos[g] when h_pos is set: single-stream batches), ev_pack[g]
// when the slice's bytes are packed.
constexpr uint32_t kAuxStreams = 4;      // side streams of a context; slice g of an encode runs on stream g mod kAuxStreams
constexpr uint32_t kMaxSlices = 8;       // = kFixSlices
struct SlicePipe { cudaEvent_t ev_scan[kMaxSlices], ev_pack[kMaxSlices]; uint64_t *h_pos; uint32_t n_slices; };

cudaError_t enc_init_attributes();
cudaError_t enc_launch_lz(const EncDev &E, const uint32_t *h_seg0, const uint32_t *h_pt0, const uint32_t *h_tile0, const uint32_t *h_grp0,
                          cudaStream_t st, StageTimer *tm, cudaStream_t *aux, cudaEvent_t *ev, uint32_t n_aux, const SliceFeed *feed,
                          const ChunkDesc *h_chunks, bool *sliced, SlicePipe *pipe = nullptr);   // h_chunks (host copy of E.chunks) lets slices end on DEFLATE block boundaries and
                                                                      // run k_huff_build / k_tile_bits for their own blocks; *sliced tells the entropy stage
cudaError_t enc_launch_entropy(const EncDev &E, cudaStream_t st, StageTimer *tm, bool sliced);
cudaError_t enc_launch_compact(const EncDev &E, uint64_t *tile_symoff, uint64_t *total, uint32_t *dst, cudaStream_t st);
uint32_t enc_launch_count_lz(bool sliced);
uint32_t enc_launch_count_entropy(bool has_tiles, bool sliced);

}  // namespace b2f

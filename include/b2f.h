/*
 * b2f.h -- C ABI of the B200-native DEFLATE hot path (libb2f.so).
 *
 * This is the drop-in boundary a libflate maintainer binds with `extern "C"` from a sibling
 * crate (libflate itself is #![forbid(unsafe_code)], /root/reference/src/lib.rs:3).  Every entry
 * point names the reference interface it replaces (file:line under sile/libflate @ v2.3.0).
 * Plain pointers and sizes only; no C++ or torch types.  See INTEGRATION.md for the Rust side.
 *
 * Threading: a b2f_ctx owns one CUDA device, its streams and scratch memory.  It is NOT
 * thread-safe; use one ctx per host thread (many ctxs per process / device are fine).
 * All calls are synchronous: when they return, host output buffers are complete and no input
 * pointer is retained.
 *
 * There is no CPU fallback anywhere behind this header: if no CUDA device is usable,
 * b2f_ctx_create fails with B2F_ERR_CUDA and nothing else can be called.
 */
#ifndef B2F_H
#define B2F_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (per call and per stream) ------------------------------------------- */
#define B2F_OK 0
#define B2F_ERR_INVALID_DATA (-1)     /* -> io::ErrorKind::InvalidData   (src/lib.rs:10-29)            */
#define B2F_ERR_UNEXPECTED_EOF (-2)   /* -> io::ErrorKind::UnexpectedEof (read_exact, src/bit.rs:137)   */
#define B2F_ERR_OUTPUT_TOO_SMALL (-3) /* caller retries with a larger buffer                            */
#define B2F_ERR_NOMEM (-4)
#define B2F_ERR_CUDA (-5)             /* message via b2f_last_error                                     */
#define B2F_ERR_INVALID_ARG (-6)

/* ---- container formats ------------------------------------------------------------------ */
#define B2F_FMT_DEFLATE 0     /* libflate::deflate::{Encoder,Decoder}  (src/deflate/{encode,decode}.rs) */
#define B2F_FMT_ZLIB 1        /* libflate::zlib::{Encoder,Decoder}     (src/zlib.rs)                    */
#define B2F_FMT_GZIP 2        /* libflate::gzip::{Encoder,Decoder}     (src/gzip.rs)  first member only */
#define B2F_FMT_GZIP_MULTI 3  /* libflate::gzip::MultiDecoder          (src/gzip.rs:1052-1167) decode   */

/* ---- block modes: deflate::EncodeOptions (src/deflate/encode.rs:17-128) ------------------ */
#define B2F_MODE_DYNAMIC 0    /* default                         */
#define B2F_MODE_FIXED 1      /* .fixed_huffman_codes()          */
#define B2F_MODE_STORED 2     /* .no_compression()               */

/* A schedule entry is either a write size (>= 0) or B2F_SCHED_FLUSH (io::Write::flush()). */
#define B2F_SCHED_FLUSH (-1)

typedef struct b2f_ctx b2f_ctx;

/* Encode options = deflate::EncodeOptions + DefaultLz77EncoderBuilder + the header fields of
 * gzip::HeaderBuilder / zlib::EncodeOptions that influence output bytes. Zero-init then
 * b2f_encode_opts_default(). */
typedef struct b2f_encode_opts {
    uint64_t block_size;       /* EncodeOptions::block_size            (src/deflate/encode.rs:84-87), default 1<<20 */
    uint32_t window_size;      /* DefaultLz77EncoderBuilder::window_size (libflate_lz77/src/default.rs:222-227), default 32768 */
    uint32_t max_length;       /* DefaultLz77EncoderBuilder::max_length  (default.rs:234-239), default 258 */
    int32_t mode;              /* B2F_MODE_*                                                            */
    int32_t zlib_flush_sync;   /* zlib::FlushMode::Sync (src/zlib.rs:150-157): flush() emits 00 00 FF FF */
    uint32_t gzip_mtime;       /* HeaderBuilder::modification_time (src/gzip.rs:171-174); the reference defaults to now() */
    uint8_t gzip_os;           /* HeaderBuilder::os, default 3 (Unix)                                   */
    uint8_t gzip_is_text;      /* HeaderBuilder::text()                                                 */
    uint8_t gzip_is_verified;  /* HeaderBuilder::verify(): header CRC16 as libflate computes it (src/gzip.rs:356-367) */
    uint8_t gzip_has_extra;    /* HeaderBuilder::extra_field()                                          */
    const uint8_t *gzip_extra; /* serialised subfields (id[2], len u16 LE, data)*, total <= 65535       */
    uint32_t gzip_extra_len;
    const char *gzip_filename; /* HeaderBuilder::filename(), NUL terminated or NULL                     */
    const char *gzip_comment;  /* HeaderBuilder::comment(),  NUL terminated or NULL                     */
} b2f_encode_opts;

void b2f_encode_opts_default(b2f_encode_opts *o);

/* ---- context ---------------------------------------------------------------------------- */
int b2f_ctx_create(int device, b2f_ctx **out);
void b2f_ctx_destroy(b2f_ctx *ctx);
const char *b2f_last_error(const b2f_ctx *ctx);  /* valid until the next call on ctx */
const char *b2f_version(void);

/* ---- host memory ---------------------------------------------------------------------------
 * Every batch call accepts ORDINARY (pageable) host memory: the library stages the payload through its own page-locked buffers
 * (three 8 MiB buffers per direction, a few copy threads -- B2F_COPY_THREADS, default 4), overlapping the staging copies with the
 * DMA transfers and the kernels.  A caller that can allocate its buffers here (or register existing ones) saves that copy: the DMA
 * engines then read and write the caller's memory in place.  This is the reference's `Vec<u8>` / `W: Write` boundary
 * (src/deflate/encode.rs:241-249, src/deflate/decode.rs:136-164) -- no CUDA type crosses it. */
int b2f_host_alloc(size_t bytes, void **out);          /* page-locked, usable from any context/device */
void b2f_host_free(void *p);
int b2f_host_register(void *p, size_t bytes);          /* pin an existing allocation in place (page granular; costly: do it once) */
int b2f_host_unregister(void *p);

/* ---- E1: segmentation plan (pure host arithmetic) ----------------------------------------
 * Replaces the bookkeeping of Block::write / CompressBuf::append (src/deflate/encode.rs:277-286,
 * 405-425) and DefaultLz77Encoder::encode (libflate_lz77/src/default.rs:60-68): given the
 * sequence of write sizes (and flush markers) it returns the end offsets of every LZ77 chunk and
 * of every DEFLATE block, including the always-emitted final block.  block_chunks[b] = number of
 * chunks flushed into block b (0 for an empty block).  Arrays need room for n_sched + 2 entries
 * plus in_len / min(block_size, 8*window) entries; pass NULL arrays to only count. */
int b2f_plan_from_writes(const int64_t *sched, size_t n_sched, uint64_t in_len, uint64_t block_size, uint32_t window_size,
                         uint64_t *chunk_ends, size_t *n_chunks, uint64_t *block_ends, uint32_t *block_chunks,
                         uint8_t *block_after_flush, size_t *n_blocks);

/* ---- E2: Lz77Encode trait backend --------------------------------------------------------
 * Replaces DefaultLz77Encoder::{encode,flush} (libflate_lz77/src/default.rs:59-113) for one
 * buffered chunk: codes[k] = byte for Code::Literal, 0x80000000 | length<<16 | distance for
 * Code::Pointer (libflate_lz77/src/lib.rs:28-42).  codes capacity must be >= len.
 * A Rust `B200Lz77Encoder: Lz77Encode` buffers in encode() and replays these into sink.consume(). */
int b2f_lz77_default(b2f_ctx *ctx, const uint8_t *buf, size_t len, uint32_t window_size, uint32_t max_length,
                     uint32_t *codes, size_t *n_codes);

/* ---- whole-stream batch encode -----------------------------------------------------------
 * Replaces {deflate,zlib,gzip}::Encoder::{with_options, write*, flush*, finish}
 * (src/deflate/encode.rs:132-258, src/zlib.rs:522-681, src/gzip.rs:754-908) for n independent
 * streams.  sched[s] == NULL means one write_all(in[s]) (no write at all when in_len[s] == 0).
 * Output is byte-identical to libflate's for the same input, options and write schedule.
 * out_len[s] is always set to the size needed; status[s] is B2F_OK or B2F_ERR_OUTPUT_TOO_SMALL. */
int b2f_encode_batch(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, size_t n_streams,
                     const uint8_t *const *in, const size_t *in_len,
                     const int64_t *const *sched, const size_t *n_sched,
                     uint8_t *const *out, const size_t *out_cap, size_t *out_len, int *status);
/* Upper bound on the encoded size of one stream (any mode, any schedule with n_sched entries). */
size_t b2f_encode_bound(size_t in_len, size_t n_sched, const b2f_encode_opts *opts);

/* ---- whole-stream batch decode -----------------------------------------------------------
 * Replaces {deflate,zlib,gzip}::Decoder::new + read_to_end and gzip::MultiDecoder
 * (src/deflate/decode.rs:8-165, src/zlib.rs:284-410, src/gzip.rs:912-1167).
 * out_len[s]   = bytes decoded; on error, the bytes decoded before the error (what read_to_end
 *               returned plus Decoder::unread_decoded_data()).
 * in_consumed[s] = bytes pulled from the underlying reader (the decoder never reads past its stream).
 * Device-resident inputs (b2f_decode_device): the sub-block decoder fetches aligned 32-bit words, so up to 3 bytes after
 * in_off[s] + in_len[s] may be READ (never interpreted: a block that would end past the stream is rejected); keep d_in at
 * least 4 bytes longer than the last stream.
 * status[s]    = B2F_OK / B2F_ERR_INVALID_DATA / B2F_ERR_UNEXPECTED_EOF / B2F_ERR_OUTPUT_TOO_SMALL. */
int b2f_decode_batch(b2f_ctx *ctx, int fmt, size_t n_streams,
                     const uint8_t *const *in, const size_t *in_len,
                     uint8_t *const *out, const size_t *out_cap, size_t *out_len, size_t *in_consumed, int *status);

/* ---- C1/C2: checksums --------------------------------------------------------------------
 * Replaces checksum::Adler32 / checksum::Crc32 update()+value() (src/checksum.rs:4-33).
 * init[s] is the value() of the bytes hashed so far (NULL: Adler 1 / CRC 0), so calls chain. */
int b2f_adler32_batch(b2f_ctx *ctx, size_t n, const uint8_t *const *buf, const size_t *len, const uint32_t *init, uint32_t *out);
int b2f_crc32_batch(b2f_ctx *ctx, size_t n, const uint8_t *const *buf, const size_t *len, const uint32_t *init, uint32_t *out);

/* ---- device-resident variants (inputs/outputs already in HBM; used by bench.py `value`) ---
 * Same semantics as the batch calls for ONE format/options, but `d_in` is a device pointer to
 * the concatenated inputs (stream s at in_off[s]) and outputs stay on the device at
 * d_out + out_off[s] (capacity out_cap[s]).  Sizes/status come back to the host arrays. */
int b2f_encode_device(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, size_t n_streams,
                      const uint8_t *d_in, const uint64_t *in_off, const size_t *in_len,
                      const int64_t *const *sched, const size_t *n_sched,
                      uint8_t *d_out, const uint64_t *out_off, const size_t *out_cap, size_t *out_len, int *status);
int b2f_decode_device(b2f_ctx *ctx, int fmt, size_t n_streams,
                      const uint8_t *d_in, const uint64_t *in_off, const size_t *in_len,
                      uint8_t *d_out, const uint64_t *out_off, const size_t *out_cap,
                      size_t *out_len, size_t *in_consumed, int *status);
/* framing helpers for the device path: header/trailer bytes are tiny and are produced on the host */
size_t b2f_header_len(int fmt, const b2f_encode_opts *opts);

/* ---- ONE stream split into parts (several contexts / GPUs) ---------------------
 * libflate's blocks are independent (fresh LZ77 table per chunk, libflate_lz77/src/default.rs:73,108; own Huffman codes per block,
 * src/deflate/symbol.rs:321-342), so contiguous runs of whole blocks can be encoded separately.  b2f_plan_from_writes gives the
 * block boundaries for a write schedule; a part = the writes of a run of blocks (the schedule entries are the caller's).
 * b2f_encode_part_device writes the part's raw DEFLATE bits starting at bit 0 of d_out (no container framing; BFINAL + the
 * finish() block only when is_last) and returns their length in BITS plus the part's own CRC-32 / Adler-32.  The assembler --
 * any host code -- places part k at bit offset 8 * header + sum of the earlier parts' bits: shift it on its GPU by (offset mod 8)
 * with b2f_bits_shift_device, copy it out, OR the seam bytes, fold the checksums with the combine functions (Crc32 / Adler32 are
 * linear: src/checksum.rs:4-33) and append b2f_stream_trailer.  The result is byte-identical to the single-call encode. */
int b2f_encode_part_device(b2f_ctx *ctx, const b2f_encode_opts *opts, const uint8_t *d_in, size_t in_len,
                           const int64_t *sched, size_t n_sched, int is_last,
                           uint8_t *d_out, size_t out_cap, uint64_t *out_bits, uint32_t *crc32, uint32_t *adler32);
/* d_dst bit (i + shift) = d_src bit i for i < n_bits (LSB first); d_dst needs (n_bits + 7) / 8 + 1 bytes; shift 0..7 */
int b2f_bits_shift_device(b2f_ctx *ctx, const uint8_t *d_src, uint64_t n_bits, uint32_t shift, uint8_t *d_dst);
uint32_t b2f_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2);       /* value of the concatenation (pure host arithmetic) */
uint32_t b2f_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2);
size_t b2f_stream_header(int fmt, const b2f_encode_opts *opts, uint8_t *out, size_t cap);     /* returns the header length */
size_t b2f_stream_trailer(int fmt, uint32_t crc32, uint32_t adler32, uint64_t total_len, uint8_t *out /* >= 8 bytes */);

/* ---- streaming handles: the Read/Write-shaped surface -----------------------------------
 * A b2f_encoder mirrors Encoder<W,E>: write()/flush() record the schedule and buffer the input,
 * finish() runs the batch path and returns the complete stream (src/deflate/encode.rs:241-249,
 * 196-201).  A b2f_decoder mirrors Decoder<R>: it is constructed over the whole input, decodes on
 * the first read() and then serves read() calls (src/deflate/decode.rs:132-165). */
typedef struct b2f_encoder b2f_encoder;
typedef struct b2f_decoder b2f_decoder;
int b2f_encoder_new(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, b2f_encoder **out);
int b2f_encoder_write(b2f_encoder *e, const uint8_t *buf, size_t len);   /* always consumes all, like Encoder::write */
int b2f_encoder_flush(b2f_encoder *e);
int b2f_encoder_finish(b2f_encoder *e, const uint8_t **out, size_t *out_len); /* buffer owned by e until free */
void b2f_encoder_free(b2f_encoder *e);
int b2f_decoder_new(b2f_ctx *ctx, int fmt, const uint8_t *in, size_t in_len, b2f_decoder **out);
int64_t b2f_decoder_read(b2f_decoder *d, uint8_t *buf, size_t len);      /* >=0 bytes read (0 = EOS), <0 = B2F_ERR_* */
size_t b2f_decoder_unread(const b2f_decoder *d, const uint8_t **ptr);    /* Decoder::unread_decoded_data */
size_t b2f_decoder_consumed(const b2f_decoder *d);
void b2f_decoder_free(b2f_decoder *d);

/* ---- introspection for benchmarks ---------------------------------------------------------*/
typedef struct b2f_stats {
    uint64_t kernel_launches;     /* kernels of this library launched since ctx creation */
    float last_kernel_ms[16];     /* per-stage device time of the last encode/decode call (CUDA events on ctx's stream) */
    uint32_t last_n_stages;
    float last_device_ms;         /* whole device section of the last call */
    uint64_t decode_parallel_streams;  /* DEFLATE streams decoded by the block/sub-block parallel path since ctx creation */
    uint64_t decode_inorder_streams;   /* streams decoded by the in-order kernel (small, irregular or erroneous streams)  */
    uint64_t staged_h2d_bytes;         /* payload bytes that went through the internal pinned staging (pageable caller memory) */
    uint64_t staged_d2h_bytes;
} b2f_stats;
int b2f_get_stats(b2f_ctx *ctx, b2f_stats *out);
const char *b2f_stage_name(b2f_ctx *ctx, uint32_t stage);   /* name of stage i of the last call */
void *b2f_ctx_stream(b2f_ctx *ctx);   /* cudaStream_t the ctx launches on */
/* 1 (default): independent chunk slices of the LZ77 stage run concurrently on internal streams; 0: every kernel runs alone on
 * the ctx stream, so that b2f_get_stats reports one duration per kernel (used for roofline measurements). */
int b2f_ctx_set_overlap(b2f_ctx *ctx, int on);

#ifdef __cplusplus
}
#endif
#endif /* B2F_H */

// b2f_api.cu -- C ABI (include/b2f.h): context, segmentation plan, container framing and the host
// orchestration of the encode / decode / checksum kernels.  Host code here is bookkeeping only
// (sizes, offsets, 10-20 byte headers/trailers); every byte of LZ77 / Huffman / bit-packing / inflate /
// checksum work is done by the CUDA kernels.  There is no CPU fallback: without a device the context cannot
// be created and every entry point needs a context.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "../../include/b2f.h"
#include "common.cuh"
#include "encode_dev.cuh"
#include "decode_dev.cuh"
#include "checksum_dev.cuh"
#include "spec_dev.cuh"
#include "inflate_core.cuh"

using namespace b2f;

// ------------------------------------------------------------------------------------------ context
struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};
struct PinBuf {
    void *p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 4096;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ------------------------------------------------------------------------------------------ host staging
// The ABI takes ordinary (pageable) caller memory ("pinned-memory registration is internal").  A cudaMemcpyAsync from
// or to pageable memory is staged by the driver through one small buffer at ~11 GB/s and blocks the calling thread, so the library
// does the staging itself: the payload moves in kStageChunk pieces through three page-locked buffers per direction; a small pool of
// worker threads copies between caller memory and the staging buffer while the DMA engine moves the previous piece and the kernels
// of earlier pieces run.  Buffers that are already page-locked (b2f_host_alloc / b2f_host_register / cudaHostAlloc) are used in place.
struct CopyPool {
    std::vector<std::thread> th;
    std::mutex mu; std::condition_variable cv_work, cv_done;
    uint8_t *dst = nullptr; const uint8_t *src = nullptr; size_t n = 0, piece = 0;
    std::atomic<size_t> next{0};
    size_t pending = 0; uint64_t gen = 0; bool stop = false;
    void start(unsigned nthreads) {
        for (unsigned i = 0; i + 1 < nthreads; i++) th.emplace_back([this] { run(); });   // the calling thread is the last worker
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> lk(mu); cv_work.wait(lk, [&] { return stop || gen != seen; }); if (stop) return; seen = gen; }
            work();
            { std::lock_guard<std::mutex> lk(mu); if (--pending == 0) cv_done.notify_all(); }
        }
    }
    void work() {
        for (;;) {
            const size_t o = next.fetch_add(piece);
            if (o >= n) return;
            memcpy(dst + o, src + o, std::min(piece, n - o));
        }
    }
    void copy(void *d, const void *s_, size_t bytes) {
        if (th.empty() || bytes < (1u << 20)) { memcpy(d, s_, bytes); return; }
        { std::lock_guard<std::mutex> lk(mu); dst = (uint8_t *)d; src = (const uint8_t *)s_; n = bytes; piece = 256u << 10; next = 0; pending = th.size(); gen++; }
        cv_work.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu); cv_done.wait(lk, [&] { return pending == 0; });
    }
    void shutdown() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_work.notify_all();
        for (auto &t : th) t.join();
        th.clear();
    }
};
constexpr size_t kStageChunk = 8u << 20;
struct Stager {
    PinBuf buf[3]; cudaEvent_t ev[3] = { nullptr, nullptr, nullptr }; bool busy[3] = { false, false, false };
    cudaError_t init() {
        for (int i = 0; i < 3; i++) {
            cudaError_t e = buf[i].ensure(kStageChunk); if (e != cudaSuccess) return e;
            if (!ev[i]) { e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
        }
        return cudaSuccess;
    }
    void release() { for (int i = 0; i < 3; i++) { buf[i].release(); if (ev[i]) cudaEventDestroy(ev[i]); ev[i] = nullptr; busy[i] = false; } }
};

enum { NB_IN, NB_LINK, NB_MD, NB_SYM, NB_EXIT, NB_TILE, NB_BLK, NB_DESC, NB_OUT, NB_MISC, NB_CK, NB_DEC_META, NB_DEC_OUT, NB_DEC_CAND, NB_DEC_BLK, NB_DEC_SER, NB_SPEC_TAB, NB_SPEC_SEG, NB_SPEC_TOK, NB_SPEC_SEL, NB_SPEC_SYM, NB_COUNT };

struct b2f_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DevBuf buf[NB_COUNT];
    PinBuf pin_meta, pin_res, pin_ck, pin_win, pin_cand, pin_blk, pin_ser, pin_sel;
    StageTimer tm;
    std::string err;
    b2f_stats stats;
    const char *stage_names[16];
    bool last_is_decode = false;
    uint64_t n_spec_members = 0, n_inorder_members = 0;
    cudaStream_t aux[4] = { nullptr, nullptr, nullptr, nullptr };
    // decode: a page-locked input arrives in pieces on copy_st; the block finder follows piece by piece (inflate_round)
    static constexpr uint32_t kFeedMax = 8;
    cudaEvent_t feed_ev[kFeedMax] = {};
    uint64_t feed_end[kFeedMax] = {};   // device offset up to which the input has arrived with piece k
    uint32_t feed_n = 0;
    cudaStream_t copy_st = nullptr;   // early D2H of an encode's finished slices (the aux streams are busy with the slices themselves)
    cudaEvent_t aux_ev[kMaxSlices + 1] = {};
    uint32_t enc_slices = 4;       // slices of the encode pipeline (B2F_ENC_SLICES, 2..kMaxSlices)
    cudaEvent_t part_ev[b2f::kMaxParts] = {};
    int overlap = 1;               // run independent chunk slices of the LZ77 stage on separate streams
    CopyPool pool; Stager st_in, st_out; unsigned copy_threads = 4; size_t stage_rr = 0;
    SlicePipe pipe;                // events of the encode's per-slice entropy stage
    PinBuf pin_pos;                // bit positions after each slice (single-stream encodes: the packed bytes leave slice by slice)
    uint32_t max_parts = 4;        // pipeline depth of the decode's LZ77 resolution + device->host copies (B2F_DECODE_PARTS, <= kMaxParts)
    uint64_t staged_h2d = 0, staged_d2h = 0;     // bytes that went through the internal staging (pageable caller memory)
    std::vector<uint64_t> last_good;             // per stream of the last decode call: bytes of the blocks that completed before an error
};

static thread_local std::string g_create_err;

#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__); return B2F_ERR_CUDA; } } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static bool is_pinned_host(const void *p) {
    cudaPointerAttributes pa;
    const bool pinned = p && cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}
// host -> device, stream ordered.  Pinned source: one DMA.  Pageable source: staged (the calling thread and the pool copy piece k+1
// into a staging buffer while the DMA of piece k and everything queued before it run).
static cudaError_t h2d_copy(b2f_ctx *ctx, uint8_t *d_dst, const uint8_t *h_src, size_t n, cudaStream_t st, bool pinned) {
    if (!n) return cudaSuccess;
    if (pinned) return cudaMemcpyAsync(d_dst, h_src, n, cudaMemcpyHostToDevice, st);
    Stager &S = ctx->st_in;
    cudaError_t e = S.init(); if (e != cudaSuccess) return e;
    for (size_t o = 0; o < n; o += kStageChunk) {
        const size_t len = std::min(kStageChunk, n - o);
        const size_t k = ctx->stage_rr++ % 3;
        if (S.busy[k]) { e = cudaEventSynchronize(S.ev[k]); if (e != cudaSuccess) return e; }
        ctx->pool.copy(S.buf[k].p, h_src + o, len);
        e = cudaMemcpyAsync(d_dst + o, S.buf[k].p, len, cudaMemcpyHostToDevice, st); if (e != cudaSuccess) return e;
        e = cudaEventRecord(S.ev[k], st); if (e != cudaSuccess) return e;
        S.busy[k] = true;
    }
    ctx->staged_h2d += n;
    return cudaSuccess;
}
// device -> host.  The source must be complete in stream order on `st`.  Returns when the bytes are in h_dst (pageable) or queued (pinned).
static cudaError_t d2h_copy(b2f_ctx *ctx, uint8_t *h_dst, const uint8_t *d_src, size_t n, cudaStream_t st, bool pinned) {
    if (!n) return cudaSuccess;
    if (pinned) return cudaMemcpyAsync(h_dst, d_src, n, cudaMemcpyDeviceToHost, st);
    Stager &S = ctx->st_out;
    cudaError_t e = S.init(); if (e != cudaSuccess) return e;
    const size_t nchunks = (n + kStageChunk - 1) / kStageChunk;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t o = c * kStageChunk, len = std::min(kStageChunk, n - o);
        cudaError_t e2 = cudaMemcpyAsync(S.buf[c % 3].p, d_src + o, len, cudaMemcpyDeviceToHost, st); if (e2 != cudaSuccess) return e2;
        return cudaEventRecord(S.ev[c % 3], st);
    };
    for (size_t c = 0; c < std::min<size_t>(2, nchunks); c++) { e = issue(c); if (e != cudaSuccess) return e; }
    for (size_t c = 0; c < nchunks; c++) {
        if (c + 2 < nchunks) { e = issue(c + 2); if (e != cudaSuccess) return e; }      // slot (c+2)%3 was drained by iteration c-1
        e = cudaEventSynchronize(S.ev[c % 3]); if (e != cudaSuccess) return e;
        const size_t o = c * kStageChunk, len = std::min(kStageChunk, n - o);
        ctx->pool.copy(h_dst + o, S.buf[c % 3].p, len);
    }
    ctx->staged_d2h += n;
    return cudaSuccess;
}

extern "C" const char *b2f_version(void) { return "libb2f 0.2 (sm_100a)"; }

// ---- page-locked host memory for callers that want the DMA engines to read/write their buffers in place
extern "C" int b2f_host_alloc(size_t bytes, void **out) {
    if (!out) return B2F_ERR_INVALID_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable) == cudaSuccess ? B2F_OK : B2F_ERR_NOMEM;
}
extern "C" void b2f_host_free(void *p) { if (p) cudaFreeHost(p); }
extern "C" int b2f_host_register(void *p, size_t bytes) {
    if (!p || !bytes) return B2F_ERR_INVALID_ARG;
    return cudaHostRegister(p, bytes, cudaHostRegisterPortable) == cudaSuccess ? B2F_OK : B2F_ERR_CUDA;
}
extern "C" int b2f_host_unregister(void *p) { return (p && cudaHostUnregister(p) == cudaSuccess) ? B2F_OK : B2F_ERR_INVALID_ARG; }

extern "C" int b2f_ctx_create(int device, b2f_ctx **out) {
    if (!out) return B2F_ERR_INVALID_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        g_create_err = e != cudaSuccess ? cudaGetErrorString(e) : "no such CUDA device";
        return B2F_ERR_CUDA;                   // no CPU fallback by design
    }
    b2f_ctx *ctx = new b2f_ctx();
    ctx->device = device;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        enc_init_attributes() != cudaSuccess || dec_init_attributes() != cudaSuccess || spec_init_attributes() != cudaSuccess || checksum_init_tables() != cudaSuccess) {
        g_create_err = cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return B2F_ERR_CUDA;
    }
    ctx->tm.create();
    for (auto &a : ctx->aux) cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->copy_st, cudaStreamNonBlocking);
    for (auto &ev : ctx->feed_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : ctx->aux_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (auto &ev : ctx->part_ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (uint32_t i = 0; i < kMaxSlices; i++) { cudaEventCreateWithFlags(&ctx->pipe.ev_scan[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&ctx->pipe.ev_pack[i], cudaEventDisableTiming); }
    if (const char *e = getenv("B2F_ENC_SLICES")) { const int v = atoi(e); if (v >= 2 && v <= (int)kMaxSlices) ctx->enc_slices = (uint32_t)v; }
    ctx->pipe.h_pos = nullptr; ctx->pipe.n_slices = 0;
    ctx->pin_pos.ensure(64);
    if (const char *o = getenv("B2F_OVERLAP")) ctx->overlap = atoi(o);
    if (const char *o = getenv("B2F_COPY_THREADS")) ctx->copy_threads = (unsigned)std::max(1, atoi(o));
    if (const char *o = getenv("B2F_DECODE_PARTS")) ctx->max_parts = (uint32_t)std::min<int>(kMaxParts, std::max(1, atoi(o)));
    ctx->copy_threads = std::min(ctx->copy_threads, std::max(1u, std::thread::hardware_concurrency()));
    ctx->pool.start(ctx->copy_threads);
    *out = ctx;
    return B2F_OK;
}
extern "C" void b2f_ctx_destroy(b2f_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->pool.shutdown();
    ctx->st_in.release(); ctx->st_out.release();
    for (auto &b : ctx->buf) b.release();
    ctx->pin_meta.release(); ctx->pin_res.release(); ctx->pin_ck.release(); ctx->pin_win.release(); ctx->pin_cand.release(); ctx->pin_blk.release(); ctx->pin_ser.release(); ctx->pin_sel.release();
    ctx->tm.destroy();
    for (auto &a : ctx->aux) if (a) cudaStreamDestroy(a);
    if (ctx->copy_st) cudaStreamDestroy(ctx->copy_st);
    for (auto &ev : ctx->feed_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->aux_ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : ctx->part_ev) if (ev) cudaEventDestroy(ev);
    for (uint32_t i = 0; i < kMaxSlices; i++) { if (ctx->pipe.ev_scan[i]) cudaEventDestroy(ctx->pipe.ev_scan[i]); if (ctx->pipe.ev_pack[i]) cudaEventDestroy(ctx->pipe.ev_pack[i]); }
    ctx->pin_pos.release();
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char *b2f_last_error(const b2f_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
extern "C" int b2f_ctx_set_overlap(b2f_ctx *ctx, int on) { if (!ctx) return B2F_ERR_INVALID_ARG; ctx->overlap = on ? 1 : 0; return B2F_OK; }
extern "C" void *b2f_ctx_stream(b2f_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int b2f_get_stats(b2f_ctx *ctx, b2f_stats *out) { if (!ctx || !out) return B2F_ERR_INVALID_ARG; ctx->stats.decode_parallel_streams = ctx->n_spec_members; ctx->stats.decode_inorder_streams = ctx->n_inorder_members; ctx->stats.staged_h2d_bytes = ctx->staged_h2d; ctx->stats.staged_d2h_bytes = ctx->staged_d2h; *out = ctx->stats; return B2F_OK; }
extern "C" const char *b2f_stage_name(b2f_ctx *ctx, uint32_t i) { return (ctx && i < 16 && ctx->stage_names[i]) ? ctx->stage_names[i] : ""; }

extern "C" void b2f_encode_opts_default(b2f_encode_opts *o) {
    memset(o, 0, sizeof *o);
    o->block_size = 1u << 20; o->window_size = 32768; o->max_length = 258; o->mode = B2F_MODE_DYNAMIC; o->gzip_os = 3;
}

static void collect_stats(b2f_ctx *ctx, bool is_decode) {
    ctx->last_is_decode = is_decode;
    // stages that ran several times (retries, sync points) are summed under one name, in order of first appearance
    uint32_t n = 0;
    for (uint32_t i = 0; i < 16; i++) { ctx->stats.last_kernel_ms[i] = 0.f; ctx->stage_names[i] = nullptr; }
    for (int i = 0; i < ctx->tm.n; i++) {
        uint32_t k = 0;
        while (k < n && strcmp(ctx->stage_names[k], ctx->tm.name[i]) != 0) k++;
        if (k == n) { if (n == 16) continue; ctx->stage_names[n++] = ctx->tm.name[i]; }
        ctx->stats.last_kernel_ms[k] += ctx->tm.stage_ms(i);
    }
    ctx->stats.last_n_stages = n;
    ctx->stats.last_device_ms = ctx->tm.total_ms();
}

// ------------------------------------------------------------------------------------------ E1: plan
namespace {
struct PlanOut {
    std::vector<uint64_t> chunk_ends;      // stream-relative
    std::vector<uint64_t> block_ends;
    std::vector<uint32_t> block_chunks;
    std::vector<uint8_t> block_after_flush;   // block was closed by an explicit flush() (candidate for the zlib sync marker)
};
// Block::write / CompressBuf::append / DefaultLz77Encoder::encode bookkeeping (encode.rs:277-286,405-425; default.rs:60-68)
// emit_final = false: the stream continues in another part (b2f_encode_part_device): no finish() block; returns false when the part
// does not end exactly on a block boundary.
bool plan_stream(const int64_t *sched, size_t n_sched, uint64_t in_len, uint64_t block_size, uint32_t window, PlanOut &P, bool emit_final = true) {
    const uint64_t chunk_thresh = (uint64_t)window * 8;
    uint64_t lz = 0, orig = 0, pos = 0; uint32_t chunks_in_block = 0;
    auto end_chunk = [&]() { P.chunk_ends.push_back(pos); chunks_in_block++; lz = 0; };
    auto flush_block = [&](bool after_flush) {
        if (lz > 0) end_chunk();
        P.block_ends.push_back(pos); P.block_chunks.push_back(chunks_in_block); P.block_after_flush.push_back(after_flush ? 1 : 0);
        chunks_in_block = 0; orig = 0;
    };
    int64_t one = (int64_t)in_len;
    if (!sched) { sched = &one; n_sched = in_len ? 1 : 0; }
    for (size_t k = 0; k < n_sched; k++) {
        if (sched[k] < 0) { flush_block(true); continue; }
        uint64_t w = (uint64_t)sched[k]; if (pos + w > in_len) w = in_len - pos;
        pos += w; orig += w; lz += w;
        if (lz >= chunk_thresh) end_chunk();
        while (orig >= block_size) flush_block(false);
    }
    if (!emit_final) return lz == 0 && orig == 0 && !P.block_ends.empty();
    flush_block(false);                    // finish(): always one more block, possibly empty (encode.rs:296-303)
    return true;
}
}  // namespace

extern "C" int b2f_plan_from_writes(const int64_t *sched, size_t n_sched, uint64_t in_len, uint64_t block_size, uint32_t window_size,
                                    uint64_t *chunk_ends, size_t *n_chunks, uint64_t *block_ends, uint32_t *block_chunks,
                                    uint8_t *block_after_flush, size_t *n_blocks) {
    if (block_size == 0 || window_size == 0) return B2F_ERR_INVALID_ARG;
    PlanOut P;
    plan_stream(sched, n_sched, in_len, block_size, window_size > 32768 ? 32768 : window_size, P);
    if (n_chunks) *n_chunks = P.chunk_ends.size();
    if (n_blocks) *n_blocks = P.block_ends.size();
    if (chunk_ends) memcpy(chunk_ends, P.chunk_ends.data(), P.chunk_ends.size() * 8);
    if (block_ends) memcpy(block_ends, P.block_ends.data(), P.block_ends.size() * 8);
    if (block_chunks) memcpy(block_chunks, P.block_chunks.data(), P.block_chunks.size() * 4);
    if (block_after_flush) memcpy(block_after_flush, P.block_after_flush.data(), P.block_after_flush.size());
    return B2F_OK;
}

// ------------------------------------------------------------------------------------------ framing (G1)
namespace {
uint32_t host_crc32(const uint8_t *p, size_t n) {   // only for the <= ~64 KiB gzip header CRC16 (gzip.rs:356-367), not a data path
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) { c ^= p[i]; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; }
    return ~c;
}
void gzip_header(const b2f_encode_opts &o, bool with_hcrc, std::vector<uint8_t> &h) {   // Header::write_to (gzip.rs:368-389)
    uint8_t flags = 0;
    if (o.gzip_is_text) flags |= 1;
    if (with_hcrc && o.gzip_is_verified) flags |= 2;
    if (o.gzip_has_extra) flags |= 4;
    if (o.gzip_filename) flags |= 8;
    if (o.gzip_comment) flags |= 16;
    const uint8_t b[10] = { 31, 139, 8, flags, (uint8_t)o.gzip_mtime, (uint8_t)(o.gzip_mtime >> 8), (uint8_t)(o.gzip_mtime >> 16),
                            (uint8_t)(o.gzip_mtime >> 24), 0 /* XFL: CompressionLevel::Unknown */, o.gzip_os };
    h.insert(h.end(), b, b + 10);
    if (o.gzip_has_extra) { h.push_back((uint8_t)o.gzip_extra_len); h.push_back((uint8_t)(o.gzip_extra_len >> 8)); h.insert(h.end(), o.gzip_extra, o.gzip_extra + o.gzip_extra_len); }
    if (o.gzip_filename) h.insert(h.end(), (const uint8_t *)o.gzip_filename, (const uint8_t *)o.gzip_filename + strlen(o.gzip_filename) + 1);
    if (o.gzip_comment) h.insert(h.end(), (const uint8_t *)o.gzip_comment, (const uint8_t *)o.gzip_comment + strlen(o.gzip_comment) + 1);
    if (with_hcrc && o.gzip_is_verified) {
        std::vector<uint8_t> t; gzip_header(o, false, t);
        uint32_t c = host_crc32(t.data(), t.size());
        h.push_back((uint8_t)c); h.push_back((uint8_t)(c >> 8));
    }
}
void zlib_header(const b2f_encode_opts &o, std::vector<uint8_t> &h) {                   // zlib.rs:212-220, 267-279
    uint32_t ws = o.window_size > 32768 ? 32768 : o.window_size;
    uint8_t cinfo = ws > 16384 ? 7 : ws > 8192 ? 6 : ws > 4096 ? 5 : ws > 2048 ? 4 : ws > 1024 ? 3 : ws > 512 ? 2 : ws > 256 ? 1 : 0;
    uint8_t level = o.mode == B2F_MODE_STORED ? 0 : 2;
    uint8_t cmf = (uint8_t)((cinfo << 4) | 8), flg = (uint8_t)(level << 6);
    uint16_t check = (uint16_t)(((uint16_t)cmf << 8) + flg);
    if (check % 31 != 0) flg = (uint8_t)(flg + (31 - check % 31));
    h.push_back(cmf); h.push_back(flg);
}
void make_header(int fmt, const b2f_encode_opts &o, std::vector<uint8_t> &h) {
    h.clear();
    if (fmt == B2F_FMT_GZIP) gzip_header(o, true, h);
    else if (fmt == B2F_FMT_ZLIB) zlib_header(o, h);
}
size_t trailer_len(int fmt) { return fmt == B2F_FMT_GZIP ? 8 : fmt == B2F_FMT_ZLIB ? 4 : 0; }
}  // namespace

extern "C" size_t b2f_header_len(int fmt, const b2f_encode_opts *opts) {
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    std::vector<uint8_t> h; make_header(fmt, *opts, h); return h.size();
}
extern "C" size_t b2f_encode_bound(size_t in_len, size_t n_sched, const b2f_encode_opts *opts) {
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    size_t bs = opts->block_size ? (size_t)opts->block_size : 1;
    if (opts->mode == B2F_MODE_STORED && bs > 0xFFFF) bs = 0xFFFF;
    size_t nblocks = n_sched + in_len / bs + 2;
    size_t hdr = 32 + (opts->gzip_has_extra ? opts->gzip_extra_len + 2 : 0) + (opts->gzip_filename ? strlen(opts->gzip_filename) + 1 : 0) +
                 (opts->gzip_comment ? strlen(opts->gzip_comment) + 1 : 0);
    return hdr + in_len + in_len / 2 + nblocks * 640 + 64;
}

// ------------------------------------------------------------------------------------------ checksums on device data
namespace {
// d_base + off[s], len[s] for n buffers that are already on the device; results to host arrays.
int run_checksums(b2f_ctx *ctx, const uint8_t *d_base, const std::vector<uint64_t> &off, const std::vector<uint64_t> &len,
                  bool do_crc, bool do_adler, const uint32_t *init, std::vector<uint32_t> &crc, std::vector<uint32_t> &adler) {
    const size_t n = off.size();
    crc.assign(n, 0); adler.assign(n, 1);
    if (!n) return B2F_OK;
    std::vector<uint64_t> piece0(n + 1, 0); uint32_t span_rows = 0;
    checksum_plan(d_base, off.data(), len.data(), n, piece0.data(), &span_rows);
    // layout in NB_CK: off | len | span0 | acc_a | acc_b | acc_crc | init | out_crc | out_adler
    size_t o_off = 0, o_len = o_off + n * 8, o_p0 = o_len + n * 8, o_a = o_p0 + (n + 1) * 8, o_b = o_a + n * 8, o_c = o_b + n * 8,
           o_init = o_c + n * 4, o_oc = o_init + n * 4, o_oa = o_oc + n * 4, total = align_up(o_oa + n * 4, 16);
    CK(ctx->buf[NB_CK].ensure(total));
    CK(ctx->pin_ck.ensure(total));
    uint8_t *hm = ctx->pin_ck.as<uint8_t>();
    memset(hm, 0, total);
    memcpy(hm + o_off, off.data(), n * 8); memcpy(hm + o_len, len.data(), n * 8); memcpy(hm + o_p0, piece0.data(), (n + 1) * 8);
    if (init) memcpy(hm + o_init, init, n * 4);
    uint8_t *dm = ctx->buf[NB_CK].as<uint8_t>();
    CK(cudaMemcpyAsync(dm, hm, total, cudaMemcpyHostToDevice, ctx->stream));
    ChecksumDev C;
    C.in = d_base; C.off = (const uint64_t *)(dm + o_off); C.len = (const uint64_t *)(dm + o_len); C.span0 = (const uint64_t *)(dm + o_p0);
    C.n_spans = piece0[n]; C.span_rows = span_rows; C.n_streams = (uint32_t)n;
    C.acc_a = (uint64_t *)(dm + o_a); C.acc_b = (uint64_t *)(dm + o_b); C.acc_crc = (uint32_t *)(dm + o_c);
    C.init_crc = (init && do_crc) ? (const uint32_t *)(dm + o_init) : nullptr;
    C.init_adler = (init && do_adler) ? (const uint32_t *)(dm + o_init) : nullptr;
    C.out_crc = (uint32_t *)(dm + o_oc); C.out_adler = (uint32_t *)(dm + o_oa);
    ctx->tm.mark(ctx->stream, "checksum");
    CK(checksum_launch(C, do_crc, do_adler, ctx->stream, &ctx->tm));
    ctx->stats.kernel_launches += (C.n_spans ? 1 : 0) + 1;
    CK(ctx->pin_res.ensure(n * 8));
    uint32_t *hr = ctx->pin_res.as<uint32_t>();
    CK(cudaMemcpyAsync(hr, dm + o_oc, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t s = 0; s < n; s++) { crc[s] = hr[s]; adler[s] = hr[n + s]; }
    return B2F_OK;
}

int checksum_batch_host(b2f_ctx *ctx, size_t n, const uint8_t *const *buf, const size_t *len, const uint32_t *init, uint32_t *out, bool is_crc) {
    if (!ctx || (n && (!buf || !len || !out))) return B2F_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    std::vector<uint64_t> off(n), ln(n); size_t total = 0;
    for (size_t s = 0; s < n; s++) { off[s] = total; ln[s] = len[s]; total += align_up(len[s], 16); }
    CK(ctx->buf[NB_IN].ensure(total + 256));
    uint8_t *d = ctx->buf[NB_IN].as<uint8_t>();
    for (size_t s = 0; s < n; s++) if (len[s]) CK(h2d_copy(ctx, d + off[s], buf[s], len[s], ctx->stream, is_pinned_host(buf[s])));
    std::vector<uint32_t> crc, adler;
    int rc = run_checksums(ctx, d, off, ln, is_crc, !is_crc, init, crc, adler);
    if (rc) return rc;
    ctx->tm.finish(ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    for (size_t s = 0; s < n; s++) out[s] = is_crc ? crc[s] : adler[s];
    collect_stats(ctx, false);
    return B2F_OK;
}
}  // namespace

extern "C" int b2f_adler32_batch(b2f_ctx *ctx, size_t n, const uint8_t *const *buf, const size_t *len, const uint32_t *init, uint32_t *out) {
    return checksum_batch_host(ctx, n, buf, len, init, out, false);
}
extern "C" int b2f_crc32_batch(b2f_ctx *ctx, size_t n, const uint8_t *const *buf, const size_t *len, const uint32_t *init, uint32_t *out) {
    return checksum_batch_host(ctx, n, buf, len, init, out, true);
}

// ------------------------------------------------------------------------------------------ encode core
namespace {
struct EncPlan {
    std::vector<ChunkDesc> chunks;
    std::vector<BlockDesc> blocks;
    std::vector<uint32_t> seg0, pt0, tile0, grp0;      // per chunk prefixes (+1)
    std::vector<uint32_t> stream_blk0;                  // per stream (+1)
    uint64_t total_in = 0;
};

// Fills plan for streams laid out at in_off[] in the device input.
int build_plan(b2f_ctx *ctx, const b2f_encode_opts &o, size_t n_streams, const uint64_t *in_off, const size_t *in_len,
               const int64_t *const *sched, const size_t *n_sched, int fmt, EncPlan &P, bool emit_final = true) {
    const uint32_t window = o.window_size > 32768 ? 32768 : o.window_size;
    P.stream_blk0.assign(1, 0);
    P.seg0.assign(1, 0); P.pt0.assign(1, 0); P.tile0.assign(1, 0); P.grp0.assign(1, 0);
    for (size_t s = 0; s < n_streams; s++) {
        PlanOut po;
        if (!plan_stream(sched ? sched[s] : nullptr, (sched && n_sched) ? n_sched[s] : 0, in_len[s], o.block_size, window, po, emit_final)) {
            ctx->err = "a part that is not the last one must end exactly on a DEFLATE block boundary (see b2f_plan_from_writes)"; return B2F_ERR_INVALID_ARG;
        }
        uint64_t cstart = 0; size_t ci = 0;
        for (size_t b = 0; b < po.block_ends.size(); b++) {
            BlockDesc bd; memset(&bd, 0, sizeof bd);
            bd.stream = (uint32_t)s; bd.chunk0 = (uint32_t)P.chunks.size(); bd.nchunks = po.block_chunks[b];
            bd.tile0 = P.tile0.back();
            bd.is_final = emit_final && b + 1 == po.block_ends.size();
            bd.sync_after = (fmt == B2F_FMT_ZLIB && o.zlib_flush_sync && po.block_after_flush[b]) ? 1 : 0;
            bd.fixed = o.mode == B2F_MODE_FIXED;
            for (uint32_t k = 0; k < bd.nchunks; k++, ci++) {
                uint64_t cend = po.chunk_ends[ci];
                uint64_t len = cend - cstart;
                if (len >= (1ull << 31)) { ctx->err = "LZ77 chunk larger than 2 GiB is not supported"; return B2F_ERR_INVALID_ARG; }
                ChunkDesc cd; cd.off = in_off[s] + cstart; cd.len = (uint32_t)len; cd.block = (uint32_t)P.blocks.size();
                P.chunks.push_back(cd);
                uint32_t nt = (uint32_t)((len + kTile - 1) / kTile);
                P.seg0.push_back(P.seg0.back() + (uint32_t)((len + kSeg - 1) / kSeg));
                P.pt0.push_back(P.pt0.back() + (uint32_t)((len + kPTile - 1) / kPTile));
                P.tile0.push_back(P.tile0.back() + nt);
                P.grp0.push_back(P.grp0.back() + (nt + kGrpTiles - 1) / kGrpTiles);
                cstart = cend;
            }
            bd.ntiles = P.tile0.back() - bd.tile0;
            P.blocks.push_back(bd);
        }
        P.stream_blk0.push_back((uint32_t)P.blocks.size());
        P.total_in += in_len[s];
    }
    return B2F_OK;
}

template <class T> T *carve(uint8_t *&p, size_t count) { T *r = reinterpret_cast<T *>(p); p += align_up(count * sizeof(T), 256); return r; }
template <class T> size_t carve_size(size_t count) { return align_up(count * sizeof(T), 256); }

// Framing kernel: container header bytes + trailer (checksums are already on the device).
__global__ void k_write_framing(uint8_t *out, const uint64_t *out_base, const uint64_t *stream_end_bits, const uint8_t *hdr, uint32_t hdr_len,
                                int fmt, const uint32_t *crc, const uint32_t *adler, const uint64_t *in_len, uint64_t *out_len, uint32_t n) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    uint8_t *o = out + out_base[s];
    for (uint32_t i = 0; i < hdr_len; i++) o[i] = hdr[i];
    uint64_t end = (stream_end_bits[s] + 7) >> 3;
    if (fmt == B2F_FMT_GZIP) {          // Trailer::write_to (gzip.rs:114-121): CRC32 LE, ISIZE LE (mod 2^32)
        uint32_t c = crc[s], z = (uint32_t)in_len[s];
        for (int k = 0; k < 4; k++) o[end + k] = (uint8_t)(c >> (8 * k));
        for (int k = 0; k < 4; k++) o[end + 4 + k] = (uint8_t)(z >> (8 * k));
        end += 8;
    } else if (fmt == B2F_FMT_ZLIB) {   // zlib.rs:630-638: Adler-32 big endian
        uint32_t a = adler[s];
        for (int k = 0; k < 4; k++) o[end + k] = (uint8_t)(a >> (8 * (3 - k)));
        end += 4;
    }
    out_len[s] = end;
}

struct EncodeJob {
    // device-resident inputs
    const uint8_t *d_in; std::vector<uint64_t> in_off; std::vector<uint64_t> in_len;
    const uint8_t *const *h_in = nullptr;     // when set, the inputs still live on the host: encode_on_device copies them slice by slice
    std::vector<char> h_pinned;               // per stream: the host input is page-locked (DMA in place) or pageable (staged)
    uint8_t *h_out = nullptr; bool h_out_pinned = false; size_t h_out_cap = 0;   // single-stream host encode: destination of the early copies
    uint64_t h_copied = 0;                    // bytes of the stream already copied to h_out when encode_on_device returns
    bool part = false, part_last = true;      // b2f_encode_part_device: raw blocks of one part of a stream (no container framing)
    std::vector<uint64_t> out_bits;           // per stream: length of the DEFLATE bits (parts are not byte aligned)
    // results
    std::vector<uint64_t> out_base; std::vector<uint64_t> out_len;   // in ctx->buf[NB_OUT]
};

// Runs the whole device pipeline for compressed modes.  On return the encoded streams sit in ctx->buf[NB_OUT] at job.out_base[s].
int encode_on_device(b2f_ctx *ctx, int fmt, const b2f_encode_opts &o, size_t n_streams, const int64_t *const *sched, const size_t *n_sched, EncodeJob &job) {
    job.out_base.clear(); job.out_len.clear();
    if (n_streams == 0) return B2F_OK;
    EncPlan P;
    std::vector<size_t> lens(n_streams);
    for (size_t s = 0; s < n_streams; s++) lens[s] = (size_t)job.in_len[s];
    int rc = build_plan(ctx, o, n_streams, job.in_off.data(), lens.data(), sched, n_sched, fmt, P, !job.part || job.part_last);
    if (rc) return rc;
    std::vector<uint8_t> hdr; make_header(fmt, o, hdr);
    const size_t tl = trailer_len(fmt);
    // output slots
    job.out_base.resize(n_streams); job.out_len.assign(n_streams, 0);
    uint64_t out_total = 0;
    for (size_t s = 0; s < n_streams; s++) {
        job.out_base[s] = out_total;
        size_t nb = P.stream_blk0[s + 1] - P.stream_blk0[s];
        out_total += align_up(hdr.size() + (size_t)job.in_len[s] + (size_t)job.in_len[s] / 2 + nb * 640 + tl + 64, 256);
    }
    const uint32_t n_chunks = (uint32_t)P.chunks.size(), n_blocks = (uint32_t)P.blocks.size(), n_tiles = P.tile0.back();
    // span of the device input touched by chunks (scratch arrays are indexed by global input offset)
    uint64_t in_span = 0;
    for (size_t s = 0; s < n_streams; s++) in_span = std::max<uint64_t>(in_span, job.in_off[s] + job.in_len[s]);

    CK(ctx->buf[NB_LINK].ensure(enc_link_bytes(in_span)));
    CK(ctx->buf[NB_MD].ensure(in_span * 4 + 256));
    CK(ctx->buf[NB_SYM].ensure(in_span * 4 + 256));
    CK(ctx->buf[NB_EXIT].ensure((size_t)n_tiles * kExitW * 2 + 256));
    size_t tile_bytes = carve_size<uint16_t>(n_tiles) + 2 * carve_size<uint32_t>(n_tiles) + carve_size<uint64_t>(n_tiles);
    CK(ctx->buf[NB_TILE].ensure(tile_bytes + 256));
    size_t blk_bytes = carve_size<uint32_t>((size_t)n_blocks * kHistStride) + carve_size<uint32_t>((size_t)n_blocks * kLitStride) +
                       carve_size<uint32_t>((size_t)n_blocks * kDistStride) + carve_size<uint32_t>((size_t)n_blocks * kHdrWords) +
                       carve_size<uint32_t>(n_blocks) + 3 * carve_size<uint64_t>(n_blocks);
    CK(ctx->buf[NB_BLK].ensure(blk_bytes + 256));
    CK(ctx->buf[NB_OUT].ensure(out_total + 256));
    // descriptors: chunks | blocks | 4 prefixes | stream_blk0 | out_base | hdr_len | in_len | hdr bytes  (one H2D copy)
    size_t d_total = carve_size<ChunkDesc>(n_chunks) + carve_size<BlockDesc>(n_blocks) + 4 * carve_size<uint32_t>(n_chunks + 1) +
                     carve_size<uint32_t>(n_streams + 1) + carve_size<uint64_t>(n_streams) + carve_size<uint32_t>(n_streams) +
                     carve_size<uint64_t>(n_streams) + carve_size<uint8_t>(hdr.size() + 1) + 16 * 256;
    size_t misc_total = 2 * carve_size<uint64_t>(n_streams) + 256;     // stream_end_bits | out_len
    CK(ctx->buf[NB_DESC].ensure(d_total + 256));
    CK(ctx->buf[NB_MISC].ensure(misc_total));
    CK(ctx->pin_meta.ensure(d_total + 256));
    uint8_t *hp = ctx->pin_meta.as<uint8_t>(), *hp0 = hp;
    uint8_t *dp = ctx->buf[NB_DESC].as<uint8_t>(), *dp0 = dp;
    auto put = [&](const void *src, size_t bytes) -> uint8_t * { memcpy(hp, src, bytes); uint8_t *d = dp0 + (hp - hp0); hp += align_up(bytes ? bytes : 1, 256); return d; };
    std::vector<uint32_t> hdr_len(n_streams, (uint32_t)hdr.size());
    EncDev E; memset(&E, 0, sizeof E);
    E.in = job.d_in; E.in_size = in_span;
    E.chunks = (const ChunkDesc *)put(P.chunks.data(), n_chunks * sizeof(ChunkDesc)); E.n_chunks = n_chunks;
    E.blocks = (const BlockDesc *)put(P.blocks.data(), n_blocks * sizeof(BlockDesc)); E.n_blocks = n_blocks;
    E.seg0 = (const uint32_t *)put(P.seg0.data(), (n_chunks + 1) * 4); E.pt0 = (const uint32_t *)put(P.pt0.data(), (n_chunks + 1) * 4);
    E.tile0 = (const uint32_t *)put(P.tile0.data(), (n_chunks + 1) * 4); E.grp0 = (const uint32_t *)put(P.grp0.data(), (n_chunks + 1) * 4);
    E.n_segs = P.seg0.back(); E.n_ptiles = P.pt0.back(); E.n_tiles = n_tiles; E.n_grps = P.grp0.back();
    E.window = o.window_size > 32768 ? 32768 : o.window_size; E.max_len = o.max_length > 258 ? 258 : o.max_length;
    E.n_streams = (uint32_t)n_streams;
    E.stream_blk0 = (const uint32_t *)put(P.stream_blk0.data(), (n_streams + 1) * 4);
    E.out_base = (const uint64_t *)put(job.out_base.data(), n_streams * 8);
    E.hdr_len = (const uint32_t *)put(hdr_len.data(), n_streams * 4);
    const uint64_t *d_in_len = (const uint64_t *)put(job.in_len.data(), n_streams * 8);
    const uint8_t *d_hdr = put(hdr.data(), hdr.size());
    CK(cudaMemcpyAsync(dp0, hp0, (size_t)(hp - hp0), cudaMemcpyHostToDevice, ctx->stream));
    E.link = ctx->buf[NB_LINK].as<uint16_t>(); E.md = ctx->buf[NB_MD].as<uint32_t>(); E.sym = ctx->buf[NB_SYM].as<uint32_t>();
    enc_set_fix(E, in_span);
    CK(cudaMemsetAsync(E.fix_count, 0, 256, ctx->stream));
    E.exit_tab = ctx->buf[NB_EXIT].as<uint16_t>();
    uint8_t *tp = ctx->buf[NB_TILE].as<uint8_t>();
    E.tile_entry = carve<uint16_t>(tp, n_tiles); E.tile_nsym = carve<uint32_t>(tp, n_tiles); E.tile_bits = carve<uint32_t>(tp, n_tiles); E.tile_bitrel = carve<uint64_t>(tp, n_tiles);
    uint8_t *bp = ctx->buf[NB_BLK].as<uint8_t>();
    E.hist = carve<uint32_t>(bp, (size_t)n_blocks * kHistStride);
    const size_t hist_bytes = (size_t)n_blocks * kHistStride * 4;
    E.litcode = carve<uint32_t>(bp, (size_t)n_blocks * kLitStride); E.distcode = carve<uint32_t>(bp, (size_t)n_blocks * kDistStride);
    E.hdr_words = carve<uint32_t>(bp, (size_t)n_blocks * kHdrWords); E.hdr_bits = carve<uint32_t>(bp, n_blocks);
    E.blk_bits = carve<uint64_t>(bp, n_blocks); E.blk_bitoff = carve<uint64_t>(bp, n_blocks); E.blk_markpos = carve<uint64_t>(bp, n_blocks);
    uint8_t *mp = ctx->buf[NB_MISC].as<uint8_t>();
    E.stream_end_bits = carve<uint64_t>(mp, n_streams);
    uint64_t *d_out_len = carve<uint64_t>(mp, n_streams);
    E.out_words = ctx->buf[NB_OUT].as<uint32_t>();

    ctx->tm.mark(ctx->stream, "clear");
    CK(cudaMemsetAsync(E.hist, 0, hist_bytes, ctx->stream));
    CK(cudaMemsetAsync(E.hdr_bits, 0xFF, (size_t)n_blocks * 4, ctx->stream));       // "codes not built yet" (k_huff_build)
    CK(cudaMemsetAsync(ctx->buf[NB_OUT].p, 0, out_total, ctx->stream));
    struct FeedCtx { b2f_ctx *ctx; const EncPlan *P; const EncodeJob *job; uint8_t *d_in; } fc = { ctx, &P, &job, const_cast<uint8_t *>(job.d_in) };
    SliceFeed feed = { &fc, [](void *self, uint32_t c0, uint32_t c1, cudaStream_t st) -> cudaError_t {
        FeedCtx *f = (FeedCtx *)self;
        if (c1 <= c0) return cudaSuccess;
        const uint64_t lo = f->P->chunks[c0].off, hi = f->P->chunks[c1 - 1].off + f->P->chunks[c1 - 1].len;   // device byte range of the slice
        for (size_t s = 0; s < f->job->in_off.size(); s++) {     // streams are laid out in increasing in_off order
            const uint64_t a = std::max<uint64_t>(lo, f->job->in_off[s]), b = std::min<uint64_t>(hi, f->job->in_off[s] + f->job->in_len[s]);
            if (a < b) { cudaError_t e = h2d_copy(f->ctx, f->d_in + a, f->job->h_in[s] + (a - f->job->in_off[s]), b - a, st, f->job->h_pinned[s] != 0); if (e != cudaSuccess) return e; }
        }
        return cudaSuccess; } };
    bool sliced = false;
    const bool early_out = job.h_out && n_streams == 1;      // the packed bytes of a slice are copied out while the next slices are matched
    ctx->pipe.n_slices = 0;
    ctx->pipe.h_pos = early_out ? ctx->pin_pos.as<uint64_t>() : nullptr;
    if (early_out && !hdr.empty()) CK(cudaMemcpyAsync(ctx->buf[NB_OUT].as<uint8_t>() + job.out_base[0], d_hdr, hdr.size(), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(enc_launch_lz(E, P.seg0.data(), P.pt0.data(), P.tile0.data(), P.grp0.data(), ctx->stream, &ctx->tm, ctx->aux, ctx->aux_ev, ctx->overlap ? ctx->enc_slices : 0u,
                     job.h_in ? &feed : nullptr, P.chunks.data(), &sliced, ctx->overlap ? &ctx->pipe : nullptr));
    const bool piped = sliced && ctx->pipe.n_slices > 0;
    if (n_chunks) ctx->stats.kernel_launches += enc_launch_count_lz(sliced) * (sliced ? ctx->enc_slices : 1);
    // checksums over the inputs (C1/C2) -> trailers
    const bool ck_async = ctx->overlap && ctx->aux[0] != nullptr;
    const uint32_t *d_crc = nullptr, *d_adler = nullptr;
    if (fmt == B2F_FMT_GZIP || fmt == B2F_FMT_ZLIB) {
        std::vector<uint64_t> piece0(n_streams + 1, 0); uint32_t span_rows = 0;
        checksum_plan(job.d_in, job.in_off.data(), job.in_len.data(), n_streams, piece0.data(), &span_rows);
        size_t n = n_streams;
        size_t o_off = 0, o_len = o_off + n * 8, o_p0 = o_len + n * 8, o_a = o_p0 + (n + 1) * 8, o_b = o_a + n * 8, o_c = o_b + n * 8,
               o_oc = o_c + n * 4, o_oa = o_oc + n * 4, total = align_up(o_oa + n * 4, 16);
        CK(ctx->buf[NB_CK].ensure(total));
        CK(ctx->pin_ck.ensure(total));
        uint8_t *hm = ctx->pin_ck.as<uint8_t>(); memset(hm, 0, total);
        memcpy(hm + o_off, job.in_off.data(), n * 8); memcpy(hm + o_len, job.in_len.data(), n * 8); memcpy(hm + o_p0, piece0.data(), (n + 1) * 8);
        uint8_t *dm = ctx->buf[NB_CK].as<uint8_t>();
        CK(cudaMemcpyAsync(dm, hm, total, cudaMemcpyHostToDevice, ctx->stream));
        ChecksumDev C; memset(&C, 0, sizeof C);
        C.in = job.d_in; C.off = (const uint64_t *)(dm + o_off); C.len = (const uint64_t *)(dm + o_len); C.span0 = (const uint64_t *)(dm + o_p0);
        C.n_spans = piece0[n]; C.span_rows = span_rows; C.n_streams = (uint32_t)n;
        C.acc_a = (uint64_t *)(dm + o_a); C.acc_b = (uint64_t *)(dm + o_b); C.acc_crc = (uint32_t *)(dm + o_c);
        C.out_crc = (uint32_t *)(dm + o_oc); C.out_adler = (uint32_t *)(dm + o_oa);
        // The checksum only needs the input: with overlap on it runs on an aux stream next to the entropy stage (huff_build, the scans
        // and write_headers are a few hundred threads each and leave the GPU idle); framing waits for it.
        if (ck_async) {
            CK(cudaEventRecord(ctx->aux_ev[0], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->aux[0], ctx->aux_ev[0], 0));
            CK(checksum_launch(C, fmt == B2F_FMT_GZIP, fmt == B2F_FMT_ZLIB, ctx->aux[0]));
            CK(cudaEventRecord(ctx->aux_ev[1], ctx->aux[0]));
        } else {
            ctx->tm.mark(ctx->stream, "checksum");
            CK(checksum_launch(C, fmt == B2F_FMT_GZIP, fmt == B2F_FMT_ZLIB, ctx->stream, &ctx->tm));
        }
        ctx->stats.kernel_launches += (C.n_spans ? 1 : 0) + 1;
        d_crc = C.out_crc; d_adler = C.out_adler;
    }
    if (!piped) {
        CK(enc_launch_entropy(E, ctx->stream, &ctx->tm, sliced));
        ctx->stats.kernel_launches += enc_launch_count_entropy(n_tiles != 0, sliced);
    } else ctx->stats.kernel_launches += 4 * ctx->pipe.n_slices;
    if (ck_async && (fmt == B2F_FMT_GZIP || fmt == B2F_FMT_ZLIB)) CK(cudaStreamWaitEvent(ctx->stream, ctx->aux_ev[1], 0));
    ctx->tm.mark(ctx->stream, "framing");
    k_write_framing<<<(unsigned)((n_streams + 63) / 64), 64, 0, ctx->stream>>>(ctx->buf[NB_OUT].as<uint8_t>(), E.out_base, E.stream_end_bits, d_hdr,
                                                                             (uint32_t)hdr.size(), fmt, d_crc, d_adler, d_in_len, d_out_len, (uint32_t)n_streams);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    ctx->tm.finish(ctx->stream);
    CK(ctx->pin_res.ensure(n_streams * 16 + 64));
    uint64_t *h_out_len = ctx->pin_res.as<uint64_t>();
    CK(cudaMemcpyAsync(h_out_len, d_out_len, n_streams * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(h_out_len + n_streams, E.stream_end_bits, n_streams * 8, cudaMemcpyDeviceToHost, ctx->stream));
    job.h_copied = 0;
    if (piped && early_out) {
        // Everything is queued; now follow the slices: as soon as the bit position after slice k is known, the bytes below it (minus the
        // byte shared with the next slice) are final once slice k is packed -- copy them out on a side stream while later slices run.
        const uint64_t *h_pos = ctx->pin_pos.as<uint64_t>();
        const uint8_t *d_stream = ctx->buf[NB_OUT].as<uint8_t>() + job.out_base[0];
        for (uint32_t k = 0; k + 1 < ctx->pipe.n_slices; k++) {          // (the last slice leaves with the trailer, below)
            CK(cudaEventSynchronize(ctx->pipe.ev_scan[k]));
            const uint64_t upto = std::min<uint64_t>(h_pos[k] >> 3, job.h_out_cap);
            CK(cudaStreamWaitEvent(ctx->copy_st, ctx->pipe.ev_pack[k], 0));
            if (upto > job.h_copied) {
                CK(d2h_copy(ctx, job.h_out + job.h_copied, d_stream + job.h_copied, upto - job.h_copied, ctx->copy_st, job.h_out_pinned));
                job.h_copied = upto;
            }
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (getenv("B2F_DEBUG_FIX")) {                    // developer aid: how many positions k_lz_find handed to the fix-up levels
        uint32_t fc[2 * kFixSlices];
        cudaMemcpy(fc, E.fix_count, sizeof fc, cudaMemcpyDeviceToHost);
        uint64_t a = 0, b = 0;
        for (uint32_t i = 0; i < kFixSlices; i++) { a += fc[i]; b += fc[kFixSlices + i]; }
        fprintf(stderr, "[b2f] deferred positions: level 1 %llu, level 2 %llu (of %llu)\n", (unsigned long long)a, (unsigned long long)b, (unsigned long long)in_span);
    }
    job.out_bits.resize(n_streams);
    for (size_t s = 0; s < n_streams; s++) { job.out_len[s] = h_out_len[s]; job.out_bits[s] = h_out_len[n_streams + s] - 8ull * hdr.size(); }
    return B2F_OK;
}

// host arrays packed into one pinned staging block and one device block (one H2D copy)
struct Packer {
    PinBuf &pin; DevBuf &dev; size_t off = 0; std::vector<std::pair<const void *, size_t>> items; std::vector<size_t> offs;
    Packer(PinBuf &p, DevBuf &d) : pin(p), dev(d) {}
    size_t add(const void *src, size_t bytes) { size_t o = off; items.push_back({ src, bytes }); offs.push_back(o); off += align_up(bytes ? bytes : 1, 256); return o; }
    size_t reserve(size_t bytes) { size_t o = off; off += align_up(bytes ? bytes : 1, 256); return o; }     // device-only region
    cudaError_t commit(cudaStream_t st) {
        cudaError_t e = dev.ensure(off + 256); if (e != cudaSuccess) return e;
        e = pin.ensure(off + 256); if (e != cudaSuccess) return e;
        size_t hi = 0;
        for (size_t i = 0; i < items.size(); i++) { if (items[i].second) memcpy(pin.as<uint8_t>() + offs[i], items[i].first, items[i].second); hi = std::max(hi, offs[i] + items[i].second); }
        if (hi) e = cudaMemcpyAsync(dev.p, pin.p, hi, cudaMemcpyHostToDevice, st);
        return e;
    }
    template <class T> T *ptr(size_t o) const { return reinterpret_cast<T *>(dev.as<uint8_t>() + o); }
};

// Stored mode (EncodeOptions::no_compression, RawBuf encode.rs:354-383): no compute besides the checksum, but the stream is still
// assembled on the device -- the host only lays it out.  A stream is a list of copy records: literal bytes (container header, the
// five header bytes of every stored block, sync markers, trailer; gathered in `blob`) and payload ranges of the input.
struct CopyRec { uint64_t src; uint64_t dst; uint64_t n; };     // src: offset into the blob (bit 63 set) or into the input buffer
constexpr uint64_t kRecBlob = 1ull << 63;
uint64_t stored_layout(int fmt, const b2f_encode_opts &o, uint64_t in_base, size_t n, const int64_t *sched, size_t n_sched,
                       uint32_t crc, uint32_t adler, uint64_t out_base, std::vector<uint8_t> &blob, std::vector<CopyRec> &recs) {
    uint64_t out = 0;                                  // stream-relative output position
    auto lit = [&](const uint8_t *b, size_t k) { recs.push_back({ kRecBlob | blob.size(), out_base + out, k }); blob.insert(blob.end(), b, b + k); out += k; };
    std::vector<uint8_t> hdr; make_header(fmt, o, hdr);
    if (!hdr.empty()) lit(hdr.data(), hdr.size());
    size_t bs = (size_t)o.block_size; if (bs > 0xFFFF) bs = 0xFFFF;
    size_t buf_start = 0, pos = 0;                     // RawBuf holds in[buf_start, pos)
    auto flush = [&](bool fin) {
        const size_t size = std::min<size_t>(pos - buf_start, 0xFFFF);
        const uint16_t ns = (uint16_t)~size;
        const uint8_t h[5] = { (uint8_t)(fin ? 1 : 0), (uint8_t)size, (uint8_t)(size >> 8), (uint8_t)ns, (uint8_t)(ns >> 8) };   // BFINAL + BTYPE=00, byte-aligned
        lit(h, 5);
        if (size) { recs.push_back({ in_base + buf_start, out_base + out, size }); out += size; }
        buf_start += size;
    };
    int64_t one = (int64_t)n;
    if (!sched) { sched = &one; n_sched = n ? 1 : 0; }
    for (size_t k = 0; k < n_sched; k++) {
        if (sched[k] < 0) {
            flush(false);
            if (fmt == B2F_FMT_ZLIB && o.zlib_flush_sync) { const uint8_t m[5] = { 0, 0, 0, 255, 255 }; lit(m, 5); }
            continue;
        }
        size_t w = (size_t)sched[k]; if (pos + w > n) w = n - pos;
        pos += w;
        while (pos - buf_start >= bs) flush(false);
    }
    flush(true);
    if (fmt == B2F_FMT_GZIP) { uint8_t t[8]; for (int k = 0; k < 4; k++) { t[k] = (uint8_t)(crc >> (8 * k)); t[4 + k] = (uint8_t)((uint32_t)n >> (8 * k)); } lit(t, 8); }
    else if (fmt == B2F_FMT_ZLIB) { uint8_t t[4]; for (int k = 0; k < 4; k++) t[k] = (uint8_t)(adler >> (8 * (3 - k))); lit(t, 4); }
    return out;
}
__global__ void __launch_bounds__(256) k_copy_records(const CopyRec *__restrict__ recs, const uint8_t *__restrict__ in, const uint8_t *__restrict__ blob,
                                                     uint8_t *__restrict__ out) {
    const CopyRec r = recs[blockIdx.x];
    const uint8_t *__restrict__ src = (r.src & kRecBlob) ? blob + (r.src & ~kRecBlob) : in + r.src;
    uint8_t *__restrict__ dst = out + r.dst;
    for (uint64_t i = threadIdx.x; i < r.n; i += 256) dst[i] = src[i];
}
// Stored-mode encode of a batch whose inputs are in d_in: checksums, layout, one copy kernel.  d_dst = where stream s goes
// (dst_off[s]); streams whose output does not fit get OUTPUT_TOO_SMALL and are not written.
int stored_on_device(b2f_ctx *ctx, int fmt, const b2f_encode_opts &o, size_t n_streams, const uint8_t *d_in, const std::vector<uint64_t> &in_off,
                     const std::vector<uint64_t> &in_len, const int64_t *const *sched, const size_t *n_sched,
                     uint8_t *d_dst, const std::vector<uint64_t> &dst_off, const size_t *out_cap, size_t *out_len, int *status) {
    std::vector<uint32_t> crc, adler;
    int rc = run_checksums(ctx, d_in, in_off, in_len, fmt == B2F_FMT_GZIP, fmt == B2F_FMT_ZLIB, nullptr, crc, adler);
    if (rc) return rc;
    std::vector<uint8_t> blob; std::vector<CopyRec> recs;
    for (size_t s = 0; s < n_streams; s++) {
        const size_t r0 = recs.size(), b0 = blob.size();
        out_len[s] = (size_t)stored_layout(fmt, o, in_off[s], (size_t)in_len[s], sched ? sched[s] : nullptr, (sched && n_sched) ? n_sched[s] : 0,
                                           crc[s], adler[s], dst_off[s], blob, recs);
        if (out_len[s] > out_cap[s]) { status[s] = B2F_ERR_OUTPUT_TOO_SMALL; recs.resize(r0); blob.resize(b0); } else status[s] = B2F_OK;
    }
    if (!recs.empty()) {
        Packer P(ctx->pin_meta, ctx->buf[NB_DEC_META]);
        const size_t a_r = P.add(recs.data(), recs.size() * sizeof(CopyRec)), a_b = P.add(blob.data(), blob.size());
        CK(P.commit(ctx->stream));
        ctx->tm.mark(ctx->stream, "stored_copy");
        k_copy_records<<<(unsigned)recs.size(), 256, 0, ctx->stream>>>(P.ptr<CopyRec>(a_r), d_in, P.ptr<uint8_t>(a_b), d_dst);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches += 1;
    }
    return B2F_OK;
}
}  // namespace

// bytes of in[] that the schedule actually writes (a schedule may stop short of in_len; those bytes are never written)
static size_t effective_len(const int64_t *sched, size_t n_sched, size_t in_len) {
    if (!sched) return in_len;
    size_t pos = 0;
    for (size_t k = 0; k < n_sched && pos < in_len; k++) if (sched[k] > 0) pos += std::min<size_t>((size_t)sched[k], in_len - pos);
    return pos;
}

static int validate_opts(b2f_ctx *ctx, int fmt, const b2f_encode_opts &o) {
    if (fmt < B2F_FMT_DEFLATE || fmt > B2F_FMT_GZIP) { ctx->err = "bad format"; return B2F_ERR_INVALID_ARG; }
    if (o.block_size == 0 || o.window_size == 0 || o.max_length < 3 || o.mode < 0 || o.mode > 2) { ctx->err = "bad encode options"; return B2F_ERR_INVALID_ARG; }
    // gzip::ExtraField::write_to fails with "extra field too long" beyond a 16-bit XLEN (src/gzip.rs:489-500)
    if (o.gzip_has_extra && (o.gzip_extra_len > 0xFFFFu || (o.gzip_extra_len && !o.gzip_extra))) { ctx->err = "gzip extra field: NULL or longer than 65535 bytes"; return B2F_ERR_INVALID_ARG; }
    return B2F_OK;
}

extern "C" int b2f_encode_device(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, size_t n_streams,
                                 const uint8_t *d_in, const uint64_t *in_off, const size_t *in_len,
                                 const int64_t *const *sched, const size_t *n_sched,
                                 uint8_t *d_out, const uint64_t *out_off, const size_t *out_cap, size_t *out_len, int *status) {
    if (!ctx) return B2F_ERR_INVALID_ARG;
    if (n_streams && (!d_in || !in_off || !in_len || !d_out || !out_off || !out_cap || !out_len || !status)) { ctx->err = "NULL argument"; return B2F_ERR_INVALID_ARG; }
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    int rc = validate_opts(ctx, fmt, *opts); if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    if (opts->mode == B2F_MODE_STORED) {
        std::vector<uint64_t> io(in_off, in_off + n_streams), il(n_streams), oo(out_off, out_off + n_streams);
        for (size_t s = 0; s < n_streams; s++) il[s] = effective_len(sched ? sched[s] : nullptr, (sched && n_sched) ? n_sched[s] : 0, in_len[s]);
        rc = stored_on_device(ctx, fmt, *opts, n_streams, d_in, io, il, sched, n_sched, d_out, oo, out_cap, out_len, status);
        if (rc) return rc;
        ctx->tm.finish(ctx->stream);
        CK(cudaStreamSynchronize(ctx->stream));
        collect_stats(ctx, false);
        return B2F_OK;
    }
    EncodeJob job; job.d_in = d_in;
    job.in_off.assign(in_off, in_off + n_streams); job.in_len.resize(n_streams);
    for (size_t s = 0; s < n_streams; s++) job.in_len[s] = effective_len(sched ? sched[s] : nullptr, (sched && n_sched) ? n_sched[s] : 0, in_len[s]);
    rc = encode_on_device(ctx, fmt, *opts, n_streams, sched, n_sched, job);
    if (rc) return rc;
    for (size_t s = 0; s < n_streams; s++) {
        out_len[s] = (size_t)job.out_len[s];
        if (job.out_len[s] > out_cap[s]) { status[s] = B2F_ERR_OUTPUT_TOO_SMALL; continue; }
        status[s] = B2F_OK;
        CK(cudaMemcpyAsync(d_out + out_off[s], ctx->buf[NB_OUT].as<uint8_t>() + job.out_base[s], job.out_len[s], cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    collect_stats(ctx, false);
    return B2F_OK;
}

extern "C" int b2f_encode_batch(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, size_t n_streams,
                                const uint8_t *const *in, const size_t *in_len,
                                const int64_t *const *sched, const size_t *n_sched,
                                uint8_t *const *out, const size_t *out_cap, size_t *out_len, int *status) {
    if (!ctx) return B2F_ERR_INVALID_ARG;
    if (n_streams && (!in || !in_len || !out || !out_cap || !out_len || !status)) { ctx->err = "NULL argument"; return B2F_ERR_INVALID_ARG; }
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    int rc = validate_opts(ctx, fmt, *opts); if (rc) return rc;
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    // stage inputs: stream s at a 256-aligned offset
    EncodeJob job; job.in_off.resize(n_streams); job.in_len.resize(n_streams);
    uint64_t total = 0;
    for (size_t s = 0; s < n_streams; s++) {
        job.in_off[s] = total;
        job.in_len[s] = effective_len(sched ? sched[s] : nullptr, (sched && n_sched) ? n_sched[s] : 0, in_len[s]);
        total += align_up(job.in_len[s], 256);
    }
    CK(ctx->buf[NB_IN].ensure(total + 512));
    uint8_t *d_in = ctx->buf[NB_IN].as<uint8_t>();
    job.d_in = d_in;
    if (opts->mode == B2F_MODE_STORED) {
        ctx->tm.mark(ctx->stream, "h2d");
        for (size_t s = 0; s < n_streams; s++) if (job.in_len[s]) CK(h2d_copy(ctx, d_in + job.in_off[s], in[s], job.in_len[s], ctx->stream, is_pinned_host(in[s])));
        // the streams are assembled in the device output buffer (bound: 5 bytes per 65535 + container) and copied out
        std::vector<uint64_t> dst_off(n_streams); uint64_t tout = 0;
        for (size_t s = 0; s < n_streams; s++) { dst_off[s] = tout; tout += align_up(b2f_encode_bound((size_t)job.in_len[s], (sched && n_sched) ? n_sched[s] : 0, opts) + 16, 256); }
        CK(ctx->buf[NB_OUT].ensure(tout + 512));
        std::vector<size_t> cap(out_cap, out_cap + n_streams);
        rc = stored_on_device(ctx, fmt, *opts, n_streams, d_in, job.in_off, job.in_len, sched, n_sched, ctx->buf[NB_OUT].as<uint8_t>(), dst_off, cap.data(), out_len, status);
        if (rc) return rc;
        ctx->tm.finish(ctx->stream);
        for (size_t s = 0; s < n_streams; s++)
            if (status[s] == B2F_OK) CK(d2h_copy(ctx, out[s], ctx->buf[NB_OUT].as<uint8_t>() + dst_off[s], out_len[s], ctx->stream, is_pinned_host(out[s])));
        CK(cudaStreamSynchronize(ctx->stream));
        collect_stats(ctx, false);
        return B2F_OK;
    }
    job.h_in = in;                         // the H2D copies are issued per slice inside the LZ77 stage
    job.h_pinned.resize(n_streams);
    for (size_t s = 0; s < n_streams; s++) job.h_pinned[s] = is_pinned_host(in[s]) ? 1 : 0;
    if (n_streams == 1 && out[0]) { job.h_out = out[0]; job.h_out_pinned = is_pinned_host(out[0]); job.h_out_cap = out_cap[0]; }
    rc = encode_on_device(ctx, fmt, *opts, n_streams, sched, n_sched, job);
    if (rc) return rc;
    collect_stats(ctx, false);
    for (size_t s = 0; s < n_streams; s++) {
        out_len[s] = (size_t)job.out_len[s];
        if (job.out_len[s] > out_cap[s]) { status[s] = B2F_ERR_OUTPUT_TOO_SMALL; continue; }
        status[s] = B2F_OK;
        const size_t have = n_streams == 1 ? (size_t)std::min<uint64_t>(job.h_copied, job.out_len[s]) : 0;     // copied while the encode was still running
        CK(d2h_copy(ctx, out[s] + have, ctx->buf[NB_OUT].as<uint8_t>() + job.out_base[s] + have, job.out_len[s] - have, ctx->stream, is_pinned_host(out[s])));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_st));
    return B2F_OK;
}

// ------------------------------------------------------------------------------------------ one stream split into parts
// The blocks of a DEFLATE stream are independent in libflate's output (fresh LZ77 table per chunk, own Huffman codes per block),
// so contiguous runs of blocks can be encoded by different contexts / GPUs.  What has to be exchanged afterwards is 8 bytes per
// part: its length in BITS (blocks are not byte aligned, src/deflate/encode.rs:291-293) -- an exclusive scan of those gives every
// part's position -- plus its checksum, folded with the algebraic combine below.  Payload bytes never cross GPUs.
namespace {
__global__ void k_shift_copy(uint8_t *dst, const uint8_t *src, uint64_t nbytes_in, uint32_t shift) {
    // dst bit (i + shift) = src bit i, LSB first; dst has nbytes_in + 1 bytes (bits below `shift` of dst[0] are zero)
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nbytes_in) return;
    const uint32_t lo = i ? src[i - 1] : 0u, hi = i < nbytes_in ? src[i] : 0u;
    dst[i] = (uint8_t)(((hi << 8 | lo) << shift) >> 8);
}
uint32_t crc_mulmod(uint32_t a, uint32_t b) {                  // product mod the reflected CRC-32 polynomial
    uint32_t p = 0;
    for (int i = 31; i >= 0; i--) { if ((a >> i) & 1u) p ^= b; b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1; }
    return p;
}
}  // namespace

extern "C" uint32_t b2f_crc32_combine(uint32_t crc1, uint32_t crc2, uint64_t len2) {
    // crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B)   (what checksum::Crc32 would give after update(A), update(B))
    uint32_t xp = 1u << 31, sq = 1u << 30;                       // x^0, x^1
    for (uint64_t e = len2 * 8; e; e >>= 1) { if (e & 1) xp = crc_mulmod(sq, xp); sq = crc_mulmod(sq, sq); }
    return crc_mulmod(xp, crc1) ^ crc2;
}
extern "C" uint32_t b2f_adler32_combine(uint32_t adler1, uint32_t adler2, uint64_t len2) {
    const uint64_t M = 65521, a1 = adler1 & 0xFFFF, b1 = adler1 >> 16, a2 = adler2 & 0xFFFF, b2 = adler2 >> 16, r = len2 % M;
    const uint64_t a = (a1 + a2 + M - 1) % M;
    const uint64_t b = (b1 + b2 + r * ((a1 + M - 1) % M)) % M;
    return (uint32_t)((b << 16) | a);
}
extern "C" size_t b2f_stream_header(int fmt, const b2f_encode_opts *opts, uint8_t *out, size_t cap) {
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    std::vector<uint8_t> h; make_header(fmt, *opts, h);
    if (out && cap >= h.size()) memcpy(out, h.data(), h.size());
    return h.size();
}
extern "C" size_t b2f_stream_trailer(int fmt, uint32_t crc32, uint32_t adler32, uint64_t total_len, uint8_t *out) {
    if (fmt == B2F_FMT_GZIP) { for (int k = 0; k < 4; k++) { out[k] = (uint8_t)(crc32 >> (8 * k)); out[4 + k] = (uint8_t)((uint32_t)total_len >> (8 * k)); } return 8; }
    if (fmt == B2F_FMT_ZLIB) { for (int k = 0; k < 4; k++) out[k] = (uint8_t)(adler32 >> (8 * (3 - k))); return 4; }
    return 0;
}
extern "C" int b2f_encode_part_device(b2f_ctx *ctx, const b2f_encode_opts *opts, const uint8_t *d_in, size_t in_len, const int64_t *sched, size_t n_sched,
                                      int is_last, uint8_t *d_out, size_t out_cap, uint64_t *out_bits, uint32_t *crc32, uint32_t *adler32) {
    if (!ctx || !d_in || !d_out || !out_bits) return B2F_ERR_INVALID_ARG;
    b2f_encode_opts d; if (!opts) { b2f_encode_opts_default(&d); opts = &d; }
    int rc = validate_opts(ctx, B2F_FMT_DEFLATE, *opts); if (rc) return rc;
    if (opts->mode == B2F_MODE_STORED) { ctx->err = "stored mode has no part variant"; return B2F_ERR_INVALID_ARG; }
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    EncodeJob job; job.d_in = d_in; job.part = true; job.part_last = is_last != 0;
    job.in_off.assign(1, 0); job.in_len.assign(1, effective_len(sched, n_sched, in_len));
    const int64_t *sp = sched; size_t sn = n_sched;
    rc = encode_on_device(ctx, B2F_FMT_DEFLATE, *opts, 1, sched ? &sp : nullptr, sched ? &sn : nullptr, job);
    if (rc) return rc;
    *out_bits = job.out_bits[0];
    if (job.out_len[0] + 1 > out_cap) return B2F_ERR_OUTPUT_TOO_SMALL;
    CK(cudaMemcpyAsync(d_out, ctx->buf[NB_OUT].as<uint8_t>() + job.out_base[0], job.out_len[0], cudaMemcpyDeviceToDevice, ctx->stream));
    if (crc32 || adler32) {
        std::vector<uint32_t> c, a;
        rc = run_checksums(ctx, d_in, job.in_off, job.in_len, true, true, nullptr, c, a);
        if (rc) return rc;
        if (crc32) *crc32 = c[0];
        if (adler32) *adler32 = a[0];
    }
    ctx->tm.finish(ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    collect_stats(ctx, false);
    return B2F_OK;
}
extern "C" int b2f_bits_shift_device(b2f_ctx *ctx, const uint8_t *d_src, uint64_t n_bits, uint32_t shift, uint8_t *d_dst) {
    if (!ctx || !d_src || !d_dst || shift > 7) return B2F_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    const uint64_t nb = (n_bits + 7) / 8;
    k_shift_copy<<<(unsigned)((nb + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_dst, d_src, nb, shift);
    CK(cudaGetLastError());
    ctx->stats.kernel_launches += 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return B2F_OK;
}

// ------------------------------------------------------------------------------------------ E2: Lz77Encode backend
extern "C" int b2f_lz77_default(b2f_ctx *ctx, const uint8_t *buf, size_t len, uint32_t window_size, uint32_t max_length, uint32_t *codes, size_t *n_codes) {
    if (!ctx || !n_codes || (len && (!buf || !codes))) return B2F_ERR_INVALID_ARG;
    if (window_size == 0 || max_length < 3) return B2F_ERR_INVALID_ARG;
    *n_codes = 0;
    if (len == 0) return B2F_OK;
    if (len >= (1ull << 31)) return B2F_ERR_INVALID_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    CK(ctx->buf[NB_IN].ensure(len + 512));
    uint8_t *d_in = ctx->buf[NB_IN].as<uint8_t>();
    CK(cudaMemcpyAsync(d_in, buf, len, cudaMemcpyHostToDevice, ctx->stream));
    const uint32_t n_tiles = (uint32_t)((len + kTile - 1) / kTile);
    ChunkDesc cd; cd.off = 0; cd.len = (uint32_t)len; cd.block = 0;
    uint32_t pref[8] = { 0, (uint32_t)((len + kSeg - 1) / kSeg), 0, (uint32_t)((len + kPTile - 1) / kPTile), 0, n_tiles, 0, (n_tiles + kGrpTiles - 1) / kGrpTiles };
    CK(ctx->buf[NB_LINK].ensure(enc_link_bytes(len))); CK(ctx->buf[NB_MD].ensure(len * 4 + 256)); CK(ctx->buf[NB_SYM].ensure(len * 4 + 256));
    CK(ctx->buf[NB_EXIT].ensure((size_t)n_tiles * kExitW * 2 + 256));
    size_t tile_bytes = carve_size<uint16_t>(n_tiles) + carve_size<uint32_t>(n_tiles) + carve_size<uint64_t>(n_tiles) + 512;
    CK(ctx->buf[NB_TILE].ensure(tile_bytes));
    CK(ctx->buf[NB_BLK].ensure(kHistStride * 4 + 256));
    CK(ctx->buf[NB_DESC].ensure(1024));
    CK(ctx->buf[NB_OUT].ensure(len * 4 + 256));          // compacted codes
    CK(ctx->pin_meta.ensure(1024));
    uint8_t *hp = ctx->pin_meta.as<uint8_t>();
    memcpy(hp, &cd, sizeof cd); memcpy(hp + 256, pref, sizeof pref);
    uint8_t *dp = ctx->buf[NB_DESC].as<uint8_t>();
    CK(cudaMemcpyAsync(dp, hp, 512, cudaMemcpyHostToDevice, ctx->stream));
    EncDev E; memset(&E, 0, sizeof E);
    E.in = d_in; E.in_size = len; E.chunks = (const ChunkDesc *)dp; E.n_chunks = 1;
    const uint32_t *dpref = (const uint32_t *)(dp + 256);
    E.seg0 = dpref; E.pt0 = dpref + 2; E.tile0 = dpref + 4; E.grp0 = dpref + 6;
    E.n_segs = pref[1]; E.n_ptiles = pref[3]; E.n_tiles = n_tiles; E.n_grps = pref[7];
    E.window = window_size > 32768 ? 32768 : window_size; E.max_len = max_length > 258 ? 258 : max_length;
    E.link = ctx->buf[NB_LINK].as<uint16_t>(); E.md = ctx->buf[NB_MD].as<uint32_t>(); E.sym = ctx->buf[NB_SYM].as<uint32_t>();
    enc_set_fix(E, len);
    CK(cudaMemsetAsync(E.fix_count, 0, 256, ctx->stream));
    E.exit_tab = ctx->buf[NB_EXIT].as<uint16_t>();
    uint8_t *tp = ctx->buf[NB_TILE].as<uint8_t>();
    E.tile_entry = carve<uint16_t>(tp, n_tiles); E.tile_nsym = carve<uint32_t>(tp, n_tiles);
    uint64_t *tile_symoff = carve<uint64_t>(tp, n_tiles);
    uint64_t *d_total = reinterpret_cast<uint64_t *>(tp);
    E.hist = ctx->buf[NB_BLK].as<uint32_t>();
    CK(cudaMemsetAsync(E.hist, 0, kHistStride * 4, ctx->stream));
    CK(enc_launch_lz(E, pref, pref + 2, pref + 4, pref + 6, ctx->stream, &ctx->tm, nullptr, nullptr, 0, nullptr, nullptr, nullptr));
    uint32_t *d_codes = ctx->buf[NB_OUT].as<uint32_t>();
    CK(enc_launch_compact(E, tile_symoff, d_total, d_codes, ctx->stream));
    ctx->stats.kernel_launches += enc_launch_count_lz(false) + 2;
    ctx->tm.finish(ctx->stream);
    CK(ctx->pin_res.ensure(64));
    uint64_t *h_total = ctx->pin_res.as<uint64_t>();
    CK(cudaMemcpyAsync(h_total, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_codes = (size_t)*h_total;
    CK(cudaMemcpyAsync(codes, d_codes, *n_codes * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    collect_stats(ctx, false);
    return B2F_OK;
}

// ------------------------------------------------------------------------------------------ decode
namespace {
struct HostReader { const uint8_t *p; size_t n, pos; };
bool rd_exact(HostReader &r, uint8_t *dst, size_t k) { if (r.n - r.pos < k) { r.pos = r.n; return false; } memcpy(dst, r.p + r.pos, k); r.pos += k; return true; }

// gzip::Header::read_from (gzip.rs:390-446) incl. libflate's CRC16 convention; returns B2F status
int parse_gzip_header(HostReader &r) {
    uint8_t b[10];
    if (!rd_exact(r, b, 10)) return B2F_ERR_UNEXPECTED_EOF;
    if (b[0] != 31 || b[1] != 139) return B2F_ERR_INVALID_DATA;
    if (b[2] != 8) return B2F_ERR_INVALID_DATA;
    const uint8_t flags = b[3];
    std::vector<uint8_t> re;                     // the header as libflate re-serialises it for the CRC16 (is_text is never restored)
    const uint8_t xfl = b[8] == 4 ? 4 : b[8] == 2 ? 2 : 0;
    const uint8_t h[10] = { 31, 139, 8, (uint8_t)(flags & (4 | 8 | 16)), b[4], b[5], b[6], b[7], xfl, b[9] };
    re.insert(re.end(), h, h + 10);
    if (flags & 4) {
        uint8_t l2[2];
        if (!rd_exact(r, l2, 2)) return B2F_ERR_UNEXPECTED_EOF;
        size_t limit = (size_t)(l2[0] | (l2[1] << 8));
        size_t total_pos = re.size(); re.push_back(0); re.push_back(0); size_t total = 0;
        while (limit > 0) {
            uint8_t sf[4];
            if (limit < 4 || r.n - r.pos < 4) { r.pos = std::min(r.n, r.pos + limit); return B2F_ERR_UNEXPECTED_EOF; }
            rd_exact(r, sf, 4); limit -= 4;
            size_t dl = (size_t)(sf[2] | (sf[3] << 8));
            if (dl > limit || r.n - r.pos < dl) { r.pos = std::min(r.n, r.pos + std::min(dl, limit)); return B2F_ERR_UNEXPECTED_EOF; }
            re.insert(re.end(), sf, sf + 4); re.insert(re.end(), r.p + r.pos, r.p + r.pos + dl); r.pos += dl; limit -= dl; total += 4 + dl;
        }
        re[total_pos] = (uint8_t)total; re[total_pos + 1] = (uint8_t)(total >> 8);
    }
    for (int f = 8; f <= 16; f <<= 1) if (flags & f) {
        for (;;) { uint8_t c; if (!rd_exact(r, &c, 1)) return B2F_ERR_UNEXPECTED_EOF; re.push_back(c); if (!c) break; }
    }
    if (flags & 2) {
        uint8_t c2[2];
        if (!rd_exact(r, c2, 2)) return B2F_ERR_UNEXPECTED_EOF;
        uint16_t crc = (uint16_t)(c2[0] | (c2[1] << 8)), expected = (uint16_t)host_crc32(re.data(), re.size());
        if (crc != expected) return B2F_ERR_INVALID_DATA;
    }
    return B2F_OK;
}
int parse_zlib_header(HostReader &r) {          // zlib::Header::read_from (zlib.rs:221-266)
    uint8_t b[2];
    if (!rd_exact(r, b, 2)) return B2F_ERR_UNEXPECTED_EOF;
    if (((((uint32_t)b[0]) << 8) + b[1]) % 31 != 0) return B2F_ERR_INVALID_DATA;
    if ((b[0] & 15) != 8) return B2F_ERR_INVALID_DATA;
    if ((b[0] >> 4) > 7) return B2F_ERR_INVALID_DATA;
    if (b[1] & 0x20) { uint8_t d[4]; if (!rd_exact(r, d, 4)) return B2F_ERR_UNEXPECTED_EOF; return B2F_ERR_INVALID_DATA; }
    return B2F_OK;
}
int map_inf_status(int s) { return s == kInfOk ? B2F_OK : s == kInfInvalid ? B2F_ERR_INVALID_DATA : s == kInfEof ? B2F_ERR_UNEXPECTED_EOF : B2F_ERR_OUTPUT_TOO_SMALL; }

struct Member { size_t stream; uint64_t def_off; uint64_t def_len; uint64_t out_off; uint64_t out_cap; uint8_t *h_out; bool h_pinned;      // h_out: host destination of this member's output (or NULL); h_pinned: page-locked
                uint64_t scan_len; };   // bytes of def_len the block finder / speculative parse may look at (0: in-order kernel)

// Packs several host arrays into one pinned staging area + one H2D copy; returns device pointers.

constexpr uint64_t kParallelMinBytes = 128 * 1024;    // smaller streams are decoded in order by one warp each

// Decodes one "round": members[] are raw DEFLATE streams inside d_in; outputs to d_out.  Results to host vectors.
//  - large members: block-boundary finder -> speculative sub-block parse of every candidate block -> verify the chain of
//    block ends from bit 0 -> tokens -> independent LZ77 units resolved in parallel (spec_kernels.cu);
//  - small members, and any member whose chain is not clean (non-dynamic blocks, cross-block back-references,
//    errors, too-small output): in-order kernel, which reproduces libflate's error kinds and partial output.
int inflate_round(b2f_ctx *ctx, const uint8_t *d_in, uint8_t *d_out, const std::vector<Member> &mem,
                  std::vector<int> &st, std::vector<uint64_t> &olen, std::vector<uint64_t> &cons, std::vector<char> &copied, std::vector<uint64_t> &good) {
    const size_t n = mem.size();
    st.assign(n, 0); olen.assign(n, 0); cons.assign(n, 0); copied.assign(n, 0); good.assign(n, 0);
    if (!n) return B2F_OK;
    // in_len: what the speculative path may look at (the whole rest of the container for a first member, a window sized after the
    // previous member for later members of a multi-member file); full_len: what the in-order kernel may read
    std::vector<uint64_t> in_off(n), in_len(n), full_len(n), out_off(n), out_cap(n), out_end(n);
    for (size_t i = 0; i < n; i++) { in_off[i] = mem[i].def_off; in_len[i] = std::min(mem[i].scan_len, mem[i].def_len); full_len[i] = mem[i].def_len; out_off[i] = mem[i].out_off; out_cap[i] = mem[i].out_cap; out_end[i] = mem[i].out_off + mem[i].out_cap; }
    std::vector<uint32_t> big;
    for (size_t i = 0; i < n; i++) if (in_len[i] >= kParallelMinBytes) big.push_back((uint32_t)i);
    std::vector<uint32_t> serial;                          // member indices for the in-order kernel
    std::vector<char> is_par(n, 0);
    // ---- phase A: members + finder
    Packer PA(ctx->pin_meta, ctx->buf[NB_DEC_META]);
    const size_t a_io = PA.add(in_off.data(), n * 8), a_il = PA.add(in_len.data(), n * 8), a_oo = PA.add(out_off.data(), n * 8),
                 a_oc = PA.add(out_cap.data(), n * 8), a_oe = PA.add(out_end.data(), n * 8);
    std::vector<uint32_t> seg0(big.size() + 1, 0);
    uint64_t big_bytes = 0;
    for (size_t k = 0; k < big.size(); k++) { seg0[k + 1] = seg0[k] + (uint32_t)((in_len[big[k]] + 1023) / 1024); big_bytes += in_len[big[k]]; }
    const size_t a_sel = PA.add(big.data(), big.size() * 4), a_seg = PA.add(seg0.data(), seg0.size() * 4);
    const uint32_t cand_cap = (uint32_t)std::min<uint64_t>(1u << 24, big_bytes / 512 + 4096);
    const size_t a_cm = PA.reserve((size_t)cand_cap * 4), a_cb = PA.reserve((size_t)cand_cap * 8), a_cc = PA.reserve(64);
    const uint32_t q_cap = (uint32_t)std::min<uint64_t>(1u << 27, big_bytes / 32 + 8192);      // ~0.1 % of bit offsets pass the cheap tests
    const size_t a_qm = PA.reserve((size_t)q_cap * 4), a_qb = PA.reserve((size_t)q_cap * 8);
    CK(PA.commit(ctx->stream));
    std::vector<uint32_t> c_member; std::vector<uint64_t> c_bit;
    if (ctx->feed_n && big.empty()) { CK(cudaStreamWaitEvent(ctx->stream, ctx->feed_ev[ctx->feed_n - 1], 0)); ctx->feed_n = 0; }
    if (!big.empty()) {
        FindDev F;
        F.in = d_in; F.in_off = PA.ptr<uint64_t>(a_io); F.in_len = PA.ptr<uint64_t>(a_il);
        F.members = PA.ptr<uint32_t>(a_sel); F.seg0 = PA.ptr<uint32_t>(a_seg); F.n_sel = (uint32_t)big.size(); F.n_segs = seg0.back();
        F.cand_member = PA.ptr<uint32_t>(a_cm); F.cand_bit = PA.ptr<uint64_t>(a_cb); F.cand_count = PA.ptr<uint32_t>(a_cc); F.cand_cap = cand_cap;
        F.q_member = PA.ptr<uint32_t>(a_qm); F.q_bit = PA.ptr<uint64_t>(a_qb); F.q_count = PA.ptr<uint32_t>(a_cc) + 1; F.q_cap = q_cap;
        F.q_done = PA.ptr<uint32_t>(a_cc) + 2;
        CK(cudaMemsetAsync(F.cand_count, 0, 16, ctx->stream));
        ctx->tm.mark(ctx->stream, "find_blocks");
        if (ctx->feed_n) {
            // the input is still arriving (b2f_decode_batch): scan what each piece completes.  A segment is ready when its 1 KiB, the
            // look-ahead of the cheap tests and the longest dynamic header behind its last offset (< 1 KiB) are there.
            uint32_t s_lo = 0;
            for (uint32_t k = 0; k < ctx->feed_n; k++) {
                uint32_t s_hi = F.n_segs;
                if (k + 1 < ctx->feed_n) {
                    s_hi = 0;
                    for (size_t j = 0; j < big.size(); j++) {
                        const uint64_t o = in_off[big[j]], l = in_len[big[j]], have = ctx->feed_end[k];
                        const uint32_t ns = seg0[j + 1] - seg0[j];
                        if (o + l <= have) { s_hi = seg0[j + 1]; continue; }
                        if (have > o + 2048) s_hi = seg0[j] + (uint32_t)std::min<uint64_t>(ns, (have - o - 2048) / 1024 + 1);
                        else s_hi = seg0[j];
                        break;
                    }
                }
                CK(cudaStreamWaitEvent(ctx->stream, ctx->feed_ev[k], 0));
                if (s_hi > s_lo) { CK(dec_launch_find(F, s_lo, s_hi, ctx->stream)); ctx->stats.kernel_launches += 3; s_lo = s_hi; }
            }
            ctx->feed_n = 0;
        } else {
            CK(dec_launch_find(F, 0, F.n_segs, ctx->stream));
            ctx->stats.kernel_launches += 3;
        }
        ctx->tm.mark(ctx->stream, "sync");
        CK(ctx->pin_res.ensure(64));
        uint32_t *h_cnt = ctx->pin_res.as<uint32_t>();
        CK(cudaMemcpyAsync(h_cnt, F.cand_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        uint32_t nc = h_cnt[0];
        if (nc > cand_cap || h_cnt[1] > q_cap) { for (uint32_t m : big) serial.push_back(m); big.clear(); nc = 0; }    // absurd candidate count: do not trust
        if (nc) {
            CK(ctx->pin_cand.ensure((size_t)nc * 12 + 64));
            uint8_t *hc = ctx->pin_cand.as<uint8_t>();
            CK(cudaMemcpyAsync(hc, F.cand_member, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(hc + align_up((size_t)nc * 4, 8), F.cand_bit, (size_t)nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            c_member.assign((uint32_t *)hc, (uint32_t *)hc + nc);
            const uint64_t *hb = (const uint64_t *)(hc + align_up((size_t)nc * 4, 8));
            c_bit.assign(hb, hb + nc);
        }
    }
    // ---- phase B: every candidate (plus bit 0 of every big member) becomes a speculative block
    std::vector<std::pair<uint32_t, uint64_t>> cands;      // (member, bit) sorted
    for (size_t i = 0; i < c_member.size(); i++) cands.push_back({ c_member[i], c_bit[i] });
    for (uint32_t m : big) cands.push_back({ m, 0 });
    std::sort(cands.begin(), cands.end());
    cands.erase(std::unique(cands.begin(), cands.end()), cands.end());
    SpecDev S; memset(&S, 0, sizeof S);
    std::vector<uint32_t> sel_blocks;                      // indices into cands of the verified chains
    std::vector<uint64_t> k_out, k_len;                    // per selected block
    std::vector<uint32_t> serial_spec;                     // members given up by the speculative path
    for (int attempt = 0; attempt < 12; attempt++) {
    bool retry = false;
    sel_blocks.clear(); k_out.clear(); k_len.clear(); serial_spec.clear();
    const size_t ncand = cands.size();
    if (ncand) {
        std::vector<uint32_t> bm(ncand), seg0(ncand + 1, 0), cta0(ncand + 1, 0); std::vector<uint64_t> bb(ncand), be(ncand);
        bool too_big = false;
        for (size_t i = 0; i < ncand; i++) {
            bm[i] = cands[i].first; bb[i] = cands[i].second;
            // A candidate's extent runs to the candidate AFTER the next one: the finder's rare false positives (about one per 300 Mbit)
            // lie inside a true block, and with the doubled extent that block's parse simply runs across it to its EndOfBlock -- no second
            // parse of the whole stream.  The parse is latency bound (one thread per subsegment), so the extra threads are nearly
            // free; two false positives in a row still take the drop-and-retry path below.
            be[i] = (i + 2 < ncand && cands[i + 2].first == bm[i]) ? cands[i + 2].second : in_len[bm[i]] * 8;
            const uint64_t bits = be[i] - bb[i];
            if (bits >= 0xFFFF0000ull) too_big = true;
            const uint32_t ns = (uint32_t)((bits + kSpecBits - 1) / kSpecBits);
            seg0[i + 1] = seg0[i] + ns; cta0[i + 1] = cta0[i] + (ns + kSpecCta - 1) / kSpecCta;
        }
        if (too_big) { for (uint32_t m : big) serial_spec.push_back(m); big.clear(); }
        else {
            const uint32_t nseg = seg0.back();
            Packer PB(ctx->pin_cand, ctx->buf[NB_DEC_CAND]);
            const size_t b_m = PB.add(bm.data(), ncand * 4), b_b = PB.add(bb.data(), ncand * 8), b_e = PB.add(be.data(), ncand * 8),
                         b_s0 = PB.add(seg0.data(), (ncand + 1) * 4), b_c0 = PB.add(cta0.data(), (ncand + 1) * 4);
            const size_t r_dr = PB.reserve(ncand * 4), r_fl = PB.reserve(ncand * 4), r_st = PB.reserve(ncand * 4), r_ee = PB.reserve(ncand * 4),
                         r_es = PB.reserve(ncand * 4), r_no = PB.reserve(ncand * 8), r_nt = PB.reserve(ncand * 8), r_end = PB.reserve(16);
            CK(PB.commit(ctx->stream));
            CK(ctx->buf[NB_SPEC_TAB].ensure(ncand * sizeof(InflateTables) + 256));
            const size_t seg_bytes = (size_t)nseg * (6 * 4 + 3 * 8) + 10 * 256 + 1024;
            CK(ctx->buf[NB_SPEC_SEG].ensure(seg_bytes));
            uint8_t *sp = ctx->buf[NB_SPEC_SEG].as<uint8_t>();
            S.in = d_in; S.in_off = PA.ptr<uint64_t>(a_io); S.in_len = PA.ptr<uint64_t>(a_il);
            S.n_blocks = (uint32_t)ncand; S.blk_member = PB.ptr<uint32_t>(b_m); S.blk_bit = PB.ptr<uint64_t>(b_b); S.blk_end = PB.ptr<uint64_t>(b_e);
            S.blk_seg0 = PB.ptr<uint32_t>(b_s0); S.blk_cta0 = PB.ptr<uint32_t>(b_c0); S.n_segs = nseg; S.n_ctas = cta0.back();
            S.tabs = ctx->buf[NB_SPEC_TAB].as<InflateTables>();
            S.blk_data_rel = PB.ptr<uint32_t>(r_dr); S.blk_flags = PB.ptr<uint32_t>(r_fl); S.blk_status = PB.ptr<uint32_t>(r_st);
            S.blk_eob_end = PB.ptr<uint32_t>(r_ee); S.blk_eob_seg = PB.ptr<uint32_t>(r_es); S.blk_nout = PB.ptr<uint64_t>(r_no); S.blk_ntok = PB.ptr<uint64_t>(r_nt);
            S.s_start = carve<uint32_t>(sp, nseg); S.s_exit = carve<uint32_t>(sp, nseg); S.s_exit_prev = carve<uint32_t>(sp, nseg);
            S.s_eob_end = carve<uint32_t>(sp, nseg); S.s_nsym = carve<uint32_t>(sp, nseg); S.s_nbytes = carve<uint32_t>(sp, nseg);
            S.s_out_rel = carve<uint64_t>(sp, nseg); S.s_tok_rel = carve<uint64_t>(sp, nseg); S.s_min_src = carve<int64_t>(sp, nseg);
            S.changed = carve<uint32_t>(sp, 64);
            CK(cudaMemsetAsync(S.changed, 0, 256, ctx->stream));
            ctx->tm.mark(ctx->stream, "spec_parse");
            uint32_t rounds = 5;
            CK(spec_launch_parse(S, 0, rounds, ctx->stream));
            ctx->stats.kernel_launches += 7;
            ctx->tm.mark(ctx->stream, "sync");
            const size_t res_bytes = r_end - r_dr;
            CK(ctx->pin_res.ensure(res_bytes + 256 + 64));
            uint8_t *hr = ctx->pin_res.as<uint8_t>();
            const uint32_t *h_changed = (const uint32_t *)(hr + align_up(res_bytes, 16));
            const uint32_t *h_fl = (const uint32_t *)(hr + (r_fl - r_dr)), *h_st = (const uint32_t *)(hr + (r_st - r_dr)), *h_ee = (const uint32_t *)(hr + (r_ee - r_dr));
            for (;;) {
                CK(cudaMemcpyAsync(hr, PB.ptr<uint8_t>(r_dr), res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(hr + align_up(res_bytes, 16), S.changed, 256, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                // Five rounds settle ordinary data.  Where codes do not self-synchronise (long runs: one- and two-bit codes) a wrong
                // start moves on by one subsegment per round; as long as the last round still changed something and a block fails
                // verification, iterate further (cheap: only the changed subsegments are decoded again) instead of giving the
                // member to the in-order kernel.
                bool any_bad = false;
                for (size_t i = 0; i < ncand && !any_bad; i++) any_bad = h_st[i] == 1;
                if (!any_bad || h_changed[rounds - 1] == 0 || rounds + 8 > 61) break;
                ctx->tm.mark(ctx->stream, "spec_parse");
                CK(spec_launch_parse(S, rounds, rounds + 8, ctx->stream));
                rounds += 8; ctx->stats.kernel_launches += 9;
                ctx->tm.mark(ctx->stream, "sync");
            }
            const uint64_t *h_no = (const uint64_t *)(hr + (r_no - r_dr)), *h_nt = (const uint64_t *)(hr + (r_nt - r_dr));
            // ---- phase C: chain walk from bit 0 of every big member
            std::vector<uint32_t> blk_sel(ncand, 0); std::vector<uint64_t> blk_out0(ncand, 0), blk_tok0(ncand, 0);
            uint64_t tok_total = 0;
            std::vector<std::pair<uint32_t, uint64_t>> drop;
            for (uint32_t m : big) {
                uint64_t pos = 0, out = 0; bool ok = true, fin = false;
                const size_t first_sel = sel_blocks.size(); const uint64_t tok_first = tok_total;
                while (!fin) {
                    auto it = std::lower_bound(cands.begin(), cands.end(), std::make_pair(m, pos));
                    if (it == cands.end() || it->first != m || it->second != pos) { ok = false; break; }
                    const size_t ci = (size_t)(it - cands.begin());
                    if (h_st[ci] != 0) { ok = false; break; }
                    // The sub-block decoder fetches whole words, so a symbol can be "completed" by bytes beyond the member (stale or zero
                    // fill): a block whose EndOfBlock lies past the member's last bit is a truncated stream -> in-order kernel (UnexpectedEof)
                    if (pos + h_ee[ci] > in_len[m] * 8) { ok = false; break; }
                    sel_blocks.push_back((uint32_t)ci); k_out.push_back(out_off[m] + out); k_len.push_back(h_no[ci]);
                    blk_out0[ci] = out_off[m] + out; blk_tok0[ci] = tok_total;
                    tok_total += h_nt[ci]; out += h_no[ci]; pos += h_ee[ci]; fin = (h_fl[ci] & 1u) != 0;
                }
                if (ok && out > out_cap[m]) ok = false;
                if (!ok && getenv("B2F_DEBUG")) {
                    auto it = std::lower_bound(cands.begin(), cands.end(), std::make_pair(m, pos));
                    const bool found = it != cands.end() && it->first == m && it->second == pos;
                    const size_t ci = found ? (size_t)(it - cands.begin()) : 0;
                    fprintf(stderr, "[b2f] member %u: chain broke at bit %llu after %zu blocks (candidate %s, status %u, out %llu cap %llu)\n", m,
                            (unsigned long long)pos, sel_blocks.size() - first_sel, found ? "found" : "missing", found ? h_st[ci] : 99u,
                            (unsigned long long)out, (unsigned long long)out_cap[m]);
                    if (found) {   // dump the verification state of the failing block
                        const uint32_t s0 = seg0[ci], ns = seg0[ci + 1] - seg0[ci];
                        std::vector<uint32_t> hs(ns), he(ns), hp(ns);
                        cudaMemcpy(hs.data(), S.s_start + s0, ns * 4, cudaMemcpyDeviceToHost);
                        cudaMemcpy(he.data(), S.s_exit + s0, ns * 4, cudaMemcpyDeviceToHost);
                        cudaMemcpy(hp.data(), S.s_exit_prev + s0, ns * 4, cudaMemcpyDeviceToHost);
                        uint32_t dr = 0; cudaMemcpy(&dr, S.blk_data_rel + ci, 4, cudaMemcpyDeviceToHost);
                        fprintf(stderr, "[b2f]   block %zu: bit %llu len %llu bits, %u subsegments, data_rel %u\n", ci, (unsigned long long)bb[ci],
                                (unsigned long long)(be[ci] - bb[ci]), ns, dr);
                        uint32_t shown = 0;
                        for (uint32_t k = 0; k < ns && shown < 12; k++) {
                            const uint32_t want = k == 0 ? dr : he[k - 1];
                            if (k == 0 || hs[k] != want || he[k] >= 0xFFFFFFFDu) { fprintf(stderr, "[b2f]     k=%u start=%u want=%u exitA=%u exitB=%u\n", k, hs[k], want, he[k], hp[k]); shown++; }
                        }
                    }
                }
                if (!ok) {
                    sel_blocks.resize(first_sel); k_out.resize(first_sel); k_len.resize(first_sel); tok_total = tok_first; serial_spec.push_back(m);
                    // A false-positive candidate inside a true block cuts that block short (no EndOfBlock before the next candidate):
                    // drop the candidate that follows the failing block and parse again.
                    auto it = std::lower_bound(cands.begin(), cands.end(), std::make_pair(m, pos));
                    if (it != cands.end() && it->first == m && it->second == pos && (it + 1) != cands.end() && (it + 1)->first == m && out <= out_cap[m]) {
                        drop.push_back(*(it + 1)); retry = true;
                        // Keep walking to collect the other false positives of this member in the same attempt: a block whose chain is
                        // consistent but hits the end of its extent without EndOfBlock (status 2) was cut by the candidate after it; the
                        // candidate after THAT one is presumed true again, and verified blocks (status 0) are followed exactly.
                        size_t idx = (size_t)(it - cands.begin());
                        while (idx < ncand && cands[idx].first == m) {
                            if (h_st[idx] == 2) {
                                if (idx + 1 >= ncand || cands[idx + 1].first != m) break;
                                drop.push_back(cands[idx + 1]);
                                idx += 2;
                            } else if (h_st[idx] == 0 && !(h_fl[idx] & 1u)) {
                                auto nx = std::lower_bound(cands.begin(), cands.end(), std::make_pair(m, cands[idx].second + h_ee[idx]));
                                if (nx == cands.end() || nx->first != m || nx->second != cands[idx].second + h_ee[idx]) break;
                                idx = (size_t)(nx - cands.begin());
                            } else break;
                        }
                    }
                    continue;
                }
                is_par[m] = 1; st[m] = kInfOk; olen[m] = out; cons[m] = (pos + 7) >> 3;
            }
            if (retry && attempt + 1 < 12) {
                std::sort(drop.begin(), drop.end()); drop.erase(std::unique(drop.begin(), drop.end()), drop.end());
                std::vector<std::pair<uint32_t, uint64_t>> kept;
                for (auto &cnd : cands) if (!std::binary_search(drop.begin(), drop.end(), cnd)) kept.push_back(cnd);
                cands.swap(kept);
                for (uint32_t m : big) is_par[m] = 0;
                ctx->tm.mark(ctx->stream, "spec_retry");
                continue;
            }
            for (uint32_t ci : sel_blocks) blk_sel[ci] = 1;
            const size_t nsel = sel_blocks.size();
            if (nsel) {
                CK(ctx->buf[NB_SPEC_TOK].ensure(tok_total * 4 + 256));
                Packer PS(ctx->pin_sel, ctx->buf[NB_SPEC_SEL]);
                const size_t s_sel = PS.add(blk_sel.data(), ncand * 4), s_o0 = PS.add(blk_out0.data(), ncand * 8), s_t0 = PS.add(blk_tok0.data(), ncand * 8),
                             s_lst = PS.add(sel_blocks.data(), nsel * 4);
                // segment slots: a block of n output bytes is resolved as ceil(n / kSegBytes) segments (spec_kernels.cu)
                std::vector<uint32_t> slot0(nsel + 1, 0);
                uint64_t out_hi = 0;
                for (size_t k = 0; k < nsel; k++) {
                    slot0[k + 1] = slot0[k] + (uint32_t)std::max<uint64_t>(1, (k_len[k] + kSegBytes - 1) / kSegBytes);
                    out_hi = std::max<uint64_t>(out_hi, k_out[k] + k_len[k]);
                }
                const uint32_t nslots = slot0.back();
                const size_t s_u0 = PS.add(slot0.data(), (nsel + 1) * 4);
                const size_t s_st = PS.reserve((size_t)nslots * 8), s_so = PS.reserve((size_t)nslots * 8), s_sn = PS.reserve((size_t)nslots * 4),
                             s_sb = PS.reserve((size_t)nslots * 4), s_sm = PS.reserve((size_t)nslots * 4), s_sr = PS.reserve((size_t)nslots * 4),
                             s_sc = PS.reserve((size_t)nslots), s_rec = PS.reserve(((size_t)nslots + 2) * 16), s_cl = PS.reserve((size_t)nslots * 4),
                             s_cc = PS.reserve(kChainCounters * kMaxParts * 4), s_sl = PS.reserve((size_t)nslots * 4), s_tl = PS.reserve((size_t)nslots * 4),
                             s_ce = PS.reserve((size_t)nslots * 8), s_cn = PS.reserve((size_t)nslots * 4);
                const size_t s_err = PS.reserve(n * 4);
                CK(PS.commit(ctx->stream));
                CK(ctx->buf[NB_SPEC_SYM].ensure(out_hi * 2 + 256));
                S.blk_sel = PS.ptr<uint32_t>(s_sel); S.blk_out0 = PS.ptr<uint64_t>(s_o0); S.blk_tok0 = PS.ptr<uint64_t>(s_t0); S.sel_blocks = PS.ptr<uint32_t>(s_lst);
                S.mem_out_off = PA.ptr<uint64_t>(a_oo);
                S.tokens = ctx->buf[NB_SPEC_TOK].as<uint32_t>(); S.out = d_out;
                S.n_sel = (uint32_t)nsel; S.n_slots = nslots; S.sel_slot0 = PS.ptr<uint32_t>(s_u0);
                S.seg_tok = PS.ptr<uint64_t>(s_st); S.seg_out = PS.ptr<uint64_t>(s_so); S.seg_ntok = PS.ptr<uint32_t>(s_sn); S.seg_nout = PS.ptr<uint32_t>(s_sb);
                S.seg_member = PS.ptr<uint32_t>(s_sm); S.seg_reach = PS.ptr<uint32_t>(s_sr); S.seg_cut = PS.ptr<uint8_t>(s_sc);
                S.seg_rec = PS.ptr<uint4>(s_rec); S.chain_list = PS.ptr<uint32_t>(s_cl); S.chain_count = PS.ptr<uint32_t>(s_cc);
                S.soft_list = PS.ptr<uint32_t>(s_sl); S.tail_list = PS.ptr<uint32_t>(s_tl); S.chain_end = PS.ptr<uint64_t>(s_ce); S.chain_next_soft = PS.ptr<uint32_t>(s_cn);
                S.sym16 = ctx->buf[NB_SPEC_SYM].as<uint16_t>(); S.mem_err = PS.ptr<uint32_t>(s_err);
                CK(cudaMemsetAsync(S.mem_err, 0, n * 4, ctx->stream));
                CK(cudaMemsetAsync(S.chain_count, 0, kChainCounters * kMaxParts * 4, ctx->stream));
                CK(cudaMemsetAsync(S.seg_rec + nslots, 0, 32, ctx->stream));       // the two read-ahead records behind the last slot
                // The LZ77 resolution is a pipeline over PARTS (runs of whole blocks): tokens -> segments (markers) -> substitution for
                // part p, then part p+1, ...; as soon as a part is final its output is copied to the caller's memory on a side stream
                // (page-locked destination: one DMA per member run; pageable: staged through the pinned buffers, the calling thread
                // copying piece k out while piece k+1 arrives) while the following parts are still being resolved.
                {
                    uint64_t total_len = 0; for (size_t k = 0; k < nsel; k++) total_len += k_len[k];
                    bool any_host = false;
                    for (size_t k = 0; k < nsel; k++) if (mem[cands[sel_blocks[k]].first].h_out) { any_host = true; break; }
                    const uint32_t nparts = any_host ? (uint32_t)std::min<uint64_t>(ctx->max_parts, std::max<uint64_t>(1, total_len / (48u << 20))) : 1u;
                    size_t part_b[kMaxParts + 1] = { 0 }; uint64_t acc = 0;
                    for (uint32_t part = 0; part < nparts; part++) {
                        size_t b1 = part_b[part];
                        const uint64_t want = total_len * (part + 1) / nparts;
                        while (b1 < nsel && (acc < want || part + 1 == nparts)) { acc += k_len[b1]; b1++; }
                        part_b[part + 1] = b1;
                    }
                    S.n_parts = nparts;
                    for (uint32_t part = 0; part <= nparts; part++) S.part_slot0[part] = slot0[part_b[part]];
                    // the token pass is one launch for all parts (a thread per subsegment: its duration is the latency of one thread, so
                    // per-part launches would pay that latency once per part)
                    ctx->tm.mark(ctx->stream, "spec_tokens");
                    CK(spec_launch_tokens(S, cta0[sel_blocks[0]], cta0[sel_blocks[nsel - 1] + 1], ctx->stream));
                    ctx->stats.kernel_launches += 1;
                    for (uint32_t part = 0; part < nparts; part++) {
                        if (part_b[part + 1] == part_b[part]) continue;
                        ctx->tm.mark(ctx->stream, "lz_resolve");
                        CK(spec_launch_segments(S, part, ctx->stream));
                        ctx->tm.mark(ctx->stream, "lz_subst");
                        CK(spec_launch_subst(S, part, ctx->stream));
                        ctx->stats.kernel_launches += 8;
                        if (any_host) CK(cudaEventRecord(ctx->part_ev[part], ctx->stream));
                    }
                    ctx->tm.mark(ctx->stream, "sync");
                    if (any_host) {
                        for (uint32_t part = 0; part < nparts; part++) {
                            if (part_b[part + 1] == part_b[part]) continue;
                            CK(cudaStreamWaitEvent(ctx->aux[0], ctx->part_ev[part], 0));
                            for (size_t k = part_b[part]; k < part_b[part + 1];) {
                                // consecutive blocks of one member are contiguous in out: one copy per run
                                const uint32_t m = cands[sel_blocks[k]].first;
                                size_t k1 = k; uint64_t run = 0;
                                while (k1 < part_b[part + 1] && cands[sel_blocks[k1]].first == m) { run += k_len[k1]; k1++; }
                                if (mem[m].h_out && run)
                                    CK(d2h_copy(ctx, mem[m].h_out + (k_out[k] - out_off[m]), d_out + k_out[k], run, ctx->aux[0], mem[m].h_pinned));
                                k = k1;
                            }
                        }
                    }
                    for (uint32_t m : big) if (is_par[m] && mem[m].h_out) copied[m] = 1;
                }
                CK(ctx->pin_res.ensure(n * 4 + 64));
                uint32_t *h_err = ctx->pin_res.as<uint32_t>();
                CK(cudaMemcpyAsync(h_err, S.mem_err, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
                // a member with a match that reaches before its first byte (libflate: "Too long backword reference") or an inconsistent
                // size: redo it with the in-order kernel, which reproduces the reference's error and partial output
                std::vector<char> redo(n, 0);
                for (uint32_t m : big) if (is_par[m] && h_err[m]) redo[m] = 1;
                bool any_redo = false;
                for (uint32_t m : big) if (is_par[m] && redo[m]) { is_par[m] = 0; copied[m] = 0; serial_spec.push_back(m); any_redo = true; }
                if (any_redo) CK(cudaStreamSynchronize(ctx->aux[0]));     // early copies of a redone member must not land after its final copy
            }
        }
    }
    break;
    }   // attempts
    for (uint32_t m : serial_spec) if (!is_par[m]) serial.push_back(m);
    for (size_t i = 0; i < n; i++) if (in_len[i] < kParallelMinBytes) serial.push_back((uint32_t)i);
    // ---- in-order kernel for everything that is not on a verified chain
    const size_t nser = serial.size();
    ctx->n_inorder_members += nser; ctx->n_spec_members += n - nser;
    if (nser) {
        Packer PC(ctx->pin_blk, ctx->buf[NB_DEC_BLK]);
        std::vector<uint64_t> s_io(nser), s_il(nser), s_oo(nser), s_oc(nser);
        for (size_t k = 0; k < nser; k++) { uint32_t m = serial[k]; s_io[k] = in_off[m]; s_il[k] = full_len[m]; s_oo[k] = out_off[m]; s_oc[k] = out_cap[m]; }
        const size_t s_a = PC.add(s_io.data(), nser * 8), s_b = PC.add(s_il.data(), nser * 8), s_c = PC.add(s_oo.data(), nser * 8), s_d = PC.add(s_oc.data(), nser * 8);
        const size_t s_st = PC.reserve(nser * 4), s_ol = PC.reserve(nser * 8), s_cs = PC.reserve(nser * 8), s_gl = PC.reserve(nser * 8);
        CK(PC.commit(ctx->stream));
        ctx->tm.mark(ctx->stream, "inflate_inorder");
        DecDev D;
        D.in = d_in; D.in_off = PC.ptr<uint64_t>(s_a); D.in_len = PC.ptr<uint64_t>(s_b);
        D.out = d_out; D.out_off = PC.ptr<uint64_t>(s_c); D.out_cap = PC.ptr<uint64_t>(s_d); D.n = (uint32_t)nser;
        D.status = PC.ptr<int32_t>(s_st); D.out_len = PC.ptr<uint64_t>(s_ol); D.consumed = PC.ptr<uint64_t>(s_cs); D.good_len = PC.ptr<uint64_t>(s_gl);
        CK(dec_launch_serial(D, ctx->stream));
        ctx->stats.kernel_launches += 1;
        ctx->tm.mark(ctx->stream, "results");
        const size_t res_bytes = PC.off - s_st;
        CK(ctx->pin_res.ensure(res_bytes + 64));
        uint8_t *hr = ctx->pin_res.as<uint8_t>();
        CK(cudaMemcpyAsync(hr, PC.ptr<uint8_t>(s_st), res_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (size_t k = 0; k < nser; k++) {
            uint32_t m = serial[k];
            st[m] = ((int32_t *)hr)[k]; olen[m] = ((uint64_t *)(hr + (s_ol - s_st)))[k]; cons[m] = ((uint64_t *)(hr + (s_cs - s_st)))[k];
            good[m] = ((uint64_t *)(hr + (s_gl - s_st)))[k];
        }
    }
    for (size_t i = 0; i < n; i++) if (is_par[i]) good[i] = olen[i];
    return B2F_OK;
}

// warp per request: copies len[i] bytes from src + soff[i] to dst + doff[i]
__global__ void k_gather_windows(const uint8_t *src, const uint64_t *soff, const uint64_t *len, const uint64_t *doff, uint8_t *dst, uint32_t n) {
    uint32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint8_t *s = src + soff[i]; uint8_t *d = dst + doff[i];
    for (uint64_t k = lane; k < len[i]; k += 32) d[k] = s[k];
}

// Access to the few container bytes the host needs (headers, trailers).  Host inputs are read in place; device-resident
// inputs are fetched as small windows with ONE gather kernel + ONE copy per round.
struct InputAccess {
    b2f_ctx *ctx; const uint8_t *const *h_in; const uint8_t *d_in; const uint64_t *in_off; const size_t *in_len;
    int fetch(const std::vector<size_t> &streams, const std::vector<size_t> &pos, size_t want, std::vector<std::vector<uint8_t>> &out) {
        const size_t n = streams.size();
        out.assign(n, std::vector<uint8_t>());
        if (!n) return B2F_OK;
        std::vector<uint64_t> meta(3 * n); uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            size_t s = streams[i];
            size_t avail = in_len[s] > pos[i] ? in_len[s] - pos[i] : 0, k = std::min(avail, want);
            if (h_in) { out[i].assign(h_in[s] + pos[i], h_in[s] + pos[i] + k); continue; }
            meta[i] = in_off[s] + pos[i]; meta[n + i] = k; meta[2 * n + i] = total; total += k;
        }
        if (h_in || total == 0) return B2F_OK;
        CK(ctx->buf[NB_DEC_META].ensure(3 * n * 8 + total + 256));
        CK(ctx->pin_win.ensure(3 * n * 8 + total + 256));
        uint8_t *hw = ctx->pin_win.as<uint8_t>(); uint8_t *dw = ctx->buf[NB_DEC_META].as<uint8_t>();
        memcpy(hw, meta.data(), 3 * n * 8);
        CK(cudaMemcpyAsync(dw, hw, 3 * n * 8, cudaMemcpyHostToDevice, ctx->stream));
        const uint64_t *dm = (const uint64_t *)dw;
        k_gather_windows<<<(unsigned)((n + 3) / 4), 128, 0, ctx->stream>>>(d_in, dm, dm + n, dm + 2 * n, dw + 3 * n * 8, (uint32_t)n);
        CK(cudaGetLastError());
        ctx->stats.kernel_launches += 1;
        CK(cudaMemcpyAsync(hw + 3 * n * 8, dw + 3 * n * 8, total, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < n; i++) out[i].assign(hw + 3 * n * 8 + meta[2 * n + i], hw + 3 * n * 8 + meta[2 * n + i] + meta[n + i]);
        return B2F_OK;
    }
};

// Shared decode driver: container framing on the host (a few bytes per stream), DEFLATE + checksums on the device.
int decode_core(b2f_ctx *ctx, int fmt, size_t n_streams, InputAccess &IA, const uint8_t *d_in, const uint64_t *in_off, const size_t *in_len,
                uint8_t *d_out, const uint64_t *out_off, const size_t *out_cap, size_t *out_len, size_t *in_consumed, int *status,
                uint8_t *const *h_out = nullptr, std::vector<uint64_t> *h_copied = nullptr, const std::vector<char> *h_pinned = nullptr) {
    std::vector<size_t> pos(n_streams, 0);            // reader position per stream
    std::vector<uint64_t> produced(n_streams, 0);
    std::vector<char> done(n_streams, 0);
    std::vector<uint64_t> prev_comp(n_streams, 0);    // compressed size of the stream's previous member (MultiDecoder rounds)
    ctx->last_good.assign(n_streams, 0);
    for (size_t s = 0; s < n_streams; s++) { status[s] = B2F_OK; out_len[s] = 0; in_consumed[s] = 0; }
    bool first = true;
    for (;;) {
        // ---- container headers
        std::vector<size_t> act;
        for (size_t s = 0; s < n_streams; s++) if (!done[s]) act.push_back(s);
        if (act.empty()) break;
        if (fmt != B2F_FMT_DEFLATE) {
            std::vector<size_t> todo = act, apos;
            size_t want = 4096;
            while (!todo.empty()) {
                apos.clear(); for (size_t s : todo) apos.push_back(pos[s]);
                std::vector<std::vector<uint8_t>> win;
                int rc = IA.fetch(todo, apos, want, win); if (rc) return rc;
                std::vector<size_t> again;
                for (size_t i = 0; i < todo.size(); i++) {
                    size_t s = todo[i];
                    HostReader r = { win[i].data(), win[i].size(), 0 };
                    int hrc = fmt == B2F_FMT_ZLIB ? parse_zlib_header(r) : parse_gzip_header(r);
                    if (hrc == B2F_ERR_UNEXPECTED_EOF && pos[s] + win[i].size() < in_len[s]) { again.push_back(s); continue; }   // window too small
                    if (hrc != B2F_OK) {
                        done[s] = 1; in_consumed[s] = pos[s] + r.pos;
                        // MultiDecoder: EOF while reading the NEXT header ends the stream cleanly (gzip.rs:1148-1156)
                        if (!(!first && hrc == B2F_ERR_UNEXPECTED_EOF)) status[s] = hrc;
                        continue;
                    }
                    pos[s] += r.pos;
                }
                todo.swap(again);
                want = want < (1u << 20) ? (1u << 20) : (size_t)-1;
            }
        }
        std::vector<Member> mem;
        for (size_t s : act) {
            if (done[s]) continue;
            Member m = { s, in_off[s] + pos[s], in_len[s] - pos[s], out_off[s] + produced[s], out_cap[s] > produced[s] ? out_cap[s] - produced[s] : 0,
                         (h_out && h_out[s]) ? h_out[s] + produced[s] : nullptr, h_pinned && (*h_pinned)[s] != 0, 0 };
            // A first member may be the whole file: the finder scans all of it.  Later members of a multi-member file (BGZF, pigz -i,
            // concatenated .gz) are sized like their predecessor: scanning the whole remainder for each would cost
            // O(members x bytes), so the speculative path gets a window of 4 predecessors + 1 MiB, and small members go straight to
            // the in-order kernel (which stops at BFINAL by itself).  A wrong guess only costs speed: the chain breaks at the window
            // end and the in-order kernel decodes the member.
            m.scan_len = first ? m.def_len : (prev_comp[s] < kParallelMinBytes ? 0 : std::min<uint64_t>(m.def_len, 4 * prev_comp[s] + (1u << 20)));
            mem.push_back(m);
        }
        if (mem.empty()) break;
        std::vector<int> st; std::vector<uint64_t> olen, cons, good; std::vector<char> copied;
        int rc = inflate_round(ctx, d_in, d_out, mem, st, olen, cons, copied, good);
        if (rc) return rc;
        for (size_t i = 0; i < mem.size(); i++) ctx->last_good[mem[i].stream] = produced[mem[i].stream] + std::min<uint64_t>(good[i], mem[i].out_cap);
        // bytes of each stream that are already on the host (a prefix: members are decoded in order)
        if (h_copied) for (size_t i = 0; i < mem.size(); i++) if (copied[i] && (*h_copied)[mem[i].stream] == produced[mem[i].stream]) (*h_copied)[mem[i].stream] += std::min<uint64_t>(olen[i], mem[i].out_cap);
        // ---- trailers + checksums of what was produced in this round
        std::vector<uint64_t> ck_off, ck_len; std::vector<size_t> ck_idx, ck_streams, ck_pos;
        for (size_t i = 0; i < mem.size(); i++) {
            size_t s = mem[i].stream;
            uint64_t wrote = std::min<uint64_t>(olen[i], mem[i].out_cap);
            pos[s] += cons[i]; prev_comp[s] = cons[i];
            produced[s] += olen[i]; out_len[s] = produced[s]; in_consumed[s] = pos[s];
            if (st[i] != kInfOk) { status[s] = map_inf_status(st[i]); done[s] = 1; continue; }
            if (fmt == B2F_FMT_DEFLATE) { done[s] = 1; continue; }
            ck_off.push_back(mem[i].out_off); ck_len.push_back(wrote); ck_idx.push_back(i); ck_streams.push_back(s); ck_pos.push_back(pos[s]);
        }
        if (!ck_idx.empty()) {
            std::vector<uint32_t> crc, adler;
            rc = run_checksums(ctx, d_out, ck_off, ck_len, fmt != B2F_FMT_ZLIB, fmt == B2F_FMT_ZLIB, nullptr, crc, adler);
            if (rc) return rc;
            std::vector<std::vector<uint8_t>> tw;
            rc = IA.fetch(ck_streams, ck_pos, 8, tw); if (rc) return rc;
            for (size_t j = 0; j < ck_idx.size(); j++) {
                size_t s = ck_streams[j];
                HostReader r = { tw[j].data(), tw[j].size(), 0 };
                if (fmt == B2F_FMT_ZLIB) {               // zlib.rs:377-409
                    uint8_t t[4];
                    if (!rd_exact(r, t, 4)) status[s] = B2F_ERR_UNEXPECTED_EOF;
                    else { uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3]; if (want != adler[j]) status[s] = B2F_ERR_INVALID_DATA; }
                    done[s] = 1;
                } else {                                 // gzip.rs:1018-1047: CRC32 checked, ISIZE read but not checked
                    uint8_t t[8];
                    if (!rd_exact(r, t, 4) || !rd_exact(r, t + 4, 4)) { status[s] = B2F_ERR_UNEXPECTED_EOF; done[s] = 1; }
                    else {
                        uint32_t want = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
                        if (want != crc[j]) { status[s] = B2F_ERR_INVALID_DATA; done[s] = 1; }
                        else if (fmt == B2F_FMT_GZIP) done[s] = 1;
                    }
                }
                pos[s] += r.pos; in_consumed[s] = pos[s];
            }
        }
        first = false;
        if (fmt != B2F_FMT_GZIP_MULTI) break;
    }
    return B2F_OK;
}
}  // namespace

extern "C" int b2f_decode_batch(b2f_ctx *ctx, int fmt, size_t n_streams, const uint8_t *const *in, const size_t *in_len,
                                uint8_t *const *out, const size_t *out_cap, size_t *out_len, size_t *in_consumed, int *status) {
    if (!ctx || fmt < B2F_FMT_DEFLATE || fmt > B2F_FMT_GZIP_MULTI) return B2F_ERR_INVALID_ARG;
    if (n_streams == 0) return B2F_OK;
    if (!in || !in_len || !out || !out_cap || !out_len || !in_consumed || !status) { ctx->err = "NULL argument"; return B2F_ERR_INVALID_ARG; }
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    std::vector<uint64_t> in_off(n_streams), out_off(n_streams); uint64_t tin = 0, tout = 0;
    for (size_t s = 0; s < n_streams; s++) { in_off[s] = tin; tin += align_up(in_len[s] + 16, 256); out_off[s] = tout; tout += align_up(out_cap[s] + 16, 256); }
    CK(ctx->buf[NB_IN].ensure(tin + 512));
    CK(ctx->buf[NB_DEC_OUT].ensure(tout + 512));
    uint8_t *d_in = ctx->buf[NB_IN].as<uint8_t>(), *d_out = ctx->buf[NB_DEC_OUT].as<uint8_t>();
    ctx->tm.mark(ctx->stream, "h2d");
    bool all_pinned = true; uint64_t in_bytes = 0;
    for (size_t s = 0; s < n_streams; s++) if (in_len[s]) { in_bytes += in_len[s]; all_pinned = all_pinned && is_pinned_host(in[s]); }
    ctx->feed_n = 0;
    if (ctx->overlap && all_pinned && in_bytes >= (8u << 20)) {
        // page-locked input: the copy runs in pieces on its own stream and the block finder follows it piece by piece
        const uint64_t piece = std::max<uint64_t>(4u << 20, (in_bytes + b2f_ctx::kFeedMax - 2) / (b2f_ctx::kFeedMax - 1));
        uint64_t acc = 0;
        for (size_t s = 0; s < n_streams; s++) {
            for (uint64_t a = 0; a < in_len[s]; ) {
                const uint64_t len = std::min<uint64_t>(in_len[s] - a, piece - acc);
                CK(cudaMemcpyAsync(d_in + in_off[s] + a, in[s] + a, len, cudaMemcpyHostToDevice, ctx->copy_st));
                a += len; acc += len;
                if (acc >= piece && ctx->feed_n + 1 < b2f_ctx::kFeedMax) {
                    ctx->feed_end[ctx->feed_n] = in_off[s] + a;
                    CK(cudaEventRecord(ctx->feed_ev[ctx->feed_n++], ctx->copy_st));
                    acc = 0;
                }
            }
        }
        ctx->feed_end[ctx->feed_n] = tin;
        CK(cudaEventRecord(ctx->feed_ev[ctx->feed_n++], ctx->copy_st));
    } else {
        for (size_t s = 0; s < n_streams; s++) if (in_len[s]) CK(h2d_copy(ctx, d_in + in_off[s], in[s], in_len[s], ctx->stream, is_pinned_host(in[s])));
    }
    InputAccess IA = { ctx, in, d_in, in_off.data(), in_len };
    // finished parts of the output are copied out while the rest is still being resolved (inflate_round)
    std::vector<uint8_t *> h_out(n_streams, nullptr); std::vector<char> h_pinned(n_streams, 0);
    for (size_t s = 0; s < n_streams; s++) { h_out[s] = out[s]; h_pinned[s] = is_pinned_host(out[s]) ? 1 : 0; }
    std::vector<uint64_t> h_copied(n_streams, 0);
    int rc = decode_core(ctx, fmt, n_streams, IA, d_in, in_off.data(), in_len, d_out, out_off.data(), out_cap, out_len, in_consumed, status, h_out.data(), &h_copied, &h_pinned);
    if (ctx->feed_n) { ctx->feed_n = 0; cudaStreamSynchronize(ctx->copy_st); }      // nothing on the device consumed the input (early error)
    if (rc) { cudaStreamSynchronize(ctx->copy_st); return rc; }
    ctx->tm.finish(ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    collect_stats(ctx, true);
    for (size_t s = 0; s < n_streams; s++) {
        const size_t w = std::min(out_len[s], out_cap[s]), have = (size_t)std::min<uint64_t>(h_copied[s], w);
        if (w > have) CK(d2h_copy(ctx, out[s] + have, d_out + out_off[s] + have, w - have, ctx->stream, h_pinned[s] != 0));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->aux[0]));
    return B2F_OK;
}

extern "C" int b2f_decode_device(b2f_ctx *ctx, int fmt, size_t n_streams, const uint8_t *d_in, const uint64_t *in_off, const size_t *in_len,
                                 uint8_t *d_out, const uint64_t *out_off, const size_t *out_cap, size_t *out_len, size_t *in_consumed, int *status) {
    if (!ctx || fmt < B2F_FMT_DEFLATE || fmt > B2F_FMT_GZIP_MULTI) return B2F_ERR_INVALID_ARG;
    if (n_streams == 0) return B2F_OK;
    if (!d_in || !in_off || !in_len || !d_out || !out_off || !out_cap || !out_len || !in_consumed || !status) { ctx->err = "NULL argument"; return B2F_ERR_INVALID_ARG; }
    CK(cudaSetDevice(ctx->device));
    ctx->tm.reset();
    InputAccess IA = { ctx, nullptr, d_in, in_off, in_len };
    int rc = decode_core(ctx, fmt, n_streams, IA, d_in, in_off, in_len, d_out, out_off, out_cap, out_len, in_consumed, status);
    if (rc) return rc;
    ctx->tm.finish(ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    collect_stats(ctx, true);
    return B2F_OK;
}

// ------------------------------------------------------------------------------------------ streaming handles
struct b2f_encoder { b2f_ctx *ctx; int fmt; b2f_encode_opts opts; std::vector<uint8_t> data; std::vector<int64_t> sched; std::vector<uint8_t> out; bool finished; std::string name, comment; std::vector<uint8_t> extra; };
struct b2f_decoder { b2f_ctx *ctx; int fmt; std::vector<uint8_t> in; std::vector<uint8_t> out; size_t rd; size_t consumed; int status; bool decoded; size_t good; };

extern "C" int b2f_encoder_new(b2f_ctx *ctx, int fmt, const b2f_encode_opts *opts, b2f_encoder **out) {
    if (!ctx || !out) return B2F_ERR_INVALID_ARG;
    b2f_encoder *e = new b2f_encoder();
    e->ctx = ctx; e->fmt = fmt; e->finished = false;
    if (opts) e->opts = *opts; else b2f_encode_opts_default(&e->opts);
    if (e->opts.gzip_filename) { e->name = e->opts.gzip_filename; e->opts.gzip_filename = e->name.c_str(); }
    if (e->opts.gzip_comment) { e->comment = e->opts.gzip_comment; e->opts.gzip_comment = e->comment.c_str(); }
    if (e->opts.gzip_has_extra) { e->extra.assign(e->opts.gzip_extra, e->opts.gzip_extra + e->opts.gzip_extra_len); e->opts.gzip_extra = e->extra.data(); }
    int rc = validate_opts(ctx, fmt, e->opts);
    if (rc) { delete e; return rc; }
    *out = e; return B2F_OK;
}
extern "C" int b2f_encoder_write(b2f_encoder *e, const uint8_t *buf, size_t len) {
    if (!e || e->finished) return B2F_ERR_INVALID_ARG;
    e->data.insert(e->data.end(), buf, buf + len); e->sched.push_back((int64_t)len); return B2F_OK;
}
extern "C" int b2f_encoder_flush(b2f_encoder *e) { if (!e || e->finished) return B2F_ERR_INVALID_ARG; e->sched.push_back(B2F_SCHED_FLUSH); return B2F_OK; }
extern "C" int b2f_encoder_finish(b2f_encoder *e, const uint8_t **out, size_t *out_len) {
    if (!e || !out || !out_len) return B2F_ERR_INVALID_ARG;
    if (!e->finished) {
        size_t cap = b2f_encode_bound(e->data.size(), e->sched.size(), &e->opts);
        e->out.resize(cap);
        const uint8_t *in = e->data.data(); size_t n = e->data.size(); const int64_t *sc = e->sched.data(); size_t ns = e->sched.size();
        static const int64_t none = 0;
        if (!ns) { sc = &none; }                         // zero writes: an explicit empty schedule (not "one write_all")
        uint8_t *op = e->out.data(); size_t ol = 0; int st = 0;
        size_t ns_eff = ns;
        int rc = b2f_encode_batch(e->ctx, e->fmt, &e->opts, 1, &in, &n, &sc, &ns_eff, &op, &cap, &ol, &st);
        if (rc) return rc;
        if (st) return st;
        e->out.resize(ol); e->finished = true;
    }
    *out = e->out.data(); *out_len = e->out.size();
    return B2F_OK;
}
extern "C" void b2f_encoder_free(b2f_encoder *e) { delete e; }

extern "C" int b2f_decoder_new(b2f_ctx *ctx, int fmt, const uint8_t *in, size_t in_len, b2f_decoder **out) {
    if (!ctx || !out) return B2F_ERR_INVALID_ARG;
    b2f_decoder *d = new b2f_decoder();
    d->ctx = ctx; d->fmt = fmt; d->in.assign(in, in + in_len); d->rd = 0; d->consumed = 0; d->status = 0; d->decoded = false;
    *out = d; return B2F_OK;
}
static int decoder_run(b2f_decoder *d) {
    size_t cap = std::max<size_t>(1 << 16, d->in.size() * 8 + 4096);
    for (;;) {
        d->out.resize(cap);
        const uint8_t *in = d->in.data(); size_t n = d->in.size(); uint8_t *op = d->out.data(); size_t ol = 0, ic = 0; int st = 0;
        int rc = b2f_decode_batch(d->ctx, d->fmt, 1, &in, &n, &op, &cap, &ol, &ic, &st);
        if (rc) return rc;
        if (st == B2F_ERR_OUTPUT_TOO_SMALL) { cap = std::max(cap * 2, ol + 64); continue; }
        d->out.resize(std::min(ol, cap)); d->consumed = ic; d->status = st; d->decoded = true;
        d->good = st == B2F_OK ? d->out.size() : std::min<size_t>(d->out.size(), d->ctx->last_good.empty() ? 0 : (size_t)d->ctx->last_good[0]);
        return B2F_OK;
    }
}
extern "C" int64_t b2f_decoder_read(b2f_decoder *d, uint8_t *buf, size_t len) {
    if (!d) return B2F_ERR_INVALID_ARG;
    if (!d->decoded) { int rc = decoder_run(d); if (rc) return rc; }
    // Decoder::read (decode.rs:136-164) decodes block by block: the bytes of every block that completed are handed out by read();
    // the error surfaces when the failing block is reached, and that block's partial bytes stay in unread_decoded_data()
    if (d->status != B2F_OK && d->rd >= d->good) return d->status;
    size_t k = std::min(len, (d->status != B2F_OK ? d->good : d->out.size()) - d->rd);
    memcpy(buf, d->out.data() + d->rd, k); d->rd += k;
    return (int64_t)k;
}
extern "C" size_t b2f_decoder_unread(const b2f_decoder *d, const uint8_t **ptr) { if (!d || !d->decoded) return 0; if (ptr) *ptr = d->out.data() + d->rd; return d->out.size() - d->rd; }
extern "C" size_t b2f_decoder_consumed(const b2f_decoder *d) { return d ? d->consumed : 0; }
extern "C" void b2f_decoder_free(b2f_decoder *d) { delete d; }

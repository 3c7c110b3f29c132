// abi_order.cpp -- drives libb2f.so through include/b2f.h in EXACTLY the order the Rust shim does
// (rust/libflate_b200/src/{lz77,stream}.rs), from C++ because the image has no Rust toolchain:
//   1. trait path  B200Lz77Encoder: encode(buf) appends and flushes when >= 8 * window bytes are buffered
//      (libflate_lz77/src/default.rs:60-68), flush() = one b2f_lz77_default call, codes replayed into a sink;
//      every chunk's codes must equal the oracle's DefaultLz77Encoder restatement (orc_lz77_default);
//   2. stream path  Encoder<W>: write() records sizes, flush() records B2F_SCHED_FLUSH, finish() = b2f_encode_bound +
//      ONE b2f_encode_batch; bytes must equal the oracle's (orc_encode) for the same schedule;
//      Decoder<R>: b2f_decode_batch with the retry-on-OUTPUT_TOO_SMALL loop.
// Test infrastructure: links the oracle as the checker.  Exit code 0 = all equal.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../include/b2f.h"

extern "C" {
int orc_lz77_default(const uint8_t *buf, size_t n, uint32_t window, uint32_t max_len, uint32_t *codes, size_t *n_codes);
int orc_encode(int fmt, const void *opts, const uint8_t *in, size_t n, const int64_t *sched, size_t n_sched, uint8_t *out, size_t cap, size_t *out_len);
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static std::vector<uint8_t> make_text(size_t n) {
    std::vector<std::vector<uint8_t>> words(700);
    for (auto &w : words) { size_t l = 2 + rnd() % 9; for (size_t i = 0; i < l; i++) w.push_back((uint8_t)("abcdefghijklmnopqrstuvwxyz_ABC"[rnd() % 30])); }
    std::vector<uint8_t> t;
    while (t.size() < n) { auto &w = words[rnd() % words.size()]; t.insert(t.end(), w.begin(), w.end()); t.push_back(rnd() % 4 ? '_' : '\n'); }
    t.resize(n);
    return t;
}

// ---- 1. the trait path ------------------------------------------------------------------------------------------------
struct Shim {                                   // rust/libflate_b200/src/lz77.rs
    b2f_ctx *ctx; uint32_t window = 32768, max_length = 258;
    std::vector<uint8_t> buf; std::vector<uint32_t> codes;
    std::vector<uint32_t> sink;                 // what Sink::consume received
    size_t flushes = 0; int mismatches = 0;
    void encode(const uint8_t *p, size_t n) { buf.insert(buf.end(), p, p + n); if (buf.size() >= (size_t)window * 8) flush(); }
    void flush() {
        if (buf.empty()) return;
        codes.resize(buf.size());
        size_t n = 0;
        int rc = b2f_lz77_default(ctx, buf.data(), buf.size(), window, max_length, codes.data(), &n);
        if (rc != B2F_OK) { fprintf(stderr, "b2f_lz77_default rc=%d: %s\n", rc, b2f_last_error(ctx)); exit(2); }
        std::vector<uint32_t> want(buf.size()); size_t wn = 0;
        orc_lz77_default(buf.data(), buf.size(), window, max_length, want.data(), &wn);
        if (wn != n || memcmp(want.data(), codes.data(), n * 4) != 0) { fprintf(stderr, "chunk %zu: codes differ (%zu vs %zu)\n", flushes, n, wn); mismatches++; }
        sink.insert(sink.end(), codes.begin(), codes.begin() + n);
        buf.clear(); flushes++;
    }
};

int main() {
    b2f_ctx *ctx = nullptr;
    if (b2f_ctx_create(0, &ctx) != B2F_OK) { fprintf(stderr, "no CUDA device: %s\n", b2f_last_error(nullptr)); return 3; }
    int bad = 0;
    const std::vector<uint8_t> text = make_text(1500000);
    {
        // Block::write feeds the Lz77Encode in the caller's write sizes (encode.rs:277-286); three schedules
        const size_t scheds[3] = { 8192, 70001, 1500000 };
        for (size_t ws : scheds) {
            Shim s; s.ctx = ctx;
            for (size_t o = 0; o < text.size(); o += ws) s.encode(text.data() + o, std::min(ws, text.size() - o));
            s.flush();
            bad += s.mismatches;
            printf("trait path, %zu-byte writes: %zu flushes, %zu codes, %d mismatching chunks\n", ws, s.flushes, s.sink.size(), s.mismatches);
        }
    }
    // ---- 2. the stream path ----------------------------------------------------------------------------------------------
    {
        b2f_encode_opts o; b2f_encode_opts_default(&o); o.gzip_mtime = 7;
        for (int fmt = B2F_FMT_DEFLATE; fmt <= B2F_FMT_GZIP; fmt++) {
            std::vector<int64_t> sched; size_t pos = 0;                      // Encoder::write x k, flush() twice in between
            while (pos < text.size()) { size_t w = std::min<size_t>(50000 + rnd() % 100000, text.size() - pos); sched.push_back((int64_t)w); pos += w; if (sched.size() % 5 == 0) sched.push_back(B2F_SCHED_FLUSH); }
            const uint8_t *ip = text.data(); size_t il = text.size(); const int64_t *sp = sched.data(); size_t sn = sched.size();
            size_t cap = b2f_encode_bound(il, sn, &o);
            std::vector<uint8_t> out(cap), want(cap); uint8_t *op = out.data(); size_t n = 0; int st = 0;
            int rc = b2f_encode_batch(ctx, fmt, &o, 1, &ip, &il, &sp, &sn, &op, &cap, &n, &st);
            size_t wn = 0;
            orc_encode(fmt, &o, text.data(), text.size(), sched.data(), sched.size(), want.data(), cap, &wn);
            const bool same = rc == 0 && st == 0 && n == wn && memcmp(out.data(), want.data(), n) == 0;
            printf("stream path fmt %d: %zu bytes, %s the oracle\n", fmt, n, same ? "equal to" : "DIFFERS from");
            bad += same ? 0 : 1;
            // Decoder: deliberately small first buffer -> OUTPUT_TOO_SMALL -> retry with out_len
            size_t dcap = 1000, ol = 0, ic = 0; std::vector<uint8_t> dec;
            for (;;) {
                dec.resize(dcap); uint8_t *dp = dec.data(); const uint8_t *ep = out.data();
                rc = b2f_decode_batch(ctx, fmt, 1, &ep, &n, &dp, &dcap, &ol, &ic, &st);
                if (rc == 0 && st == B2F_ERR_OUTPUT_TOO_SMALL) { dcap = ol + 64; continue; }
                break;
            }
            const bool rt = rc == 0 && st == 0 && ol == text.size() && ic == n && memcmp(dec.data(), text.data(), ol) == 0;
            printf("decoder fmt %d: %s\n", fmt, rt ? "round trip ok" : "MISMATCH");
            bad += rt ? 0 : 1;
        }
    }
    b2f_ctx_destroy(ctx);
    printf(bad ? "FAILED (%d)\n" : "abi order ok\n", bad);
    return bad ? 1 : 0;
}

// inflate_core.cuh -- DEFLATE block decoding (header parse, decode tables, symbol loop) as plain
// host/device-portable code.  On the device every lane of a warp runs this scalar code uniformly
// (loads broadcast, shared-memory tables), and the Out policy spreads LZ77 copies over the lanes;
// tests/native compiles the same source for the host to compare it with the oracle without a GPU.
//
// Behaviour follows libflate's decoder, including its error kinds and the points where they surface:
//   block loop / stored blocks   src/deflate/decode.rs:81-165
//   dynamic header               src/deflate/symbol.rs:387-484
//   code tables                  src/huffman.rs:35-55, 96-132  ("Bit region conflict" == over-subscribed code)
//   symbol decode                src/huffman.rs:157-179, src/deflate/symbol.rs:193-243
//   bit reader + deferred EOF    src/bit.rs:84-141
//   LZ77 copy + distance check   libflate_lz77/src/lib.rs:164-194
// The table layout (11/9-bit primary LUT + canonical slow path) is this implementation's own; the reference
// uses one flat 2^maxbits LUT.  Both decode the same prefix code, see DESIGN.md "decode tables".
#pragma once
#include "common.cuh"

namespace b2f {

constexpr int kInfOk = 0;
constexpr int kInfInvalid = -1;        // io::ErrorKind::InvalidData
constexpr int kInfEof = -2;            // io::ErrorKind::UnexpectedEof
constexpr int kInfOutFull = -3;
constexpr int kInfAbort = -4;           // probe ran past its stop bit (block-parallel candidate validation only)

constexpr uint32_t kLitBits = 11, kDistBits = 9;

// primary entry: [0:4) code width, [4:6) kind, payload from bit 8
constexpr uint32_t kKindLit = 0, kKindEob = 1, kKindLen = 2, kKindSpecial = 3;
constexpr uint32_t kSpecInvalid = 0, kSpecLong = 1, kSpecBadSym = 2;

struct alignas(16) InflateTables {
    uint32_t lit[1u << kLitBits];
    uint32_t dist[1u << kDistBits];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t lit_first[16], lit_count[16], lit_off[16];
    uint16_t dist_first[16], dist_count[16], dist_off[16];
    uint8_t lit_maxbw, dist_maxbw, lit_safe, dist_safe;
    uint8_t widths[320 + 160];          // scratch for the code-length decode (may overshoot by one run)
    uint32_t pre[128];                  // 7-bit LUT for the code-length code: width | sym<<8, 0 = unassigned
    uint8_t pre_maxbw;
};

B2F_HD uint32_t len_base(uint32_t k) {       // LENGTH_TABLE (symbol.rs:22-52): k = code - 257
    if (k < 8) return 3 + k;
    if (k == 28) return 258;
    uint32_t eb = (k - 4) >> 2;
    return 3 + ((4 + (k & 3)) << eb);
}
B2F_HD uint32_t len_extra(uint32_t k) { return (k < 8 || k == 28) ? 0 : (k - 4) >> 2; }
B2F_HD uint32_t dist_base(uint32_t k) {      // DISTANCE_TABLE (symbol.rs:56-87)
    if (k < 4) return 1 + k;
    uint32_t eb = (k - 2) >> 1;
    return 1 + ((2 + (k & 1)) << eb);
}
B2F_HD uint32_t dist_extra(uint32_t k) { return k < 4 ? 0 : (k - 2) >> 1; }

B2F_HD uint32_t make_lit_entry(uint32_t sym, uint32_t w) {
    if (sym < 256) return w | (kKindLit << 4) | (sym << 8);
    if (sym == 256) return w | (kKindEob << 4);
    if (sym >= 286) return w | (kKindSpecial << 4) | (kSpecBadSym << 8) | ((sym - 286) << 12);
    uint32_t k = sym - 257;
    return w | (kKindLen << 4) | (len_base(k) << 8) | (len_extra(k) << 20);
}
B2F_HD uint32_t make_dist_entry(uint32_t sym, uint32_t w) {
    return w | (dist_base(sym) << 8) | (dist_extra(sym) << 24);
}

// ---------------------------------------------------------------------------------- bit reader
// Mirrors BitReader (bit.rs:53-141): LSB-first, bytes pulled on demand, reading past the end yields zeros and
// a deferred UnexpectedEof that surfaces at the next check.  After the first failed refill the reference behaves
// as if exactly one zero byte had been appended and every later refill is refused (peek returns 0).
struct BitIn {
    const uint8_t *p;        // stream start
    uint64_t limit;          // 8 * byte length
    uint64_t pos;            // bit position
    uint64_t bb; uint32_t bc; uint64_t next;   // bit buffer: bc valid bits, `next` = byte offset of the next 4-byte load
    uint64_t stop;           // probes give up when pos passes this bit (~0 = never)
    int err;                 // pending error (last_error), 0 = none
    bool eof;                // a refill has failed
};
B2F_HD uint32_t load_le32_guarded(const uint8_t *p, uint64_t off, uint64_t nbytes) {
    // bytes at or beyond nbytes read as zero; 4-byte aligned fast path when fully inside
    if (off + 4 <= nbytes && ((reinterpret_cast<uintptr_t>(p + off)) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p + off);
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4; k++) if (off + k < nbytes) v |= (uint32_t)p[off + k] << (8 * k);
    return v;
}
B2F_HD void bi_seek(BitIn &b, uint64_t bitpos) {
    b.pos = bitpos; b.next = bitpos >> 3;
    uint32_t drop = (uint32_t)(bitpos & 7);
    // start on a 4-byte aligned address so that every later refill is one aligned 32-bit load
    uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(b.p + b.next) & 3);
    if (mis && b.next >= mis) { b.next -= mis; drop += 8 * mis; }
    uint32_t w = load_le32_guarded(b.p, b.next, b.limit >> 3);
    b.bb = (uint64_t)w >> drop; b.bc = 32 - drop; b.next += 4;
}
B2F_HD void bi_init(BitIn &b, const uint8_t *p, uint64_t nbytes, uint64_t bitpos) {
    b.p = p; b.limit = nbytes * 8; b.err = 0; b.eof = false; b.stop = ~0ull;
    bi_seek(b, bitpos);
}
B2F_HD void bi_refill(BitIn &b) {            // keep >= 32 valid bits
    if (b.bc < 32) {
        uint32_t w = load_le32_guarded(b.p, b.next, b.limit >> 3);
        b.bb |= (uint64_t)w << b.bc; b.bc += 32; b.next += 4;
    }
}
// peek_bits_unchecked(k), k <= 16
B2F_HD uint32_t bi_peek(BitIn &b, uint32_t k) {
    if (b.pos + k > b.limit) {               // needs a byte that does not exist (k == 0 never does)
        if (k == 0) return 0;
        if (!b.eof) { b.eof = true; b.err = kInfEof; return 0; }          // the failing refill: error recorded, 0 returned
        if (b.pos + k > b.limit + 8) return 0;                              // refused refill
    }
    bi_refill(b);
    return (uint32_t)b.bb & ((1u << k) - 1u);
}
B2F_HD void bi_skip(BitIn &b, uint32_t k) {
    b.pos += k;
    if (k <= b.bc) { b.bb >>= k; b.bc -= k; }
    else { bi_seek(b, b.pos); }              // only reachable on error paths (skip without a successful peek)
}
B2F_HD uint32_t bi_read(BitIn &b, uint32_t k) { uint32_t v = bi_peek(b, k); bi_skip(b, k); return v; }
B2F_HD int bi_check(BitIn &b) { int e = b.err; b.err = 0; return e; }    // check_last_error(): take()

// ---------------------------------------------------------------------------------- table construction
// from_bitwidthes for a DecoderBuilder (huffman.rs:80-91): returns kInfInvalid when the widths are
// over-subscribed (the reference's "Bit region conflict"); incomplete codes are accepted.
// lane/nl: cooperative striding (0/1 on the host).  SYNC() must order shared-memory writes between phases.
template <class Sync>
B2F_HD int build_decode_table(const uint8_t *w, int n, bool is_lit, InflateTables &T, int lane, int nl, Sync SYNC) {
    uint32_t cnt[16];
    for (int i = 0; i < 16; i++) cnt[i] = 0;
    for (int s = 0; s < n; s++) cnt[w[s]]++;
    cnt[0] = 0;
    uint32_t maxbw = 0;
    for (uint32_t l = 1; l < 16; l++) if (cnt[l]) maxbw = l;
    // over-subscription check == a canonical code running out of its width
    uint32_t code = 0; uint32_t first[16], off[16]; uint32_t o = 0;
    first[0] = 0; off[0] = 0;
    for (uint32_t l = 1; l < 16; l++) {
        code = (code + cnt[l - 1]) << 1;
        first[l] = code; off[l] = o; o += cnt[l];
        if (cnt[l] && code + cnt[l] > (1u << l)) return kInfInvalid;
    }
    uint32_t *prim = is_lit ? T.lit : T.dist;
    const uint32_t PB = is_lit ? kLitBits : kDistBits;
    uint16_t *sorted = is_lit ? T.lit_sorted : T.dist_sorted;
    uint16_t *tf = is_lit ? T.lit_first : T.dist_first, *tc = is_lit ? T.lit_count : T.dist_count, *to = is_lit ? T.lit_off : T.dist_off;
    for (uint32_t i = (uint32_t)lane; i < (1u << PB); i += (uint32_t)nl) prim[i] = kKindSpecial << 4;   // invalid
    if (lane == 0) {
        for (uint32_t l = 0; l < 16; l++) { tf[l] = (uint16_t)first[l]; tc[l] = (uint16_t)cnt[l]; to[l] = (uint16_t)off[l]; }
        if (is_lit) T.lit_maxbw = (uint8_t)maxbw; else T.dist_maxbw = (uint8_t)maxbw;
    }
    SYNC();
    // symbols in (width, symbol) order; each lane handles a strided subset of symbols and computes its rank
    for (int s = lane; s < n; s += nl) {
        uint32_t l = w[s];
        if (!l) continue;
        uint32_t rank = 0;
        for (int t = 0; t < s; t++) rank += (w[t] == l);
        uint32_t c = first[l] + rank;
        sorted[off[l] + rank] = (uint16_t)s;
        uint32_t r = bitrev(c, l);
        if (l <= PB) {
            uint32_t e = is_lit ? make_lit_entry((uint32_t)s, l) : make_dist_entry((uint32_t)s, l);
            for (uint32_t k = r; k < (1u << PB); k += (1u << l)) prim[k] = e;
        } else {
            prim[r & ((1u << PB) - 1u)] = (kKindSpecial << 4) | (kSpecLong << 8);
        }
    }
    SYNC();
    return kInfOk;
}

// one table lookup on `bits` (>= 15 valid bits, upper bits may be zero extended): returns entry; width 0 + special => invalid
B2F_HD uint32_t lookup_code(const InflateTables &T, bool is_lit, uint32_t bits) {
    const uint32_t PB = is_lit ? kLitBits : kDistBits;
    uint32_t e = (is_lit ? T.lit : T.dist)[bits & ((1u << PB) - 1u)];
    if ((e & 0xF) != 0 || ((e >> 4) & 3) != kKindSpecial) return e;
    if (((e >> 8) & 0xF) != kSpecLong) return e;                 // invalid
    // canonical slow path for codes longer than the primary index
    const uint16_t *tf = is_lit ? T.lit_first : T.dist_first, *tc = is_lit ? T.lit_count : T.dist_count, *to = is_lit ? T.lit_off : T.dist_off;
    const uint16_t *sorted = is_lit ? T.lit_sorted : T.dist_sorted;
    uint32_t v = bitrev(bits & 0x7FFFu, 15);                     // MSB-first view of the next 15 bits
    uint32_t mb = is_lit ? T.lit_maxbw : T.dist_maxbw;
    for (uint32_t l = PB + 1; l <= mb; l++) {
        uint32_t c = v >> (15 - l);
        uint32_t d = c - tf[l];
        if (c >= tf[l] && d < tc[l]) { uint32_t s = sorted[to[l] + d]; return is_lit ? make_lit_entry(s, l) : make_dist_entry(s, l); }
    }
    return kKindSpecial << 4;                                    // invalid
}

// huffman::Decoder::decode_unchecked (huffman.rs:157-179) with the reference's two-step peek near the end of input.
B2F_HD uint32_t decode_code(BitIn &b, const InflateTables &T, bool is_lit) {
    const uint32_t mb = is_lit ? T.lit_maxbw : T.dist_maxbw;
    const uint32_t safe = is_lit ? T.lit_safe : T.dist_safe;
    if (b.pos + 64 <= b.limit && !b.eof) {                       // far from the end: a full-width lookup is equivalent
        bi_refill(b);
        uint32_t e = lookup_code(T, is_lit, (uint32_t)b.bb & 0x7FFFu);
        uint32_t wdt = e & 0xF;
        if (wdt == 0) { b.err = kInfInvalid; wdt = 16; }         // "Invalid huffman coded stream": skips the table's 16
        bi_skip(b, wdt);
        return e;
    }
    uint32_t peek = safe, e, wdt;
    for (;;) {
        uint32_t code = bi_peek(b, peek);
        e = mb ? lookup_code(T, is_lit, code) : (kKindSpecial << 4);
        wdt = e & 0xF;
        if (wdt == 0) wdt = 16;                                  // unassigned pattern: table value 16
        if (wdt <= peek) break;
        if (wdt > mb) { b.err = kInfInvalid; break; }
        peek = wdt;
    }
    bi_skip(b, wdt);
    return e;
}

// DynamicHuffmanCodec::load (symbol.rs:387-456).  Uniform scalar code; table fills are strided over lanes.
template <class Sync>
B2F_HD int load_dynamic(BitIn &b, InflateTables &T, int lane, int nl, Sync SYNC) {
    const uint8_t ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
    int rc;
    uint32_t hlit = bi_read(b, 5); if ((rc = bi_check(b))) return rc;
    uint32_t hdist = bi_read(b, 5); if ((rc = bi_check(b))) return rc;
    uint32_t hclen = bi_read(b, 4); if ((rc = bi_check(b))) return rc;
    hlit += 257; hdist += 1; hclen += 4;
    if (hdist > 30) return kInfInvalid;                          // "The value of HDIST is too big"
    uint8_t pw[19];
    for (int i = 0; i < 19; i++) pw[i] = 0;
    for (uint32_t k = 0; k < hclen; k++) { uint32_t x = bi_read(b, 3); if ((rc = bi_check(b))) return rc; pw[ORDER[k]] = (uint8_t)x; }
    // code-length code: DecoderBuilder::from_bitwidthes(.., Some(1), None)
    uint32_t cnt[8]; for (int i = 0; i < 8; i++) cnt[i] = 0;
    for (int i = 0; i < 19; i++) cnt[pw[i]]++;
    cnt[0] = 0;
    uint32_t pmax = 0; for (uint32_t l = 1; l < 8; l++) if (cnt[l]) pmax = l;
    uint32_t first[8]; uint32_t code = 0; first[0] = 0;
    for (uint32_t l = 1; l < 8; l++) { code = (code + cnt[l - 1]) << 1; first[l] = code; if (cnt[l] && code + cnt[l] > (1u << l)) return kInfInvalid; }
    SYNC();
    for (uint32_t i = (uint32_t)lane; i < 128; i += (uint32_t)nl) T.pre[i] = 0;
    SYNC();
    if (lane == 0) {
        uint32_t nxt[8]; for (int l = 0; l < 8; l++) nxt[l] = first[l];
        for (uint32_t s = 0; s < 19; s++) {
            uint32_t l = pw[s]; if (!l) continue;
            uint32_t r = bitrev(nxt[l]++, l);
            for (uint32_t k = r; k < 128; k += (1u << l)) T.pre[k] = l | (s << 8);
        }
        T.pre_maxbw = (uint8_t)pmax;
    }
    SYNC();
    const uint32_t psafe = pmax < 1 ? pmax : 1;                  // min(max_bitwidth, Some(1))
    // literal/length widths then distance widths; runs may spill from the first list into the second (symbol.rs:422-424)
    uint32_t total = 0, want = hlit;
    int phase = 0;                                               // 0: literal list, 1: distance list
    uint32_t nl_done = 0;
    for (;;) {
        if (phase == 0 && total >= hlit) { phase = 1; nl_done = hlit; want = hlit + hdist; }
        if (phase == 1 && total >= want) break;
        // bitwidth_decoder.decode(reader)?  -- same two-step peek as any huffman::Decoder
        uint32_t sym, wdt, peek = psafe;
        for (;;) {
            uint32_t c = bi_peek(b, peek);
            uint32_t e = pmax ? T.pre[c & 127] : 0;
            wdt = e & 0xFF; sym = e >> 8;
            if (wdt == 0) wdt = 16;
            if (wdt <= peek) break;
            if (wdt > pmax) { b.err = kInfInvalid; sym = 0; break; }
            peek = wdt;
        }
        bi_skip(b, wdt);
        if ((rc = bi_check(b))) return rc;
        // load_bitwidthes (symbol.rs:457-484)
        uint32_t rep = 1; uint8_t val = (uint8_t)sym;
        if (sym == 16) {
            rep = bi_read(b, 2) + 3; if ((rc = bi_check(b))) return rc;
            // `last`: previous entry of the current list, or (distance list only) the last literal width
            bool have = phase == 0 ? total > 0 : (total > nl_done || nl_done > 0);
            if (!have) return kInfInvalid;                       // "No preceding value"
            val = T.widths[total - 1];
        } else if (sym == 17) { rep = bi_read(b, 3) + 3; if ((rc = bi_check(b))) return rc; val = 0; }
        else if (sym == 18) { rep = bi_read(b, 7) + 11; if ((rc = bi_check(b))) return rc; val = 0; }
        for (uint32_t k = 0; k < rep; k++) T.widths[total + k] = val;    // uniform value from every lane
        total += rep;
    }
    if (total > hlit + hdist) return kInfInvalid;                // "distance_code_bitwidthes is too large"
    SYNC();
    if ((rc = build_decode_table(T.widths, (int)hlit, true, T, lane, nl, SYNC))) return rc;
    if ((rc = build_decode_table(T.widths + hlit, (int)hdist, false, T, lane, nl, SYNC))) return rc;
    if (lane == 0) {
        uint32_t eobw = hlit > 256 ? T.widths[256] : 0;          // safely_peek_bitwidth (huffman.rs:98-100, 124-131)
        uint32_t ls = eobw ? eobw : 1; if (ls > T.lit_maxbw) ls = T.lit_maxbw;
        uint32_t ds = ls; if (ds > T.dist_maxbw) ds = T.dist_maxbw;
        T.lit_safe = (uint8_t)ls; T.dist_safe = (uint8_t)ds;
    }
    SYNC();
    return kInfOk;
}

// FixedHuffmanCodec::load (symbol.rs:290-315)
template <class Sync>
B2F_HD int load_fixed(InflateTables &T, int lane, int nl, Sync SYNC) {
    SYNC();
    for (int s = lane; s < 288; s += nl) T.widths[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
    for (int s = lane; s < 30; s += nl) T.widths[288 + s] = 5;
    SYNC();
    build_decode_table(T.widths, 288, true, T, lane, nl, SYNC);
    build_decode_table(T.widths + 288, 30, false, T, lane, nl, SYNC);
    if (lane == 0) { T.lit_safe = 7; T.dist_safe = 5; }         // EOB is 7 bits wide; min(5, 7)
    SYNC();
    return kInfOk;
}

// Result of decoding from some bit position up to (and including) the final block or an error.
struct InflateResult {
    int status;              // kInfOk / kInfInvalid / kInfEof / kInfOutFull
    uint64_t out_len;        // bytes produced (including the partial block before an error)
    uint64_t end_bit;        // bit position after the last consumed bit
    uint64_t consumed;       // bytes pulled from the underlying reader
    uint32_t final_seen;     // BFINAL block completed
};

// Decode blocks starting at b.pos until BFINAL (max_blocks limits the count: block-parallel callers pass 1).
// `hist_base`: number of bytes of history that precede out position 0 (for the "Too long backword reference" check).
// Out policy: lit(pos, byte), copy(pos, len, dist), raw(pos, src, n), cap(), block_end(start_bit, end_bit, out_pos, final),
// fast(b, T, out_pos, hist_base): optional accelerated symbol loop (returns 1 when it consumed EndOfBlock, else 0 and
// leaves the reader at a symbol boundary for the exact generic path).
template <class Out, class Sync>
B2F_HD void inflate_blocks(BitIn &b, InflateTables &T, Out &out, uint64_t out_pos, uint64_t hist_base, uint32_t max_blocks,
                           int lane, int nl, Sync SYNC, InflateResult &R) {
    R.status = kInfOk; R.final_seen = 0;
    uint64_t pulled = (b.pos + 7) >> 3;
    const uint64_t cap = out.cap();
    for (uint32_t nb = 0; nb < max_blocks; nb++) {
        int rc;
        const uint64_t blk_start = b.pos;
        uint32_t bfinal = bi_read(b, 1); if ((rc = bi_check(b))) { R.status = rc; break; }
        uint32_t btype = bi_read(b, 2); if ((rc = bi_check(b))) { R.status = rc; break; }
        if (btype == 0) {                                        // read_non_compressed_block (decode.rs:81-111)
            uint64_t byte = (b.pos + 7) >> 3, nbytes = b.limit >> 3;
            if (byte + 2 > nbytes) { R.status = kInfEof; b.pos = nbytes * 8; break; }
            uint32_t len = (uint32_t)b.p[byte] | ((uint32_t)b.p[byte + 1] << 8);
            if (byte + 4 > nbytes) { R.status = kInfEof; b.pos = nbytes * 8; break; }
            uint32_t nlen = (uint32_t)b.p[byte + 2] | ((uint32_t)b.p[byte + 3] << 8);
            if (((~len) & 0xFFFFu) != nlen) { R.status = kInfInvalid; b.pos = (byte + 4) * 8; break; }
            uint64_t avail = nbytes - (byte + 4), used = avail < len ? avail : len;
            uint64_t room = cap - out_pos, wr = used < room ? used : room;
            out.raw(out_pos, b.p + byte + 4, wr);
            out_pos += wr;
            if (wr < used) { R.status = kInfOutFull; out_pos += used - wr; }
            bi_seek(b, (byte + 4 + used) * 8);
            if (used != len) { R.status = kInfEof; break; }
            if (R.status) break;
        } else if (btype == 3) { R.status = kInfInvalid; break; }   // "btype 0x11 of DEFLATE is reserved(error) value"
        else {
            rc = btype == 1 ? load_fixed(T, lane, nl, SYNC) : load_dynamic(b, T, lane, nl, SYNC);
            if (rc) { R.status = rc; break; }
            for (;;) {                                           // read_compressed_block loop (decode.rs:117-128)
                if (out.fast(b, T, out_pos, hist_base)) break;   // device fast path: returns 1 after consuming EndOfBlock
                uint32_t e = decode_code(b, T, true);
                uint32_t kind = (e >> 4) & 3;
                uint32_t length = 0, distance = 0;
                if (kind == kKindSpecial) {
                    if ((e & 0xF) != 0 && ((e >> 8) & 0xF) == kSpecBadSym) b.err = kInfInvalid;   // 286/287 "must not occur"
                    kind = (e & 0xF) == 0 ? kKindLit : kKindEob;                                   // dummy values of the reference
                } else if (kind == kKindLen) {
                    uint32_t eb = (e >> 20) & 0xF;
                    length = ((e >> 8) & 0x1FF) + bi_read(b, eb);
                    uint32_t d = decode_code(b, T, false);
                    uint32_t deb = (d >> 24) & 0xF;
                    distance = ((d >> 8) & 0xFFFF) + bi_read(b, deb);
                }
                if ((rc = bi_check(b))) { R.status = rc; break; }
                if (b.pos > b.stop) { R.status = kInfAbort; break; }
                if (kind == kKindEob) break;
                if (kind == kKindLit) {
                    if (out_pos < cap) out.lit(out_pos, (uint8_t)(e >> 8)); else R.status = kInfOutFull;
                    out_pos++;
                } else {
                    if (out_pos + hist_base < distance) { R.status = kInfInvalid; break; }         // "Too long backword reference"
                    if (out_pos + length <= cap) out.copy(out_pos, length, distance);
                    else {
                        if (out_pos < cap) out.copy(out_pos, (uint32_t)(cap - out_pos), distance);
                        R.status = kInfOutFull;
                    }
                    out_pos += length;
                }
            }
            if (R.status && R.status != kInfOutFull) break;
        }
        uint64_t pb = (b.pos + 7) >> 3; if (pb > pulled) pulled = pb;
        out.block_end(blk_start, b.pos, out_pos, bfinal != 0);
        if (bfinal) { R.final_seen = 1; break; }
    }
    uint64_t pb = (b.pos + 7) >> 3; if (pb > (b.limit >> 3)) pb = b.limit >> 3;
    if (pb > pulled) pulled = pb;
    if (b.eof) pulled = b.limit >> 3;
    R.out_len = out_pos; R.end_bit = b.pos; R.consumed = pulled;
}

}  // namespace b2f

/*
 * oracle/flate_oracle.c  --  TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C, single-threaded CPU restatement of the sile/libflate (v2.3.0) DEFLATE
 * hot path, written from the reference's behaviour (citations are file:line under
 * /root/reference).  It exists so that the CUDA path in libflate_b200/ can be checked
 * byte-for-byte against "what libflate would have produced".  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (libflate_b200/csrc) never links, loads or calls it.
 *
 * Parity pin: this restatement reproduces every exact-bytes golden vector the
 * reference's own tests and doctests hold for this path (tests/test_oracle_goldens.py,
 * fixtures in tests/golden/, extracted by tests/golden/make_goldens.py).
 * The reference itself (Rust) cannot be compiled in this image (no rustc/cargo).
 *
 * The data structures deliberately mirror the reference's (65 536 small vectors as the
 * trigram table, symbol vectors, node-list package-merge, byte-at-a-time bit reader,
 * flat 2^maxbits decode LUT) so that timing it is a fair "libflate-restatement CPU"
 * baseline rather than a tuned competitor.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ORC_OK 0
#define ORC_INVALID_DATA (-1)
#define ORC_UNEXPECTED_EOF (-2)
#define ORC_OUTPUT_TOO_SMALL (-3)
#define ORC_NOMEM (-4)

#define ORC_FMT_DEFLATE 0
#define ORC_FMT_ZLIB 1
#define ORC_FMT_GZIP 2
#define ORC_FMT_GZIP_MULTI 3

#define ORC_MODE_DYNAMIC 0
#define ORC_MODE_FIXED 1
#define ORC_MODE_STORED 2

typedef struct {
    uint64_t block_size;      /* deflate::EncodeOptions::block_size, default 1<<20 (src/deflate/encode.rs:12) */
    uint32_t window_size;     /* DefaultLz77EncoderBuilder::window_size, default 32768 (libflate_lz77/src/default.rs:213) */
    uint32_t max_length;      /* DefaultLz77EncoderBuilder::max_length, default 258 */
    int32_t mode;             /* ORC_MODE_* : dynamic / fixed_huffman_codes() / no_compression() */
    int32_t zlib_flush_sync;  /* zlib::FlushMode::Sync (src/zlib.rs:150-157) */
    uint32_t gzip_mtime;      /* HeaderBuilder::modification_time */
    uint8_t gzip_os;          /* default 3 = Unix (src/gzip.rs:140) */
    uint8_t gzip_is_text;
    uint8_t gzip_is_verified; /* F_HCRC */
    uint8_t gzip_has_extra;
    const uint8_t *gzip_extra; /* raw subfield bytes (id,len,data)*; written after a u16 total length */
    uint32_t gzip_extra_len;
    const char *gzip_filename; /* NUL-terminated or NULL */
    const char *gzip_comment;  /* NUL-terminated or NULL */
} orc_opts;

/* ------------------------------------------------------------------ byte vector */
typedef struct { uint8_t *p; size_t len, cap; } bytevec;
static void bv_reserve(bytevec *v, size_t extra) {
    if (v->len + extra > v->cap) {
        size_t nc = v->cap ? v->cap : 64;
        while (nc < v->len + extra) nc *= 2;
        v->p = (uint8_t *)realloc(v->p, nc);
        v->cap = nc;
    }
}
static void bv_push(bytevec *v, uint8_t b) { bv_reserve(v, 1); v->p[v->len++] = b; }
static void bv_extend(bytevec *v, const uint8_t *s, size_t n) {
    if (!n) return;
    bv_reserve(v, n); memcpy(v->p + v->len, s, n); v->len += n;
}
static void bv_free(bytevec *v) { free(v->p); v->p = NULL; v->len = v->cap = 0; }

/* ------------------------------------------------------------------ symbol vector
 * code word layout (same as the C ABI, include/b2f.h): literal = byte value,
 * pointer = 0x80000000 | length<<16 | distance, end-of-block = 0x40000000. */
#define SYM_EOB 0x40000000u
#define SYM_PTR 0x80000000u
typedef struct { uint32_t *p; size_t len, cap; } symvec;
static void sv_push(symvec *v, uint32_t s) {
    if (v->len == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->p = (uint32_t *)realloc(v->p, v->cap * 4); }
    v->p[v->len++] = s;
}

/* ================================================================== LZ77 (libflate_lz77/src/default.rs) */
typedef struct { uint8_t key; uint32_t pos; } pt_ent;           /* (u8, u32), default.rs:156 */
typedef struct { pt_ent *e; uint32_t len, cap; } pt_vec;
typedef struct {
    int large;
    pt_vec *table;                 /* Large: 65536 vectors (default.rs:158-163) */
    uint32_t *skey; uint32_t *spos; /* Small: HashMap<[u8;3],u32> (default.rs:133-141); open addressing, map semantics only */
} prefix_table;

static void pt_new(prefix_table *t, size_t bytes) {
    memset(t, 0, sizeof *t);
    if (bytes < 32768) {            /* default.rs:138 */
        t->large = 0;
        t->skey = (uint32_t *)calloc(65536, 4);  /* key+1, 0 = empty */
        t->spos = (uint32_t *)malloc(65536 * 4);
    } else {
        t->large = 1;
        t->table = (pt_vec *)calloc(65536, sizeof(pt_vec));
    }
}
static void pt_free(prefix_table *t) {
    if (t->large) { for (int i = 0; i < 65536; i++) free(t->table[i].e); free(t->table); }
    else { free(t->skey); free(t->spos); }
}
/* returns previous position or -1 (default.rs:146-151, 166-183) */
static int64_t pt_insert(prefix_table *t, const uint8_t *p3, uint32_t pos) {
    if (t->large) {
        pt_vec *v = &t->table[((uint32_t)p3[0] << 8) + p3[1]];
        for (uint32_t k = 0; k < v->len; k++) {
            if (v->e[k].key == p3[2]) { uint32_t old = v->e[k].pos; v->e[k].pos = pos; return old; }
        }
        if (v->len == v->cap) { v->cap = v->cap ? v->cap * 2 : 4; v->e = (pt_ent *)realloc(v->e, v->cap * sizeof(pt_ent)); }
        v->e[v->len].key = p3[2]; v->e[v->len].pos = pos; v->len++;
        return -1;
    } else {
        uint32_t key = ((uint32_t)p3[0] << 16) | ((uint32_t)p3[1] << 8) | p3[2];
        uint32_t h = (key * 2654435761u) >> 16;
        for (;;) {
            if (t->skey[h] == 0) { t->skey[h] = key + 1; t->spos[h] = pos; return -1; }
            if (t->skey[h] == key + 1) { uint32_t old = t->spos[h]; t->spos[h] = pos; return old; }
            h = (h + 1) & 65535;
        }
    }
}

typedef struct {
    uint32_t window_size, max_length;
    bytevec buf;
} lz77_enc;

/* default.rs:122-129 : lcp of buf[i..] (at most max-3 items) against buf[j..] */
static uint32_t lz_lcp(const uint8_t *buf, size_t n, size_t i, size_t j, uint32_t max) {
    uint32_t k = 0, lim = max - 3;
    while (k < lim && i + k < n && buf[i + k] == buf[j + k]) k++;
    return k;
}

/* DefaultLz77Encoder::flush (default.rs:69-109) */
static void lz77_flush(lz77_enc *e, symvec *sink) {
    const uint8_t *buf = e->buf.p; size_t n = e->buf.len;
    prefix_table pt; pt_new(&pt, n);
    size_t i = 0, end = (n > 3 ? n : 3) - 3;
    while (i < end) {
        int64_t m = pt_insert(&pt, buf + i, (uint32_t)i);
        if (m >= 0) {
            size_t j = (size_t)m, distance = i - j;
            if (distance <= e->window_size) {
                uint32_t length = 3 + lz_lcp(buf, n, i + 3, j + 3, e->max_length);
                sv_push(sink, SYM_PTR | (length << 16) | (uint32_t)distance);
                for (size_t k = i + 1; k < i + length; k++) {
                    if (k >= end) break;
                    pt_insert(&pt, buf + k, (uint32_t)k);
                }
                i += length;
                continue;
            }
        }
        sv_push(sink, buf[i]);
        i++;
    }
    for (; i < n; i++) sv_push(sink, buf[i]);
    e->buf.len = 0;
    pt_free(&pt);
}
/* DefaultLz77Encoder::encode (default.rs:60-68) */
static void lz77_encode(lz77_enc *e, const uint8_t *b, size_t n, symvec *sink) {
    bv_extend(&e->buf, b, n);
    if (e->buf.len >= (size_t)e->window_size * 8) lz77_flush(e, sink);
}

/* ================================================================== bit writer (src/bit.rs:6-50) */
typedef struct { bytevec *inner; uint32_t buf; uint8_t end; } bitwriter;
static void bw_write_bits(bitwriter *w, uint8_t bitwidth, uint16_t bits) {
    w->buf |= (uint32_t)bits << w->end;
    w->end += bitwidth;
    if (w->end >= 16) {                       /* flush_if_needed, bit.rs:42-49 */
        bv_push(w->inner, (uint8_t)w->buf); bv_push(w->inner, (uint8_t)(w->buf >> 8));
        w->end -= 16; w->buf >>= 16;
    }
}
static void bw_flush(bitwriter *w) {          /* bit.rs:32-40 */
    while (w->end > 0) {
        bv_push(w->inner, (uint8_t)w->buf);
        w->buf >>= 8;
        w->end = w->end >= 8 ? w->end - 8 : 0;
    }
}

/* ================================================================== huffman (src/huffman.rs) */
typedef struct { uint8_t width; uint16_t bits; } hcode;
typedef struct { hcode *table; uint32_t n; } henc;   /* huffman::Encoder */

static uint16_t inverse_endian(uint8_t width, uint16_t bits) {   /* huffman.rs:19-28 */
    uint16_t f = bits, t = 0;
    for (int k = 0; k < width; k++) { t = (uint16_t)(t << 1); t |= f & 1; f >>= 1; }
    return t;
}

/* ordinary_huffman_codes::calc_optimal_max_bitwidth (huffman.rs:261-274).
 * Max-heap over tuples (weight = -freq, width); tuple total order makes the result container independent. */
typedef struct { int64_t w; uint8_t d; } hp_item;
static int hp_less(hp_item a, hp_item b) { return a.w < b.w || (a.w == b.w && a.d < b.d); }
static void hp_push(hp_item *h, int *n, hp_item x) {
    int i = (*n)++; h[i] = x;
    while (i > 0) { int p = (i - 1) / 2; if (hp_less(h[p], h[i])) { hp_item t = h[p]; h[p] = h[i]; h[i] = t; i = p; } else break; }
}
static hp_item hp_pop(hp_item *h, int *n) {
    hp_item top = h[0]; h[0] = h[--(*n)];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < *n && hp_less(h[m], h[l])) m = l;
        if (r < *n && hp_less(h[m], h[r])) m = r;
        if (m == i) break;
        hp_item t = h[m]; h[m] = h[i]; h[i] = t; i = m;
    }
    return top;
}
static uint8_t calc_optimal_max_bitwidth(const size_t *freq, int nf) {
    hp_item heap[320]; int n = 0;
    for (int i = 0; i < nf; i++) if (freq[i] > 0) { hp_item x = { -(int64_t)freq[i], 0 }; hp_push(heap, &n, x); }
    while (n > 1) {
        hp_item a = hp_pop(heap, &n), b = hp_pop(heap, &n);
        hp_item c = { a.w + b.w, (uint8_t)(1 + (a.d > b.d ? a.d : b.d)) };
        hp_push(heap, &n, c);
    }
    uint8_t mb = n ? heap[0].d : 0;
    return mb > 1 ? mb : 1;
}

/* length_limited_huffman_codes (huffman.rs:276-363): node lists carrying symbol vectors, as in the reference */
typedef struct { uint16_t *syms; uint32_t nsyms; size_t weight; } pm_node;
typedef struct { pm_node *v; int n; } pm_list;

static pm_node pm_clone(const pm_node *a) {
    pm_node r; r.nsyms = a->nsyms; r.weight = a->weight;
    r.syms = (uint16_t *)malloc((a->nsyms ? a->nsyms : 1) * 2);
    memcpy(r.syms, a->syms, a->nsyms * 2);
    return r;
}
static pm_list pm_clone_list(const pm_list *s) {
    pm_list r; r.n = s->n; r.v = (pm_node *)malloc((s->n ? s->n : 1) * sizeof(pm_node));
    for (int i = 0; i < s->n; i++) r.v[i] = pm_clone(&s->v[i]);
    return r;
}
static void pm_free_list(pm_list *l) { for (int i = 0; i < l->n; i++) free(l->v[i].syms); free(l->v); l->v = NULL; l->n = 0; }
/* package (huffman.rs:350-362): pair adjacent nodes, drop an odd tail; lists shorter than 2 are returned unchanged */
static pm_list pm_package(pm_list nodes) {
    if (nodes.n >= 2) {
        int new_len = nodes.n / 2;
        for (int i = 0; i < new_len; i++) {
            pm_node a = nodes.v[2 * i], b = nodes.v[2 * i + 1];
            pm_node m; m.weight = a.weight + b.weight; m.nsyms = a.nsyms + b.nsyms;
            m.syms = (uint16_t *)malloc((m.nsyms ? m.nsyms : 1) * 2);
            memcpy(m.syms, a.syms, a.nsyms * 2); memcpy(m.syms + a.nsyms, b.syms, b.nsyms * 2);
            free(a.syms); free(b.syms);
            nodes.v[i] = m;            /* slots 2i,2i+1 are consumed; i <= 2i so no live slot is clobbered */
        }
        if (nodes.n & 1) free(nodes.v[nodes.n - 1].syms);
        nodes.n = new_len;
    }
    return nodes;
}
/* merge (huffman.rs:329-349): take from x only when strictly lighter than the head of y */
static pm_list pm_merge(pm_list x, pm_list y) {
    pm_list z; z.n = 0; z.v = (pm_node *)malloc((x.n + y.n + 1) * sizeof(pm_node));
    int ix = 0, iy = 0;
    for (;;) {
        if (ix >= x.n) { while (iy < y.n) z.v[z.n++] = y.v[iy++]; break; }
        else if (iy >= y.n) { while (ix < x.n) z.v[z.n++] = x.v[ix++]; break; }
        else if (x.v[ix].weight < y.v[iy].weight) z.v[z.n++] = x.v[ix++];
        else z.v[z.n++] = y.v[iy++];
    }
    free(x.v); free(y.v);
    return z;
}
static int pm_cmp_weight(const void *a, const void *b) { /* used on (weight, original index) pairs -> stable */
    const size_t *x = (const size_t *)a, *y = (const size_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1]);
}
static void length_limited_calc(uint8_t max_bitwidth, const size_t *freq, int nf, uint8_t *bitwidth /* [nf] */) {
    memset(bitwidth, 0, nf);
    size_t keyed[320][2]; int ns = 0;
    for (int i = 0; i < nf; i++) if (freq[i] > 0) { keyed[ns][0] = freq[i]; keyed[ns][1] = (size_t)i; ns++; }
    qsort(keyed, ns, sizeof keyed[0], pm_cmp_weight);          /* sort_by_key(weight), stable (huffman.rs:315) */
    pm_list source; source.n = ns; source.v = (pm_node *)malloc((ns ? ns : 1) * sizeof(pm_node));
    for (int i = 0; i < ns; i++) {
        source.v[i].weight = keyed[i][0]; source.v[i].nsyms = 1;
        source.v[i].syms = (uint16_t *)malloc(2); source.v[i].syms[0] = (uint16_t)keyed[i][1];
    }
    pm_list weighted = pm_clone_list(&source);
    for (int r = 0; r < (int)max_bitwidth - 1; r++)            /* fold (huffman.rs:317-318) */
        weighted = pm_merge(pm_package(weighted), pm_clone_list(&source));
    weighted = pm_package(weighted);
    for (int i = 0; i < weighted.n; i++)
        for (uint32_t k = 0; k < weighted.v[i].nsyms; k++) bitwidth[weighted.v[i].syms[k]]++;
    pm_free_list(&weighted); pm_free_list(&source);
}

/* Builder::restore_canonical_huffman_codes for EncoderBuilder (huffman.rs:35-55, 192-216) */
static henc henc_from_bitwidthes(const uint8_t *bw, int n) {
    int symbol_count = 1;
    for (int i = n - 1; i >= 0; i--) if (bw[i] > 0) { symbol_count = i + 1; break; }
    henc e; e.n = (uint32_t)symbol_count; e.table = (hcode *)calloc(symbol_count, sizeof(hcode));
    uint16_t code = 0; uint8_t prev = 0;
    for (int w = 1; w <= 15; w++)                               /* stable sort by width == ascending width, then symbol */
        for (int s = 0; s < n; s++) if (bw[s] == w) {
            code = (uint16_t)(code << (w - prev));
            e.table[s].width = (uint8_t)w; e.table[s].bits = inverse_endian((uint8_t)w, code);
            code = (uint16_t)(code + 1); prev = (uint8_t)w;
        }
    return e;
}
/* EncoderBuilder::from_frequencies (huffman.rs:202-209) */
static henc henc_from_frequencies(const size_t *freq, int nf, uint8_t max_bitwidth) {
    uint8_t opt = calc_optimal_max_bitwidth(freq, nf);
    uint8_t mb = max_bitwidth < opt ? max_bitwidth : opt;
    uint8_t bw[320];
    length_limited_calc(mb, freq, nf, bw);
    return henc_from_bitwidthes(bw, nf);
}
static int henc_used_max_symbol(const henc *e) {                /* huffman.rs:247-253; -1 == None */
    for (int i = (int)e->n - 1; i >= 0; i--) if (e->table[i].width > 0) return i;
    return -1;
}
static void henc_encode(const henc *e, bitwriter *w, uint16_t sym) { bw_write_bits(w, e->table[sym].width, e->table[sym].bits); }

/* ================================================================== symbols (src/deflate/symbol.rs) */
static const uint8_t BITWIDTH_CODE_ORDER[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
static const uint16_t LENGTH_BASE[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
static const uint8_t LENGTH_EXTRA[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
static const uint16_t DIST_BASE[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577 };
static const uint8_t DIST_EXTRA[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };

static uint16_t sym_code(uint32_t s) {                          /* Symbol::code, symbol.rs:95-111 */
    if (s == SYM_EOB) return 256;
    if (!(s & SYM_PTR)) return (uint16_t)s;
    uint32_t length = (s >> 16) & 0x1FF;
    if (length <= 10) return (uint16_t)(257 + length - 3);
    if (length <= 18) return (uint16_t)(265 + (length - 11) / 2);
    if (length <= 34) return (uint16_t)(269 + (length - 19) / 4);
    if (length <= 66) return (uint16_t)(273 + (length - 35) / 8);
    if (length <= 130) return (uint16_t)(277 + (length - 67) / 16);
    if (length <= 257) return (uint16_t)(281 + (length - 131) / 32);
    return 285;
}
static int sym_extra_length(uint32_t s, uint8_t *bits, uint16_t *extra) { /* symbol.rs:112-126 */
    if (s == SYM_EOB || !(s & SYM_PTR)) return 0;
    uint32_t length = (s >> 16) & 0x1FF;
    if (length <= 10 || length == 258) return 0;
    if (length <= 18) { *bits = 1; *extra = (uint16_t)((length - 11) % 2); return 1; }
    if (length <= 34) { *bits = 2; *extra = (uint16_t)((length - 19) % 4); return 1; }
    if (length <= 66) { *bits = 3; *extra = (uint16_t)((length - 35) % 8); return 1; }
    if (length <= 130) { *bits = 4; *extra = (uint16_t)((length - 67) % 16); return 1; }
    *bits = 5; *extra = (uint16_t)((length - 131) % 32); return 1;
}
static int sym_distance(uint32_t s, uint8_t *code, uint8_t *bits, uint16_t *extra) { /* symbol.rs:127-154 */
    if (s == SYM_EOB || !(s & SYM_PTR)) return 0;
    uint32_t distance = s & 0xFFFF;
    if (distance <= 4) { *code = (uint8_t)(distance - 1); *bits = 0; *extra = 0; return 1; }
    uint8_t eb = 1, c = 4; uint32_t base = 4;
    while (base * 2 < distance) { eb++; c += 2; base *= 2; }
    uint32_t half = base / 2, delta = distance - base - 1;
    *bits = eb; *extra = (uint16_t)(delta % half);
    *code = distance <= base + half ? c : (uint8_t)(c + 1);
    return 1;
}

typedef struct { henc literal, distance; } sym_encoder;
static void sym_encoder_free(sym_encoder *e) { free(e->literal.table); free(e->distance.table); }

static void sym_encode(const sym_encoder *e, bitwriter *w, uint32_t s) {  /* symbol::Encoder::encode, symbol.rs:168-183 */
    henc_encode(&e->literal, w, sym_code(s));
    uint8_t bits, code; uint16_t extra;
    if (sym_extra_length(s, &bits, &extra)) bw_write_bits(w, bits, extra);
    if (sym_distance(s, &code, &bits, &extra)) {
        henc_encode(&e->distance, w, code);
        if (bits > 0) bw_write_bits(w, bits, extra);
    }
}

/* FixedHuffmanCodec::build (symbol.rs:260-281) */
static sym_encoder fixed_build(void) {
    sym_encoder e;
    e.literal.n = 288; e.literal.table = (hcode *)calloc(288, sizeof(hcode));
    static const struct { uint8_t bw; uint16_t lo, hi, base; } T[4] = {
        { 8, 0, 144, 0x30 }, { 9, 144, 256, 0x190 }, { 7, 256, 280, 0 }, { 8, 280, 288, 0xC0 } };
    for (int t = 0; t < 4; t++)
        for (uint16_t s = T[t].lo; s < T[t].hi; s++) {
            e.literal.table[s].width = T[t].bw;
            e.literal.table[s].bits = inverse_endian(T[t].bw, (uint16_t)(T[t].base + (s - T[t].lo)));
        }
    e.distance.n = 30; e.distance.table = (hcode *)calloc(30, sizeof(hcode));
    for (uint16_t i = 0; i < 30; i++) { e.distance.table[i].width = 5; e.distance.table[i].bits = inverse_endian(5, i); }
    return e;
}
/* DynamicHuffmanCodec::build (symbol.rs:321-342) */
static sym_encoder dynamic_build(const symvec *syms) {
    size_t lit[286], dist[30]; memset(lit, 0, sizeof lit); memset(dist, 0, sizeof dist);
    int empty_distance_table = 1;
    for (size_t i = 0; i < syms->len; i++) {
        uint32_t s = syms->p[i];
        lit[sym_code(s)]++;
        uint8_t c, b; uint16_t x;
        if (sym_distance(s, &c, &b, &x)) { empty_distance_table = 0; dist[c]++; }
    }
    if (empty_distance_table) dist[0] = 1;
    sym_encoder e;
    e.literal = henc_from_frequencies(lit, 286, 15);
    e.distance = henc_from_frequencies(dist, 30, 15);
    return e;
}
/* build_bitwidth_codes (symbol.rs:486-540) -> triples (code, bits, extra) */
typedef struct { uint8_t code, bits, extra; } bwcode;
static int build_bitwidth_codes(const sym_encoder *codec, int lit_count, int dist_count, bwcode *out /* >= 320 */) {
    struct { uint8_t value; size_t count; } runs[320]; int nr = 0;
    const henc *es[2] = { &codec->literal, &codec->distance }; int sizes[2] = { lit_count, dist_count };
    for (int t = 0; t < 2; t++)
        for (int i = 0; i < sizes[t]; i++) {
            uint8_t c = es[t]->table[i].width;
            if (i > 0 && nr > 0 && runs[nr - 1].value == c) runs[nr - 1].count++;
            else { runs[nr].value = c; runs[nr].count = 1; nr++; }
        }
    int n = 0;
    for (int r = 0; r < nr; r++) {
        if (runs[r].value == 0) {
            size_t c = runs[r].count;
            while (c >= 11) { uint8_t k = (uint8_t)(c < 138 ? c : 138); out[n++] = (bwcode){ 18, 7, (uint8_t)(k - 11) }; c -= k; }
            if (c >= 3) { out[n++] = (bwcode){ 17, 3, (uint8_t)(c - 3) }; c = 0; }
            for (size_t k = 0; k < c; k++) out[n++] = (bwcode){ 0, 0, 0 };
        } else {
            out[n++] = (bwcode){ runs[r].value, 0, 0 };
            size_t c = runs[r].count - 1;
            while (c >= 3) { uint8_t k = (uint8_t)(c < 6 ? c : 6); out[n++] = (bwcode){ 16, 2, (uint8_t)(k - 3) }; c -= k; }
            for (size_t k = 0; k < c; k++) out[n++] = (bwcode){ runs[r].value, 0, 0 };
        }
    }
    return n;
}
/* DynamicHuffmanCodec::save (symbol.rs:343-386) */
static void dynamic_save(bitwriter *w, const sym_encoder *codec) {
    int um = henc_used_max_symbol(&codec->literal); int lit_count = (um < 0 ? 0 : um) + 1; if (lit_count < 257) lit_count = 257;
    um = henc_used_max_symbol(&codec->distance); int dist_count = (um < 0 ? 0 : um) + 1; if (dist_count < 1) dist_count = 1;
    bwcode codes[640]; int nc = build_bitwidth_codes(codec, lit_count, dist_count, codes);
    size_t code_counts[19]; memset(code_counts, 0, sizeof code_counts);
    for (int i = 0; i < nc; i++) code_counts[codes[i].code]++;
    henc be = henc_from_frequencies(code_counts, 19, 7);
    int bitwidth_code_count = 0;
    for (int k = 18; k >= 0; k--) {
        int i = BITWIDTH_CODE_ORDER[k];
        if (code_counts[i] != 0 && be.table[i].width > 0) { bitwidth_code_count = k + 1; break; }
    }
    if (bitwidth_code_count < 4) bitwidth_code_count = 4;
    bw_write_bits(w, 5, (uint16_t)(lit_count - 257));
    bw_write_bits(w, 5, (uint16_t)(dist_count - 1));
    bw_write_bits(w, 4, (uint16_t)(bitwidth_code_count - 4));
    for (int k = 0; k < bitwidth_code_count; k++) {
        int i = BITWIDTH_CODE_ORDER[k];
        uint16_t width = code_counts[i] == 0 ? 0 : be.table[i].width;
        bw_write_bits(w, 3, width);
    }
    for (int i = 0; i < nc; i++) {
        henc_encode(&be, w, codes[i].code);
        if (codes[i].bits > 0) bw_write_bits(w, codes[i].bits, codes[i].extra);
    }
    free(be.table);
}

/* ================================================================== deflate::Encoder (src/deflate/encode.rs) */
typedef struct {
    int mode;                   /* BlockType: Raw(0b00)/Fixed(0b01)/Dynamic(0b10) via ORC_MODE_* */
    size_t block_size;
    bytevec rawbuf;             /* RawBuf */
    lz77_enc lz77; symvec syms; size_t original_size;  /* CompressBuf */
    bitwriter writer;
} deflate_enc;

static void denc_init(deflate_enc *d, const orc_opts *o, bytevec *out) {
    memset(d, 0, sizeof *d);
    d->mode = o->mode;
    d->block_size = (size_t)o->block_size;
    if (d->mode == ORC_MODE_STORED && d->block_size > 0xFFFF) d->block_size = 0xFFFF;  /* encode.rs:121-127 */
    d->lz77.window_size = o->window_size > 32768 ? 32768 : o->window_size;
    d->lz77.max_length = o->max_length > 258 ? 258 : o->max_length;
    d->writer.inner = out;
}
static void denc_free(deflate_enc *d) { bv_free(&d->rawbuf); bv_free(&d->lz77.buf); free(d->syms.p); }
static size_t denc_len(const deflate_enc *d) { return d->mode == ORC_MODE_STORED ? d->rawbuf.len : d->original_size; }

/* Block::flush (encode.rs:287-295) + {RawBuf,CompressBuf}::flush (:367-382, :412-425) */
static void denc_block_flush(deflate_enc *d, int is_final) {
    static const uint16_t BT[3] = { 2 /*dynamic*/, 1 /*fixed*/, 0 /*raw*/ };
    bw_write_bits(&d->writer, 1, (uint16_t)(is_final ? 1 : 0));
    bw_write_bits(&d->writer, 2, BT[d->mode]);
    if (d->mode == ORC_MODE_STORED) {
        size_t size = d->rawbuf.len < 0xFFFF ? d->rawbuf.len : 0xFFFF;
        bw_flush(&d->writer);
        bv_push(d->writer.inner, (uint8_t)size); bv_push(d->writer.inner, (uint8_t)(size >> 8));
        uint16_t ns = (uint16_t)~size;
        bv_push(d->writer.inner, (uint8_t)ns); bv_push(d->writer.inner, (uint8_t)(ns >> 8));
        bv_extend(d->writer.inner, d->rawbuf.p, size);
        memmove(d->rawbuf.p, d->rawbuf.p + size, d->rawbuf.len - size); d->rawbuf.len -= size;
    } else {
        lz77_flush(&d->lz77, &d->syms);
        sv_push(&d->syms, SYM_EOB);
        sym_encoder enc = d->mode == ORC_MODE_DYNAMIC ? dynamic_build(&d->syms) : fixed_build();
        if (d->mode == ORC_MODE_DYNAMIC) dynamic_save(&d->writer, &enc);
        for (size_t i = 0; i < d->syms.len; i++) sym_encode(&enc, &d->writer, d->syms.p[i]);
        d->syms.len = 0;
        d->original_size = 0;
        sym_encoder_free(&enc);
    }
}
/* Block::write (encode.rs:277-286) */
static void denc_write(deflate_enc *d, const uint8_t *b, size_t n) {
    if (d->mode == ORC_MODE_STORED) bv_extend(&d->rawbuf, b, n);
    else { d->original_size += n; lz77_encode(&d->lz77, b, n, &d->syms); }
    while (denc_len(d) >= d->block_size) denc_block_flush(d, 0);
}
static void denc_flush(deflate_enc *d) { denc_block_flush(d, 0); }          /* io::Write::flush, encode.rs:245-248 */
static void denc_zlib_sync_flush(deflate_enc *d) {                            /* encode.rs:225-234 */
    denc_block_flush(d, 0);
    bw_write_bits(&d->writer, 1, 0); bw_write_bits(&d->writer, 2, 0);
    bw_flush(&d->writer);
    static const uint8_t m[4] = { 0, 0, 255, 255 };
    bv_extend(d->writer.inner, m, 4);
}
static void denc_finish(deflate_enc *d) { denc_block_flush(d, 1); bw_flush(&d->writer); }  /* encode.rs:296-303 */

/* ================================================================== checksums (src/checksum.rs -> adler32 1.x, crc32fast 1.x) */
static uint32_t crc_tab[8][256]; static int crc_ready = 0;
static void crc_init(void) {
    for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; crc_tab[0][i] = c; }
    for (uint32_t i = 0; i < 256; i++) for (int t = 1; t < 8; t++) crc_tab[t][i] = (crc_tab[t - 1][i] >> 8) ^ crc_tab[0][crc_tab[t - 1][i] & 0xFF];
    crc_ready = 1;
}
uint32_t orc_crc32(uint32_t crc, const uint8_t *p, size_t n) {   /* crc = previous value() (0 to start) */
    if (!crc_ready) crc_init();
    uint32_t c = ~crc;
    while (n >= 8) {
        uint32_t a = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        a ^= c;
        c = crc_tab[7][a & 0xFF] ^ crc_tab[6][(a >> 8) & 0xFF] ^ crc_tab[5][(a >> 16) & 0xFF] ^ crc_tab[4][a >> 24]
          ^ crc_tab[3][p[4]] ^ crc_tab[2][p[5]] ^ crc_tab[1][p[6]] ^ crc_tab[0][p[7]];
        p += 8; n -= 8;
    }
    while (n--) c = crc_tab[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
    return ~c;
}
uint32_t orc_adler32(uint32_t adler, const uint8_t *p, size_t n) { /* adler = previous value() (1 to start) */
    uint32_t a = adler & 0xFFFF, b = adler >> 16;
    while (n) {
        size_t k = n < 5552 ? n : 5552; n -= k;
        while (k--) { a += *p++; b += a; }
        a %= 65521; b %= 65521;
    }
    return (b << 16) | a;
}

/* ================================================================== containers: gzip (src/gzip.rs), zlib (src/zlib.rs) */
static void gzip_header_write(bytevec *out, const orc_opts *o, int with_hcrc_flag_and_crc) { /* Header::write_to, gzip.rs:368-389 */
    size_t start = out->len;
    uint8_t flags = 0;                                       /* Header::flags, gzip.rs:343-355 */
    if (o->gzip_is_text) flags |= 1;
    if (with_hcrc_flag_and_crc && o->gzip_is_verified) flags |= 2;
    if (o->gzip_has_extra) flags |= 4;
    if (o->gzip_filename) flags |= 8;
    if (o->gzip_comment) flags |= 16;
    uint8_t xfl = 0;   /* CompressionLevel::Unknown for DefaultLz77Encoder (Balance) and for no_compression (gzip.rs:84-92, 693-697) */
    uint8_t h[10] = { 31, 139, 8, flags, (uint8_t)o->gzip_mtime, (uint8_t)(o->gzip_mtime >> 8), (uint8_t)(o->gzip_mtime >> 16), (uint8_t)(o->gzip_mtime >> 24), xfl, o->gzip_os };
    bv_extend(out, h, 10);
    if (o->gzip_has_extra) { bv_push(out, (uint8_t)o->gzip_extra_len); bv_push(out, (uint8_t)(o->gzip_extra_len >> 8)); bv_extend(out, o->gzip_extra, o->gzip_extra_len); }
    if (o->gzip_filename) bv_extend(out, (const uint8_t *)o->gzip_filename, strlen(o->gzip_filename) + 1);
    if (o->gzip_comment) bv_extend(out, (const uint8_t *)o->gzip_comment, strlen(o->gzip_comment) + 1);
    if (with_hcrc_flag_and_crc && o->gzip_is_verified) {
        /* Header::crc16 (gzip.rs:356-367): CRC-32 of the header re-serialised with is_verified=false, low 16 bits */
        bytevec tmp = { 0 }; gzip_header_write(&tmp, o, 0);
        uint32_t c = orc_crc32(0, tmp.p, tmp.len); bv_free(&tmp);
        bv_push(out, (uint8_t)c); bv_push(out, (uint8_t)(c >> 8));
    }
    (void)start;
}
static void zlib_header_write(bytevec *out, const orc_opts *o) { /* zlib::Header::from_lz77 + write_to (zlib.rs:212-220, 267-279) */
    uint32_t ws = o->window_size > 32768 ? 32768 : o->window_size;
    uint8_t cinfo = ws > 16384 ? 7 : ws > 8192 ? 6 : ws > 4096 ? 5 : ws > 2048 ? 4 : ws > 1024 ? 3 : ws > 512 ? 2 : ws > 256 ? 1 : 0;
    uint8_t level = o->mode == ORC_MODE_STORED ? 0 : 2;     /* no_compression -> Fastest (zlib.rs:466-470); Balance -> Default */
    uint8_t cmf = (uint8_t)((cinfo << 4) | 8), flg = (uint8_t)(level << 6);
    uint16_t check = (uint16_t)(((uint16_t)cmf << 8) + flg);
    if (check % 31 != 0) flg = (uint8_t)(flg + (31 - check % 31));
    bv_push(out, cmf); bv_push(out, flg);
}

/* Encode one stream under an explicit write schedule.
 * sched[k] >= 0 : io::Write::write of that many bytes (consumed in order from `in`); sched[k] == -1 : io::Write::flush().
 * sched == NULL : a single write_all(in) (no write call at all when n == 0). finish() is always called at the end. */
int orc_encode(int fmt, const orc_opts *o, const uint8_t *in, size_t n, const int64_t *sched, size_t n_sched,
               uint8_t *out, size_t cap, size_t *out_len) {
    bytevec ob = { 0 };
    if (fmt == ORC_FMT_GZIP) gzip_header_write(&ob, o, 1);
    else if (fmt == ORC_FMT_ZLIB) zlib_header_write(&ob, o);
    deflate_enc d; denc_init(&d, o, &ob);
    uint32_t crc = 0, adler = 1, isize = 0;
    size_t pos = 0;
    int64_t one = (int64_t)n; size_t ns = n_sched;
    if (!sched) { sched = &one; ns = n ? 1 : 0; }
    for (size_t k = 0; k < ns; k++) {
        if (sched[k] < 0) {
            if (fmt == ORC_FMT_ZLIB && o->zlib_flush_sync) denc_zlib_sync_flush(&d); else denc_flush(&d);
        } else {
            size_t w = (size_t)sched[k]; if (pos + w > n) w = n - pos;
            denc_write(&d, in + pos, w);
            if (fmt == ORC_FMT_GZIP) { crc = orc_crc32(crc, in + pos, w); isize += (uint32_t)w; }   /* gzip.rs:890-895 */
            else if (fmt == ORC_FMT_ZLIB) adler = orc_adler32(adler, in + pos, w);                 /* zlib.rs:661-665 */
            pos += w;
        }
    }
    denc_finish(&d);
    if (fmt == ORC_FMT_GZIP) {          /* Trailer::write_to, gzip.rs:114-121 */
        for (int k = 0; k < 4; k++) bv_push(&ob, (uint8_t)(crc >> (8 * k)));
        for (int k = 0; k < 4; k++) bv_push(&ob, (uint8_t)(isize >> (8 * k)));
    } else if (fmt == ORC_FMT_ZLIB) {   /* zlib.rs:630-638, big endian */
        for (int k = 3; k >= 0; k--) bv_push(&ob, (uint8_t)(adler >> (8 * k)));
    }
    denc_free(&d);
    *out_len = ob.len;
    int rc = ORC_OK;
    if (ob.len > cap) rc = ORC_OUTPUT_TOO_SMALL; else memcpy(out, ob.p, ob.len);
    bv_free(&ob);
    return rc;
}

/* Lz77Encode through the trait: encode(buf) then flush() of a fresh DefaultLz77Encoder. codes cap must be >= n. */
int orc_lz77_default(const uint8_t *buf, size_t n, uint32_t window, uint32_t max_len, uint32_t *codes, size_t *n_codes) {
    lz77_enc e; memset(&e, 0, sizeof e); e.window_size = window > 32768 ? 32768 : window; e.max_length = max_len > 258 ? 258 : max_len;
    symvec sv = { 0 };
    lz77_encode(&e, buf, n, &sv);
    lz77_flush(&e, &sv);
    memcpy(codes, sv.p, sv.len * 4); *n_codes = sv.len;
    free(sv.p); bv_free(&e.buf);
    return ORC_OK;
}

/* Code lengths as EncoderBuilder::from_frequencies would assign them; widths[nf]. */
int orc_huffman_lengths(const uint64_t *freq, int nf, int max_bitwidth, uint8_t *widths) {
    size_t f[320]; if (nf > 320 || nf <= 0) return ORC_INVALID_DATA;
    for (int i = 0; i < nf; i++) f[i] = (size_t)freq[i];
    uint8_t opt = calc_optimal_max_bitwidth(f, nf);
    uint8_t mb = (uint8_t)max_bitwidth < opt ? (uint8_t)max_bitwidth : opt;
    length_limited_calc(mb, f, nf, widths);
    return ORC_OK;
}

/* ================================================================== decode side */
typedef struct { const uint8_t *p; size_t n, pos; } memreader;       /* the inner `R: io::Read` */
static int mr_read_exact(memreader *r, uint8_t *dst, size_t k) {     /* short read consumes what is there, as read_exact does */
    if (r->n - r->pos < k) { r->pos = r->n; return ORC_UNEXPECTED_EOF; }
    memcpy(dst, r->p + r->pos, k); r->pos += k; return ORC_OK;
}

typedef struct { memreader *inner; uint32_t last_read; uint8_t offset; int last_error; char msg[128]; } bitreader; /* bit.rs:53-60 */
static void br_init(bitreader *b, memreader *r) { b->inner = r; b->last_read = 0; b->offset = 32; b->last_error = 0; b->msg[0] = 0; }
static void br_set_error(bitreader *b, int e, const char *m) { b->last_error = e; snprintf(b->msg, sizeof b->msg, "%s", m); }
static int br_fill_next_u8(bitreader *b) {                            /* bit.rs:132-141 */
    b->offset = (uint8_t)(b->offset - 8);
    b->last_read >>= 8;
    uint8_t x;
    if (mr_read_exact(b->inner, &x, 1)) return ORC_UNEXPECTED_EOF;
    b->last_read |= (uint32_t)x << 24;
    return ORC_OK;
}
static uint16_t br_peek(bitreader *b, uint8_t bw) {                  /* peek_bits_unchecked, bit.rs:111-125 */
    while (32 < (unsigned)b->offset + bw) {
        if (b->last_error) return 0;
        if (br_fill_next_u8(b)) { br_set_error(b, ORC_UNEXPECTED_EOF, "failed to fill whole buffer"); return 0; }
    }
    uint16_t bits = (uint16_t)(b->last_read >> (b->offset & 31));    /* wrapping_shr */
    return (uint16_t)(bits & ((1u << bw) - 1));
}
static void br_skip(bitreader *b, uint8_t bw) { b->offset = (uint8_t)(b->offset + bw); }
static uint16_t br_read_unchecked(bitreader *b, uint8_t bw) { uint16_t v = br_peek(b, bw); br_skip(b, bw); return v; }
static int br_check(bitreader *b) { int e = b->last_error; b->last_error = 0; return e; }   /* check_last_error: take() */
static int br_read_bits(bitreader *b, uint8_t bw, uint16_t *v) { *v = br_read_unchecked(b, bw); return br_check(b); }

typedef struct { uint16_t *table; uint8_t safely_peek_bitwidth, max_bitwidth; } hdec;   /* huffman::Decoder */
/* DecoderBuilder::from_bitwidthes + restore_canonical_huffman_codes + set_mapping + finish (huffman.rs:35-55, 66-132) */
static int hdec_from_bitwidthes(const uint8_t *bw, int n, int safely /* -1 == None */, int eob /* -1 == None */, hdec *out, char *msg) {
    uint8_t max_bw = 0; for (int i = 0; i < n; i++) if (bw[i] > max_bw) max_bw = bw[i];
    size_t tn = (size_t)1 << max_bw;
    uint16_t *table = (uint16_t *)malloc(tn * 2);
    for (size_t i = 0; i < tn; i++) table[i] = 16;
    uint16_t code = 0; uint8_t prev = 0;
    for (int w = 1; w <= 15; w++)
        for (int s = 0; s < n; s++) if (bw[s] == w) {
            code = (uint16_t)(code << (w - prev));
            if (s == eob) safely = w;
            uint16_t value = (uint16_t)(((uint16_t)s << 5) | (uint16_t)w);
            uint16_t be = inverse_endian((uint8_t)w, code);
            uint32_t npad = 1u << (max_bw - w);
            for (uint32_t padding = 0; padding < npad; padding++) {
                size_t i = (size_t)((padding << w) | be);
                if (table[i] != 16) { free(table); snprintf(msg, 128, "Bit region conflict"); return ORC_INVALID_DATA; }
                table[i] = value;
            }
            code = (uint16_t)(code + 1); prev = (uint8_t)w;
        }
    out->table = table; out->max_bitwidth = max_bw;
    int sp = safely < 0 ? 1 : safely;
    out->safely_peek_bitwidth = (uint8_t)(max_bw < sp ? max_bw : sp);
    return ORC_OK;
}
static uint16_t hdec_decode_unchecked(const hdec *d, bitreader *r) {  /* huffman.rs:157-179 */
    uint16_t value; uint8_t bitwidth, peek = d->safely_peek_bitwidth;
    for (;;) {
        uint16_t code = br_peek(r, peek);
        value = d->table[code];
        bitwidth = (uint8_t)(value & 31);
        if (bitwidth <= peek) break;
        if (bitwidth > d->max_bitwidth) { br_set_error(r, ORC_INVALID_DATA, "Invalid huffman coded stream"); break; }
        peek = bitwidth;
    }
    br_skip(r, bitwidth);
    return (uint16_t)(value >> 5);
}

typedef struct { hdec literal, distance; } sym_decoder;
static void sym_decoder_free(sym_decoder *d) { free(d->literal.table); free(d->distance.table); }

static int fixed_load(sym_decoder *out, char *msg) {                  /* FixedHuffmanCodec::load, symbol.rs:290-315 */
    uint8_t lw[288]; for (int s = 0; s < 288; s++) lw[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
    /* explicit (width, code) mappings equal the canonical assignment for these widths; DecoderBuilder::new(9, None, Some(256)) */
    int rc = hdec_from_bitwidthes(lw, 288, -1, 256, &out->literal, msg);
    if (rc) return rc;
    uint8_t dw[30]; memset(dw, 5, 30);
    rc = hdec_from_bitwidthes(dw, 30, out->literal.safely_peek_bitwidth, -1, &out->distance, msg);
    if (rc) free(out->literal.table);
    return rc;
}
/* load_bitwidthes (symbol.rs:457-484): appends to v */
static int load_bitwidthes(bitreader *r, uint16_t code, int last /* -1 none */, uint8_t *v, int *n, char *msg) {
    uint16_t x; int rc;
    if (code <= 15) { v[(*n)++] = (uint8_t)code; return ORC_OK; }
    if (code == 16) {
        if ((rc = br_read_bits(r, 2, &x))) { snprintf(msg, 128, "%s", r->msg); return rc; }
        if (last < 0) { snprintf(msg, 128, "No preceding value"); return ORC_INVALID_DATA; }
        for (int k = 0; k < x + 3; k++) v[(*n)++] = (uint8_t)last;
        return ORC_OK;
    }
    if (code == 17) {
        if ((rc = br_read_bits(r, 3, &x))) { snprintf(msg, 128, "%s", r->msg); return rc; }
        for (int k = 0; k < x + 3; k++) v[(*n)++] = 0;
        return ORC_OK;
    }
    if ((rc = br_read_bits(r, 7, &x))) { snprintf(msg, 128, "%s", r->msg); return rc; }
    for (int k = 0; k < x + 11; k++) v[(*n)++] = 0;
    return ORC_OK;
}
/* DynamicHuffmanCodec::load (symbol.rs:387-456) */
static int dynamic_load(bitreader *r, sym_decoder *out, char *msg) {
    uint16_t a, b, c; int rc;
#define RB(nb, var) if ((rc = br_read_bits(r, nb, &var))) { snprintf(msg, 128, "%s", r->msg); return rc; }
    RB(5, a) RB(5, b) RB(4, c)
    int literal_code_count = a + 257, distance_code_count = b + 1, bitwidth_code_count = c + 4;
    if (distance_code_count > 30) { snprintf(msg, 128, "The value of HDIST is too big: max=30, actual=%d", distance_code_count); return ORC_INVALID_DATA; }
    uint8_t bcb[19]; memset(bcb, 0, 19);
    for (int k = 0; k < bitwidth_code_count; k++) { uint16_t x; RB(3, x) bcb[BITWIDTH_CODE_ORDER[k]] = (uint8_t)x; }
#undef RB
    hdec bd;
    if ((rc = hdec_from_bitwidthes(bcb, 19, 1, -1, &bd, msg))) return rc;
    uint8_t lit[1024]; int nl = 0;      /* at most 287 + 138 entries */
    while (nl < literal_code_count) {
        uint16_t cc = hdec_decode_unchecked(&bd, r);
        if ((rc = br_check(r))) { snprintf(msg, 128, "%s", r->msg); free(bd.table); return rc; }
        if ((rc = load_bitwidthes(r, cc, nl ? lit[nl - 1] : -1, lit, &nl, msg))) { free(bd.table); return rc; }
    }
    uint8_t dist[1024]; int nd = 0;
    for (int k = literal_code_count; k < nl; k++) dist[nd++] = lit[k];   /* drain(literal_code_count..) */
    nl = literal_code_count;
    while (nd < distance_code_count) {
        uint16_t cc = hdec_decode_unchecked(&bd, r);
        if ((rc = br_check(r))) { snprintf(msg, 128, "%s", r->msg); free(bd.table); return rc; }
        int last = nd ? dist[nd - 1] : (nl ? lit[nl - 1] : -1);
        if ((rc = load_bitwidthes(r, cc, last, dist, &nd, msg))) { free(bd.table); return rc; }
    }
    free(bd.table);
    if (nd > distance_code_count) { snprintf(msg, 128, "The length of `distance_code_bitwidthes` is too large: actual=%d, expected=%d", nd, distance_code_count); return ORC_INVALID_DATA; }
    if ((rc = hdec_from_bitwidthes(lit, nl, -1, 256, &out->literal, msg))) return rc;
    if ((rc = hdec_from_bitwidthes(dist, nd, out->literal.safely_peek_bitwidth, -1, &out->distance, msg))) { free(out->literal.table); return rc; }
    return ORC_OK;
}

/* deflate::Decoder driven as read_to_end would drive it (src/deflate/decode.rs:81-165) + Lz77Decoder::decode
 * (libflate_lz77/src/lib.rs:164-194).  `hist0` = out->len at the start of the member (Lz77Decoder::clear on reset).
 * On error the bytes decoded so far stay in `out` (== bytes read + unread_decoded_data()). */
static int deflate_decode_stream(memreader *mr, bytevec *out, size_t hist0, char *msg) {
    bitreader br; br_init(&br, mr);
    int eos = 0, rc;
    while (!eos) {
        uint16_t bfinal, btype;
        if ((rc = br_read_bits(&br, 1, &bfinal))) { snprintf(msg, 128, "%s", br.msg); return rc; }
        if ((rc = br_read_bits(&br, 2, &btype))) { snprintf(msg, 128, "%s", br.msg); return rc; }
        eos = bfinal != 0;
        if (btype == 0) {                                   /* read_non_compressed_block, decode.rs:81-111 */
            br.offset = 32;                                 /* bit_reader.reset() */
            uint8_t b2[2];
            if (mr_read_exact(mr, b2, 2)) { snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
            uint16_t len = (uint16_t)(b2[0] | (b2[1] << 8));
            if (mr_read_exact(mr, b2, 2)) { snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
            uint16_t nlen = (uint16_t)(b2[0] | (b2[1] << 8));
            if ((uint16_t)~len != nlen) { snprintf(msg, 128, "LEN=%u is not the one's complement of NLEN=%u", len, nlen); return ORC_INVALID_DATA; }
            size_t avail = mr->n - mr->pos, used = avail < len ? avail : len;
            bv_extend(out, mr->p + mr->pos, used); mr->pos += used;
            if (used != len) { snprintf(msg, 128, "The reader has incorrect length: expected %u, read %zu", len, used); return ORC_UNEXPECTED_EOF; }
        } else if (btype == 1 || btype == 2) {              /* read_compressed_block, decode.rs:112-130 */
            sym_decoder sd;
            if ((rc = btype == 1 ? fixed_load(&sd, msg) : dynamic_load(&br, &sd, msg))) return rc;
            for (;;) {
                /* symbol::Decoder::decode_unchecked (symbol.rs:193-243) */
                uint16_t decoded = hdec_decode_unchecked(&sd.literal, &br);
                int kind; uint16_t length = 0, distance = 0;   /* kind 0 literal, 1 EOB, 2 pointer */
                if (decoded <= 255) kind = 0;
                else if (decoded == 256) kind = 1;
                else if (decoded == 286 || decoded == 287) {
                    char m[128]; snprintf(m, sizeof m, "The value %u must not occur in compressed data", decoded);
                    br_set_error(&br, ORC_INVALID_DATA, m); kind = 1;
                } else {
                    uint16_t extra = br_read_unchecked(&br, LENGTH_EXTRA[decoded - 257]);
                    length = (uint16_t)(LENGTH_BASE[decoded - 257] + extra); kind = 2;
                }
                if (kind == 2) {
                    uint16_t dsym = hdec_decode_unchecked(&sd.distance, &br);
                    uint16_t extra = br_read_unchecked(&br, DIST_EXTRA[dsym]);
                    distance = (uint16_t)(DIST_BASE[dsym] + extra);
                }
                if ((rc = br_check(&br))) { snprintf(msg, 128, "%s", br.msg); sym_decoder_free(&sd); return rc; }
                if (kind == 1) break;
                if (kind == 0) bv_push(out, (uint8_t)decoded);
                else {
                    size_t have = out->len - hist0;
                    if (have < distance) {
                        snprintf(msg, 128, "Too long backword reference: buffer.len=%zu, distance=%u", have, distance);
                        sym_decoder_free(&sd); return ORC_INVALID_DATA;
                    }
                    bv_reserve(out, length);
                    for (uint16_t k = 0; k < length; k++) { out->p[out->len] = out->p[out->len - distance]; out->len++; }
                }
            }
            sym_decoder_free(&sd);
        } else { snprintf(msg, 128, "btype 0x11 of DEFLATE is reserved(error) value"); return ORC_INVALID_DATA; }
    }
    return ORC_OK;
}

/* gzip::Header::read_from (gzip.rs:390-446). *crc16_expected mimics Header::crc16 of the re-serialised parsed header. */
static int gzip_header_read(memreader *mr, char *msg) {
    uint8_t b[10];
    if (mr_read_exact(mr, b, 10)) { snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
    if (b[0] != 31 || b[1] != 139) { snprintf(msg, 128, "Unexpected GZIP ID: value=[%u, %u], expected=[31, 139]", b[0], b[1]); return ORC_INVALID_DATA; }
    if (b[2] != 8) { snprintf(msg, 128, "Compression methods other than DEFLATE(8) are unsupported: method=%u", b[2]); return ORC_INVALID_DATA; }
    uint8_t flags = b[3];
    bytevec re = { 0 };          /* the header as Header::write_to would re-serialise it with is_verified = false */
    uint8_t nf = (uint8_t)(flags & (4 | 8 | 16));   /* read_from never sets is_text (gzip.rs:413-425), so F_TEXT is not re-serialised */
    uint8_t xfl = b[8] == 4 ? 4 : b[8] == 2 ? 2 : 0;         /* CompressionLevel::from_u8 -> to_u8 */
    uint8_t h[10] = { 31, 139, 8, nf, b[4], b[5], b[6], b[7], xfl, b[9] };
    bv_extend(&re, h, 10);
    if (flags & 4) {             /* ExtraField::read_from, gzip.rs:472-487 */
        uint8_t l2[2];
        if (mr_read_exact(mr, l2, 2)) { bv_free(&re); snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
        size_t data_size = (size_t)(l2[0] | (l2[1] << 8)), limit = data_size;
        size_t total_pos = re.len; bv_push(&re, 0); bv_push(&re, 0); size_t total = 0;
        while (limit > 0) {      /* subfields read through a Take: short reads are UnexpectedEof */
            uint8_t sf[4]; size_t k = limit < 4 ? limit : 4;
            if (mr->n - mr->pos < k || k < 4) { if (mr->n - mr->pos >= k) mr->pos += k; else mr->pos = mr->n; bv_free(&re); snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
            mr_read_exact(mr, sf, 4); limit -= 4;
            size_t dl = (size_t)(sf[2] | (sf[3] << 8));
            if (dl > limit || mr->n - mr->pos < dl) { size_t c = dl < limit ? dl : limit; if (mr->n - mr->pos < c) c = mr->n - mr->pos; mr->pos += c; bv_free(&re); snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
            bv_extend(&re, sf, 4); bv_extend(&re, mr->p + mr->pos, dl); mr->pos += dl; limit -= dl; total += 4 + dl;
        }
        re.p[total_pos] = (uint8_t)total; re.p[total_pos + 1] = (uint8_t)(total >> 8);
    }
    for (int f = 8; f <= 16; f <<= 1) if (flags & f) {       /* read_cstring, gzip.rs:448-461 */
        for (;;) {
            uint8_t c;
            if (mr_read_exact(mr, &c, 1)) { bv_free(&re); snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
            bv_push(&re, c);
            if (c == 0) break;
        }
    }
    if (flags & 2) {
        uint8_t c2[2];
        if (mr_read_exact(mr, c2, 2)) { bv_free(&re); snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
        uint16_t crc = (uint16_t)(c2[0] | (c2[1] << 8)), expected = (uint16_t)orc_crc32(0, re.p, re.len);
        if (crc != expected) { bv_free(&re); snprintf(msg, 128, "CRC16 of GZIP header mismatched: value=%u, expected=%u", crc, expected); return ORC_INVALID_DATA; }
    }
    bv_free(&re);
    return ORC_OK;
}
static int zlib_header_read(memreader *mr, char *msg) {      /* zlib::Header::read_from, zlib.rs:221-266 */
    uint8_t b[2];
    if (mr_read_exact(mr, b, 2)) { snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
    uint16_t check = (uint16_t)(((uint16_t)b[0] << 8) + b[1]);
    if (check % 31 != 0) { snprintf(msg, 128, "Inconsistent ZLIB check bits: `CMF(%u) * 256 + FLG(%u)` must be a multiple of 31", b[0], b[1]); return ORC_INVALID_DATA; }
    if ((b[0] & 15) != 8) { snprintf(msg, 128, "Compression methods other than DEFLATE(8) are unsupported: method=%u", b[0] & 15); return ORC_INVALID_DATA; }
    if ((b[0] >> 4) > 7) { snprintf(msg, 128, "CINFO above 7 are not allowed: value=%u", b[0] >> 4); return ORC_INVALID_DATA; }
    if (b[1] & 0x20) {
        uint8_t d[4];
        if (mr_read_exact(mr, d, 4)) { snprintf(msg, 128, "failed to fill whole buffer"); return ORC_UNEXPECTED_EOF; }
        snprintf(msg, 128, "Preset dictionaries are not supported: dictionary_id=0x%X", (unsigned)((d[0] << 24) | (d[1] << 16) | (d[2] << 8) | d[3]));
        return ORC_INVALID_DATA;
    }
    return ORC_OK;
}

/* Decoder::new(in) + read_to_end, for fmt in {DEFLATE, ZLIB, GZIP (first member only), GZIP_MULTI (MultiDecoder)}.
 * out_len = bytes decoded (including a partially decoded block when an error is returned);
 * in_consumed = bytes the decoder pulled from the underlying reader. */
int orc_decode(int fmt, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *out_len, size_t *in_consumed, char *errmsg /* >=128 or NULL */) {
    char msgbuf[128]; char *msg = errmsg ? errmsg : msgbuf; msg[0] = 0;
    memreader mr = { in, n, 0 };
    bytevec ob = { 0 };
    int rc = ORC_OK;
    if (fmt == ORC_FMT_DEFLATE) rc = deflate_decode_stream(&mr, &ob, 0, msg);
    else if (fmt == ORC_FMT_ZLIB) {
        rc = zlib_header_read(&mr, msg);
        if (!rc) rc = deflate_decode_stream(&mr, &ob, 0, msg);
        if (!rc) {                                         /* zlib.rs:377-409 */
            uint8_t t[4];
            if (mr_read_exact(&mr, t, 4)) { rc = ORC_UNEXPECTED_EOF; snprintf(msg, 128, "failed to fill whole buffer"); }
            else {
                uint32_t want = ((uint32_t)t[0] << 24) | ((uint32_t)t[1] << 16) | ((uint32_t)t[2] << 8) | t[3];
                uint32_t got = orc_adler32(1, ob.p, ob.len);
                if (want != got) { rc = ORC_INVALID_DATA; snprintf(msg, 128, "Adler32 checksum mismatched: value=%u, expected=%u", got, want); }
            }
        }
    } else {
        rc = gzip_header_read(&mr, msg);
        for (; !rc;) {
            size_t start = ob.len;
            rc = deflate_decode_stream(&mr, &ob, start, msg);
            if (rc) break;
            uint8_t t[8];                                   /* gzip.rs:1018-1047; ISIZE is read but not checked */
            if (mr_read_exact(&mr, t, 4) || mr_read_exact(&mr, t + 4, 4)) { rc = ORC_UNEXPECTED_EOF; snprintf(msg, 128, "failed to fill whole buffer"); break; }
            uint32_t want = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
            uint32_t got = orc_crc32(0, ob.p + start, ob.len - start);
            if (want != got) { rc = ORC_INVALID_DATA; snprintf(msg, 128, "CRC32 mismatched: value=%u, expected=%u", got, want); break; }
            if (fmt != ORC_FMT_GZIP_MULTI) break;
            int hrc = gzip_header_read(&mr, msg);           /* MultiDecoder::read, gzip.rs:1142-1166 */
            if (hrc == ORC_UNEXPECTED_EOF) { msg[0] = 0; break; }
            rc = hrc;
        }
    }
    *out_len = ob.len; *in_consumed = mr.pos;
    if (ob.len > cap) { if (rc == ORC_OK) rc = ORC_OUTPUT_TOO_SMALL; memcpy(out, ob.p, cap); }
    else if (ob.len) memcpy(out, ob.p, ob.len);
    bv_free(&ob);
    return rc;
}

/* Loading a dynamic block header only (reference test test_issues_3, src/deflate/decode.rs:176-190). */
int orc_dynamic_header_loads(const uint8_t *in, size_t n) {
    memreader mr = { in, n, 0 }; bitreader br; br_init(&br, &mr); uint16_t a, b; char msg[128];
    if (br_read_bits(&br, 1, &a) || br_read_bits(&br, 2, &b) || b != 2) return ORC_INVALID_DATA;
    sym_decoder sd; int rc = dynamic_load(&br, &sd, msg);
    if (!rc) sym_decoder_free(&sd);
    return rc;
}

size_t orc_opts_size(void) { return sizeof(orc_opts); }

"""CPU checks of the product's Huffman/header construction (libflate_b200/csrc/huff_build.cuh, compiled for
the host by tests/native) against the oracle.  No GPU needed: this is the same source the kernels run."""
import random

import numpy as np
import pytest

import hostcheck as hc
from oracle import oracle as orc

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
             6145, 8193, 12289, 16385, 24577]
DIST_EXTRA = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def test_length_and_distance_codes_match_rfc_tables():    # src/deflate/symbol.rs:22-87, 95-154
    for n in range(3, 259):
        code, eb, ex = hc.length_code(n)
        k = code - 257
        assert LEN_EXTRA[k] == eb and LEN_BASE[k] + ex == n and ex < (1 << eb)
        if k + 1 < 29:
            assert n < LEN_BASE[k + 1] or (n == 258)
    for d in range(1, 32769):
        code, eb, ex = hc.dist_code(d)
        assert DIST_EXTRA[code] == eb and DIST_BASE[code] + ex == d and ex < (1 << eb)


def _hists(rng, n):
    kind = rng.randrange(6)
    if kind == 0:
        return [rng.randrange(0, 5) for _ in range(n)]
    if kind == 1:
        return [rng.randrange(0, 1000) if rng.random() < 0.5 else 0 for _ in range(n)]
    if kind == 2:     # fibonacci-ish: forces deep trees / the length limit
        f, a, b = [], 1, 1
        for _ in range(n):
            f.append(a if rng.random() < 0.9 else 0); a, b = b, a + b
            if a > 1 << 28:
                a, b = 1, 1
        rng.shuffle(f)
        return f
    if kind == 3:     # many ties
        return [rng.choice([0, 1, 1, 2, 2, 4, 8]) for _ in range(n)]
    if kind == 4:     # zipf
        return [int(100000 / (i + 1) ** 1.1) for i in range(n)] if rng.random() < 0.5 else \
               sorted([int(100000 / (i + 1) ** 1.1) for i in range(n)], key=lambda _: rng.random())
    return [rng.randrange(1, 1 << 20) for _ in range(n)]


@pytest.mark.parametrize("n,cap", [(286, 15), (30, 15), (19, 7)])
def test_code_lengths_match_oracle(n, cap):               # src/huffman.rs:202-209, 257-363
    rng = random.Random(n * 100 + cap)
    for it in range(400):
        f = _hists(rng, n)
        if it == 0:
            f = [0] * n
        if it == 1:
            f = [0] * n; f[n // 2] = 7
        if it == 2:
            f = [0] * n; f[0] = 1; f[n - 1] = 1
        got = hc.code_lengths(f, cap)
        want = orc.huffman_lengths(f, cap)
        assert list(got) == list(want), (it, f)


def _sym_fields(code_word):
    """(litlen index, extra bits n, extra value, dist index or None, dist extra n, dist extra value)"""
    if code_word & 0x80000000:
        ln, dist = (code_word >> 16) & 0x1FF, code_word & 0xFFFF
        lc, leb, lex = hc.length_code(ln)
        dc, deb, dex = hc.dist_code(dist)
        return lc, leb, lex, dc, deb, dex
    return code_word, 0, 0, None, 0, 0


class BitOut:
    def __init__(self):
        self.acc, self.n = 0, 0

    def put(self, v, nb):
        self.acc |= int(v) << self.n
        self.n += nb

    def bytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


def assemble_stream(data, sched_blocks):
    """Python re-assembly of a raw DEFLATE stream from: oracle LZ77 codes per chunk + the PRODUCT's code tables/header.
    sched_blocks: list of blocks, each a list of chunk byte strings; the last block is final."""
    out = BitOut()
    for bi, chunks in enumerate(sched_blocks):
        syms = []
        for ch in chunks:
            syms += [int(x) for x in orc.lz77_default(ch)]
        hist = np.zeros(320, dtype=np.uint32)
        for s in syms:
            lc, _, _, dc, _, _ = _sym_fields(s)
            hist[lc] += 1
            if dc is not None:
                hist[286 + dc] += 1
        lit, dist, hdr, nbits = hc.block_codes(hist)
        out.put(1 if bi == len(sched_blocks) - 1 else 0, 1)
        out.put(2, 2)
        hv = int.from_bytes(hdr.tobytes(), "little") & ((1 << nbits) - 1)
        out.put(hv, nbits)
        for s in syms + [256]:
            lc, leb, lex, dc, deb, dex = _sym_fields(s)
            out.put(int(lit[lc]) & 0xFFFF, int(lit[lc]) >> 16)
            out.put(lex, leb)
            if dc is not None:
                out.put(int(dist[dc]) & 0xFFFF, int(dist[dc]) >> 16)
                out.put(dex, deb)
    return out.bytes()


def _text(rng, n):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz_") for _ in range(rng.randint(2, 9))) for _ in range(200)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b"\n"
    return bytes(b[:n])


def test_block_codes_and_header_reproduce_oracle_stream():  # symbol.rs:321-386 + 486-540, huffman.rs:35-55
    rng = random.Random(7)
    cases = [b"Hello World!", b"", b"a", b"aaaaa", bytes(range(256)) * 3, _text(rng, 3000), _text(rng, 70000),
             bytes(rng.getrandbits(8) for _ in range(5000)), b"\x00" * 100000, bytes([7]) * 2 + bytes([9]) * 300]
    for data in cases:
        got = assemble_stream(data, [[data] if data else []])   # < block_size: finish() emits one final block
        want = orc.encode(orc.FMT_DEFLATE, data)
        assert got == want, (len(data), data[:20])


def test_fixed_codes():                                     # symbol.rs:9-14, 260-281
    lit, dist = hc.fixed_codes()
    def rev(v, w):
        return int(format(v, f"0{w}b")[::-1], 2)
    for s in range(288):
        w = 8 if s < 144 else 9 if s < 256 else 7 if s < 280 else 8
        base = 0x30 + s if s < 144 else 0x190 + s - 144 if s < 256 else s - 256 if s < 280 else 0xC0 + s - 280
        assert lit[s] == (w << 16) | rev(base, w)
    for s in range(30):
        assert dist[s] == (5 << 16) | rev(s, 5)

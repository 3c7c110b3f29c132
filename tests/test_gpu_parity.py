"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Every test goes through the C ABI of libb2f.so and
compares with the CPU oracle (bit-exact) or with the reference's committed golden vectors."""
import json
import os
import random
import zlib as pyzlib
import gzip as pygzip

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "goldens.json")))


@pytest.fixture(scope="module")
def ctx():
    from libflate_b200 import native
    c = native.Context(0)
    yield c
    c.close()


def _text(rng, n, nwords=300):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz_") for _ in range(rng.randint(2, 9))) for _ in range(nwords)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b"\n"
    return bytes(b[:n])


def _cases(rng):
    return {
        "empty": b"", "one": b"a", "two": b"ab", "three": b"abc", "four": b"abcd", "aaaaa": b"aaaaa",
        "hello": b"Hello World!", "text3k": _text(rng, 3000), "text70k": _text(rng, 70000),
        "rand5k": bytes(rng.getrandbits(8) for _ in range(5000)), "zeros100k": b"\x00" * 100000,
        "w32767": _text(rng, 32767), "w32768": _text(rng, 32768), "w32769": _text(rng, 32769),
        "text300k": _text(rng, 300000), "counter1m": bytes(i & 255 for i in range(1 << 20)),
        "lowent": bytes(rng.getrandbits(8) >> 6 for _ in range(200000)),
        "periodic": (b"abcdefghij" * 30000)[:262144 + 77],
    }


# ------------------------------------------------------------------------------------------ C1 / C2
def test_checksum_kats(ctx):                                   # src/checksum.rs:45-56
    g = G["checksum_kat"]
    assert ctx.crc32([g["input"].encode()]) == [g["crc32"]]
    assert ctx.adler32([g["input"].encode()]) == [g["adler32"]]


def test_checksums_match_oracle_and_zlib(ctx):
    rng = random.Random(1)
    datas = [b"", b"a", bytes(rng.getrandbits(8) for _ in range(511)), bytes(rng.getrandbits(8) for _ in range(512)),
             bytes(rng.getrandbits(8) for _ in range(513)), _text(rng, 100003), b"\xff" * 1000000, _text(rng, 3 << 20)]
    assert ctx.crc32(datas) == [orc.crc32(d) for d in datas] == [pyzlib.crc32(d) for d in datas]
    assert ctx.adler32(datas) == [orc.adler32(d) for d in datas] == [pyzlib.adler32(d) for d in datas]
    # chaining through init values (Crc32::update called repeatedly)
    a, b = datas[5][:4321], datas[5][4321:]
    assert ctx.crc32([b], init=[pyzlib.crc32(a)]) == [pyzlib.crc32(datas[5])]
    assert ctx.adler32([b], init=[pyzlib.adler32(a)]) == [pyzlib.adler32(datas[5])]


# ------------------------------------------------------------------------------------------ E2: LZ77 codes
def test_lz77_issue21(ctx):                                    # src/lz77.rs:16-31
    assert list(ctx.lz77_default(b"aaaaa")) == [97, 0x80000000 | (4 << 16) | 1]


def test_lz77_codes_match_oracle(ctx):
    rng = random.Random(2)
    for name, d in _cases(rng).items():
        got, want = ctx.lz77_default(d), orc.lz77_default(d)
        assert len(got) == len(want), name
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (name, int(bad[0]), hex(int(got[bad[0]])), hex(int(want[bad[0]])))


def test_lz77_window_and_max_length_options(ctx):              # libflate_lz77/src/default.rs:222-239
    rng = random.Random(3)
    d = _text(rng, 120000, nwords=40)
    for window, max_len in ((1024, 258), (32768, 16), (4096, 3), (300, 100)):
        got, want = ctx.lz77_default(d, window, max_len), orc.lz77_default(d, window, max_len)
        assert np.array_equal(got, want), (window, max_len)


def test_lz77_multi_segment_chunk(ctx):                        # one chunk > 256 KiB: several chain segments + many tiles
    rng = random.Random(4)
    d = _text(rng, 3 * 262144 + 12345, nwords=2000)
    assert np.array_equal(ctx.lz77_default(d), orc.lz77_default(d))


def test_lz77_long_chains_and_long_runs(ctx):
    """k_lz_find's bounded walks: runs of identical lines fill a few hash buckets with thousands of entries, so text that follows
    (and collides with those buckets) needs k_lz_fixup (HBM walk) and k_lz_fixup2 (window scan); the runs themselves are
    period-p matches of length 258 at every position (cooperative extension, max_len cap, chunk end inside a run)."""
    rng = random.Random(11)
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyzABCDEFG_()0123456789") for _ in range(rng.randint(2, 9))) for _ in range(6000)]
    def text(n):
        b = bytearray()
        while len(b) < n:
            b += rng.choice(words) + b"\n"
        return bytes(b[:n])
    d = (b"Abc\n" * 9000 + text(50000) + b"The_quick\n" * 4000 + text(70000) + bytes(rng.getrandbits(8) for _ in range(20000)) +
         b"x" * 40000 + text(30000) + b"Zed_(film)\n" * 3000)
    for data in (d, d[:262144 + 5000], d[3:140001]):
        got, want = ctx.lz77_default(data), orc.lz77_default(data)
        assert len(got) == len(want)
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (int(bad[0]), hex(int(got[bad[0]])), hex(int(want[bad[0]])))
    for window, max_len in ((32768, 258), (5000, 258), (32768, 40)):
        assert np.array_equal(ctx.lz77_default(d[:200000], window, max_len), orc.lz77_default(d[:200000], window, max_len)), (window, max_len)
    sched = [8192] * (len(d) // 8192 + 1)
    assert ctx.encode(1, d, sched) == orc.encode(1, d, sched)                   # zlib, 256 KiB chunks: runs cut by chunk ends


# ------------------------------------------------------------------------------------------ encode goldens
def test_encode_goldens(ctx):
    from libflate_b200 import native as nv
    g = G["deflate_hello_dynamic"]
    assert list(ctx.encode(nv.FMT_DEFLATE, g["plain"].encode())) == g["bytes"]         # src/deflate/encode.rs:152-154
    g = G["zlib_hello_default"]
    assert list(ctx.encode(nv.FMT_ZLIB, g["plain"].encode())) == g["bytes"]            # src/zlib.rs:547-549
    g = G["deflate_hello_stored"]
    assert list(ctx.encode(nv.FMT_DEFLATE, g["plain"].encode(), mode=nv.MODE_STORED)) == g["bytes"]
    g = G["zlib_raw_encode"]
    assert list(ctx.encode(nv.FMT_ZLIB, g["plain"].encode(), mode=nv.MODE_STORED)) == g["bytes"]
    g = G["gzip_stored_mtime123"]
    assert list(ctx.encode(nv.FMT_GZIP, g["plain"].encode(), mode=nv.MODE_STORED, mtime=123)) == g["bytes"]
    for key, sync in (("zlib_issue27_none", False), ("zlib_issue27_sync", True)):     # src/zlib.rs:840-902
        g = G[key]
        writes = [w.encode() for w in g["writes"]]
        sched = ([len(w) for w in writes] + [nv.FLUSH]) * 2
        assert list(ctx.encode(nv.FMT_ZLIB, b"".join(writes) * 2, sched, zlib_flush_sync=sync)) == g["bytes"]


def test_issue52_size_bound(ctx):                              # src/deflate/encode.rs:435-457
    from libflate_b200 import native as nv
    data = open(os.path.join(HERE, "golden", "issue52_input.bin"), "rb").read()
    for lim in (16031, 16032):
        enc = ctx.encode(nv.FMT_DEFLATE, data[:lim])
        assert len(enc) < lim and enc == orc.encode(orc.FMT_DEFLATE, data[:lim])


# ------------------------------------------------------------------------------------------ encode == oracle
@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_encode_matches_oracle_single_write(ctx, fmt):
    rng = random.Random(10 + fmt)
    cases = _cases(rng)
    names = list(cases)
    got = ctx.encode_batch(fmt, [cases[k] for k in names], mtime=77)
    for k, enc in zip(names, got):
        want = orc.encode(fmt, cases[k], mtime=77)
        assert enc == want, (k, len(enc), len(want), next((i for i in range(min(len(enc), len(want))) if enc[i] != want[i]), -1))


def test_encode_matches_oracle_schedules(ctx):
    rng = random.Random(20)
    d = _text(rng, 2500000, nwords=3000)
    scheds = {
        "8k": [8192] * (len(d) // 8192 + 1), "7": None, "1m": [1 << 20] * 3, "odd": [100000, 1, 0, 262143, 1, 999999, 5000000],
        "flushes": [1000, -1, -1, 300000, -1, 1200000, -1], "big_first": [2000000, 10, -1, 600000],
    }
    scheds["7"] = [70001] * (len(d) // 70001 + 1)
    for fmt in (0, 1, 2):
        names = list(scheds)
        got = ctx.encode_batch(fmt, [d] * len(names), [scheds[k] for k in names], mtime=5)
        for k, enc in zip(names, got):
            want = orc.encode(fmt, d, scheds[k], mtime=5)
            assert enc == want, (fmt, k, len(enc), len(want))
    enc = ctx.encode(1, d, scheds["flushes"], zlib_flush_sync=True)
    assert enc == orc.encode(1, d, scheds["flushes"], zlib_flush_sync=True)
    assert pyzlib.decompress(enc) == d[:1501000]        # only the bytes the schedule writes belong to the stream


def test_encode_options_match_oracle(ctx):
    rng = random.Random(30)
    d = _text(rng, 400000, nwords=100)
    for kw in (dict(mode=1), dict(block_size=50000), dict(window_size=1024), dict(max_length=10), dict(window_size=2000, max_length=40, block_size=7777),
               dict(mode=2, block_size=1000), dict(mode=2)):
        for fmt in (0, 1, 2):
            for sched in (None, [30000] * 14, [5, -1, 100000, -1]):
                enc = ctx.encode(fmt, d, sched, **kw)
                assert enc == orc.encode(fmt, d, sched, **kw), (kw, fmt, sched and sched[:3])


def test_gzip_header_fields(ctx):                              # src/gzip.rs:126-288, 343-389
    kw = dict(mtime=123456, os_=11, is_text=True, is_verified=True, extra=bytes([0, 0x42, 3, 0]) + b"abc", filename=b"foo.txt", comment=b"hi")
    enc = ctx.encode(2, b"hello world hello world", **kw)
    assert enc == orc.encode(2, b"hello world hello world", **kw)


# ------------------------------------------------------------------------------------------ decode
def test_decode_goldens(ctx):
    from libflate_b200 import native as nv
    st, out, used, _ = ctx.decode(nv.FMT_DEFLATE, bytes(G["deflate_fixed_hello"]["bytes"]))        # src/deflate/decode.rs:28-33
    assert st == 0 and out == b"Hello World!" and used == 14
    st, out, _, _ = ctx.decode(nv.FMT_ZLIB, bytes(G["zlib_decode_works"]["bytes"]))                 # src/zlib.rs:708-730
    assert st == 0 and out == b"Hello World!"
    one = orc.encode(orc.FMT_GZIP, b"Hello World!")                                                  # src/gzip.rs:1217-1226
    st, out, used, _ = ctx.decode(nv.FMT_GZIP, one * 2)
    assert st == 0 and out == b"Hello World!" and used == len(one)
    st, out, used, _ = ctx.decode(nv.FMT_GZIP_MULTI, one * 2)
    assert st == 0 and out == b"Hello World!Hello World!" and used == 2 * len(one)
    st, out, _, _ = ctx.decode(nv.FMT_GZIP, open(os.path.join(HERE, "golden", "offset.gz"), "rb").read())   # src/non_blocking/gzip.rs:178-183
    assert st == 0 and out == open(os.path.join(HERE, "golden", "offset.bin"), "rb").read()


def test_decode_error_goldens(ctx):
    from libflate_b200 import native as nv
    for name in sorted(os.listdir(os.path.join(HERE, "golden"))):                                    # src/zlib.rs:799-837
        if name.startswith("issue16_crash-"):
            st, _, _, _ = ctx.decode(nv.FMT_ZLIB, open(os.path.join(HERE, "golden", name), "rb").read())
            assert st == nv.ERR_INVALID_DATA
    st, _, _, _ = ctx.decode(nv.FMT_DEFLATE, bytes(G["deflate_it_works_too_long"]["encoded"]))      # src/deflate/decode.rs:194-212
    assert st == nv.ERR_INVALID_DATA
    st, _, _, _ = ctx.decode(nv.FMT_DEFLATE, bytes(G["deflate_issue64"]["encoded"]))                # :216-220
    assert st != 0
    for k in (1, 2, 3):                                                                              # src/gzip.rs:1230-1247
        st, _, _, _ = ctx.decode(nv.FMT_GZIP, bytes(G[f"gzip_issue15_{k}"]["encoded"]))
        assert st != 0
    st, out, _, _ = ctx.decode(nv.FMT_ZLIB, bytes(G["zlib_issue71"]["encoded"]))                     # src/zlib.rs:917-934
    assert st != 0 and list(out) == G["zlib_issue71"]["partial"]
    st, _, _, _ = ctx.decode(nv.FMT_ZLIB, bytes([0, 0]))                                             # src/zlib.rs:938-943
    assert st == nv.ERR_INVALID_DATA


def test_decode_matches_oracle_valid_and_corrupt(ctx):
    rng = random.Random(40)
    d = _text(rng, 200000)
    streams, fmts = [], []
    for fmt in (0, 1, 2):
        enc = orc.encode(fmt, d, [8192] * 30)
        streams += [enc, enc[: len(enc) // 2], enc[:7], enc[:-1], enc + b"tail"]
        fmts += [fmt] * 5
        for _ in range(40):
            b = bytearray(enc)
            i = rng.randrange(len(b)); b[i] ^= 1 << rng.randrange(8)
            streams.append(bytes(b)); fmts.append(fmt)
    for lvl in (0, 1, 6, 9):
        streams.append(pyzlib.compress(d, lvl)); fmts.append(1)
    streams.append(pygzip.compress(d, 6, mtime=0)); fmts.append(2)
    for fmt in (0, 1, 2):
        idx = [i for i, f in enumerate(fmts) if f == fmt]
        res = ctx.decode_batch(fmt, [streams[i] for i in idx])
        for i, (st, out, used, _) in zip(idx, res):
            rc, want, wused, msg = orc.decode(fmt, streams[i])
            assert st == rc, (fmt, i, st, rc, msg)
            assert out == want, (fmt, i, len(out), len(want), msg)
            if rc == 0:
                assert used == wused


def test_decode_large_streams_block_parallel_path(ctx):
    """streams >= 128 KiB compressed go through the finder + speculative sub-block decode; anything irregular must fall
    back to the in-order kernel and still match the oracle / zlib bit for bit"""
    rng = random.Random(41)
    text = _text(rng, 6 << 20, nwords=5000)
    rnd = bytes(rng.getrandbits(8) for _ in range(1 << 20))
    mixed = text[: 1 << 20] + rnd[: 300000] + b"\x00" * 700000 + (b"abc" * 100000) + text[1 << 20: 2 << 20] + bytes([7]) * 500000
    cases = {
        "text_A": orc.encode(0, text, [8192] * (len(text) // 8192 + 1)),
        "text_single_write": orc.encode(0, text),                                   # one 6 MiB block
        "text_small_blocks": orc.encode(0, text, block_size=100000),
        "mixed": orc.encode(0, mixed, [8192] * (len(mixed) // 8192 + 1)),
        "random": orc.encode(0, rnd * 3, [8192] * 400),
        "zlib6_foreign": pyzlib.compress(text, 6)[2:-4],                            # cross-block references -> fallback
        "zlib1_foreign": pyzlib.compress(mixed, 1)[2:-4],
        "fixed_mode": orc.encode(0, text[: 2 << 20], mode=orc.MODE_FIXED),          # no dynamic headers -> fallback
        "stored_mode": orc.encode(0, text[: 1 << 20], mode=orc.MODE_STORED),
    }
    plain = {"text_A": text, "text_single_write": text, "text_small_blocks": text, "mixed": mixed, "random": rnd * 3,
             "zlib6_foreign": text, "zlib1_foreign": mixed, "fixed_mode": text[: 2 << 20], "stored_mode": text[: 1 << 20]}
    names = list(cases)
    before = ctx.stats()
    res = ctx.decode_batch(0, [cases[k] for k in names], caps=[len(plain[k]) + 64 for k in names])
    after = ctx.stats()
    # the four compressible libflate-style streams and the zlib-6 stream (dynamic blocks with cross-block references: markers)
    # must really take the parallel path; incompressible data (fixed-width codes never self-synchronise), fixed and stored
    # blocks fall back to the in-order kernel
    npar = after["decode_parallel_streams"] - before["decode_parallel_streams"]
    assert 5 <= npar <= 6, (before, after)
    assert after["decode_inorder_streams"] - before["decode_inorder_streams"] == 9 - npar
    for k, (st, out, used, _) in zip(names, res):
        assert st == 0, k
        assert out == plain[k], (k, len(out), len(plain[k]), next((i for i in range(min(len(out), len(plain[k]))) if out[i] != plain[k][i]), -1))
        assert used == len(cases[k]), k
    # corrupt one bit in the middle of a large stream: same status and partial output as the oracle
    b = bytearray(cases["text_A"]); b[len(b) // 2] ^= 0x10
    st, out, used, _ = ctx.decode(0, bytes(b), cap=len(text) + 64)
    rc, want, _, msg = orc.decode(0, bytes(b), cap=len(text) + 64)
    assert st == rc and out == want, (st, rc, len(out), len(want), msg)
    # too-small output for a large stream
    st, out, _, need = ctx.decode(0, cases["text_A"], cap=1 << 20)
    assert st == -3 and out == text[: 1 << 20] and need == len(text)


def test_decode_foreign_streams_take_the_parallel_path(ctx):
    """zlib / gzip output has cross-block back-references (the reference's own decode benchmark inflates flate2's output,
    flate_bench/src/main.rs:49-55): blocks are found and parsed speculatively like libflate's, references that reach before a
    segment become markers and are substituted in stream order (k_seg_resolve / k_seg_subst)."""
    from libflate_b200 import titles
    text = titles.generate(12 << 20, seed=5).tobytes()
    rng = random.Random(43)
    blockrep = bytes(rng.getrandbits(8) for _ in range(20000)) * 300             # period 20000: every match is long and far
    for name, plain, enc, fmt in (
            ("zlib6", text, pyzlib.compress(text, 6), 1), ("zlib9", text, pyzlib.compress(text, 9), 1),
            ("zlib1", text, pyzlib.compress(text, 1), 1), ("gzip6", text, pygzip.compress(text, 6, mtime=0), 2),
            ("raw6_blockrep", blockrep, pyzlib.compress(blockrep, 6)[2:-4], 0),
            ("raw6_zeros", b"\x00" * (40 << 20), pyzlib.compress(b"\x00" * (40 << 20), 6)[2:-4], 0)):
        before = ctx.stats()
        st, out, used, _ = ctx.decode(fmt, enc, cap=len(plain) + 64)
        after = ctx.stats()
        assert st == 0 and used == len(enc), (name, st, used, len(enc))
        assert out == plain, (name, len(out), next((i for i in range(min(len(out), len(plain))) if out[i] != plain[i]), -1))
        if len(enc) >= 128 * 1024 and name != "raw6_zeros":
            assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 1, name


def test_decode_truncated_and_corrupt_large_streams(ctx):
    """truncation / bit flips in streams >= 128 KiB compressed (the speculative path): status, partial output and consumed bytes
    must equal the oracle's, also right after the intact stream was decoded on the same context (stale device memory)"""
    rng = random.Random(44)
    text = _text(rng, 3 << 20, nwords=4000)
    enc = orc.encode(0, text, [8192] * (len(text) // 8192 + 1))
    assert len(enc) >= 256 * 1024
    st, out, used, _ = ctx.decode(0, enc, cap=len(text) + 64)
    assert st == 0 and out == text and used == len(enc)
    variants = [enc[:-1], enc[:-2], enc[:-3], enc[:-14], enc[: len(enc) // 2 + 1], enc[: 200 * 1024]]
    for _ in range(6):
        b = bytearray(enc); i = rng.randrange(len(b)); b[i] ^= 1 << rng.randrange(8); variants.append(bytes(b))
    for k, v in enumerate(variants):
        st, out, used, _ = ctx.decode(0, v, cap=len(text) + 64)
        rc, want, wused, msg = orc.decode(0, v, cap=len(text) + 64)
        assert st == rc, (k, st, rc, msg)
        assert out == want, (k, len(out), len(want), msg)
        if rc == 0:
            assert used == wused, k


def test_decode_multi_member_gzip_with_large_members(ctx):
    rng = random.Random(45)
    parts = [_text(rng, n, nwords=3000) for n in (1 << 20, 700000, 5000, 1500000)]
    members = [orc.encode(2, parts[0], [8192] * 200, mtime=1), pygzip.compress(parts[1], 6, mtime=2), orc.encode(2, parts[2], mtime=3),
               pygzip.compress(parts[3], 9, mtime=4)]
    blob = b"".join(members)
    st, out, used, _ = ctx.decode(3, blob, cap=sum(map(len, parts)) + 64)
    assert st == 0 and out == b"".join(parts) and used == len(blob)
    st, out, used, _ = ctx.decode(2, blob, cap=sum(map(len, parts)) + 64)          # gzip::Decoder: first member only
    assert st == 0 and out == parts[0] and used == len(members[0])


def test_decode_many_small_members(ctx):
    """BGZF / pigz -i shape: hundreds of small gzip members back to back (MultiDecoder, src/gzip.rs:1052-1167); later members are
    sized after their predecessor, so the rest of the file is not re-scanned for every member"""
    rng = random.Random(46)
    parts = [_text(rng, rng.randint(20000, 60000)) for _ in range(120)] + [_text(rng, 1 << 20, nwords=3000)] + [_text(rng, 30000) for _ in range(5)]
    blob = b"".join(pygzip.compress(p, 6, mtime=0) if i % 2 else orc.encode(2, p, mtime=0) for i, p in enumerate(parts))
    st, out, used, _ = ctx.decode(3, blob, cap=sum(map(len, parts)) + 64)
    assert st == 0 and out == b"".join(parts) and used == len(blob)


def test_pageable_and_page_locked_callers_get_the_same_bytes(ctx):
    """pinned-memory registration is internal -- ordinary memory is staged by the library, b2f_host_alloc memory is
    used in place; both give the oracle's bytes"""
    from libflate_b200 import native, titles
    d = titles.generate(24 << 20, seed=11)
    sched = np.full(d.size // 8192 + 1, 8192, dtype=np.int64)
    want = orc.encode(orc.FMT_GZIP, d.tobytes(), sched.tolist(), mtime=0)
    cap = native.lib().b2f_encode_bound(d.size, len(sched), None)
    s0 = ctx.stats()
    e_page = np.empty(cap, dtype=np.uint8)
    m = ctx.encode_into(native.FMT_GZIP, d, e_page, sched, mtime=0)
    s1 = ctx.stats()
    assert e_page[:m].tobytes() == want
    assert s1["staged_h2d_bytes"] - s0["staged_h2d_bytes"] == d.size and s1["staged_d2h_bytes"] - s0["staged_d2h_bytes"] == m
    h_in, h_enc, h_dec = native.host_alloc(d.size), native.host_alloc(cap), native.host_alloc(d.size + 64)
    try:
        h_in[:] = d
        m2 = ctx.encode_into(native.FMT_GZIP, h_in, h_enc, sched, mtime=0)
        s2 = ctx.stats()
        assert m2 == m and h_enc[:m].tobytes() == want
        assert s2["staged_h2d_bytes"] == s1["staged_h2d_bytes"] and s2["staged_d2h_bytes"] == s1["staged_d2h_bytes"]     # in place
        dl, used, st = ctx.decode_into(native.FMT_GZIP, h_enc, m, h_dec)
        assert st == 0 and dl == d.size and used == m and np.array_equal(h_dec[:dl], d)
        p_dec = np.empty(d.size + 64, dtype=np.uint8)
        dl, used, st = ctx.decode_into(native.FMT_GZIP, e_page, m, p_dec)
        assert st == 0 and dl == d.size and used == m and np.array_equal(p_dec[:dl], d)
        assert ctx.stats()["staged_d2h_bytes"] - s2["staged_d2h_bytes"] == d.size
    finally:
        for a in (h_in, h_enc, h_dec):
            native.host_free(a)


def test_page_locked_input_arrives_in_pieces(ctx):
    """b2f_decode_batch with page-locked inputs of >= 8 MiB in all: the H2D copy runs in pieces on its own stream and the block
    finder follows it piece by piece (pieces cross stream boundaries; a small and an incompressible stream ride along)"""
    import ctypes as C
    from libflate_b200 import native, titles
    rng = random.Random(77)
    plains = [titles.generate(30 << 20, seed=21).tobytes(), titles.generate(12 << 20, seed=22).tobytes(), _text(rng, 1000),
              bytes(rng.getrandbits(8) for _ in range(1 << 20)) * 2]
    encs = [orc.encode(orc.FMT_GZIP, p, [8192] * (len(p) // 8192 + 1), mtime=0) for p in plains]
    assert sum(map(len, encs)) >= (8 << 20)
    n = len(encs)
    h_in = [native.host_alloc(len(e)) for e in encs]
    h_out = [native.host_alloc(len(p) + 64) for p in plains]
    try:
        for a, e in zip(h_in, encs):
            a[:] = np.frombuffer(e, dtype=np.uint8)
        in_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in h_in]); in_len = (C.c_size_t * n)(*[len(e) for e in encs])
        out_ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in h_out]); out_cap = (C.c_size_t * n)(*[a.size for a in h_out])
        out_len, used, status = (C.c_size_t * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        for rep in range(2):                       # the second call reuses the context's feed events
            for a in h_out:
                a[:] = 0
            rc = native.lib().b2f_decode_batch(ctx._h, native.FMT_GZIP, n, in_ptrs, in_len, out_ptrs, out_cap, out_len, used, status)
            assert rc == 0
            for i in range(n):
                assert status[i] == 0 and out_len[i] == len(plains[i]) and used[i] == len(encs[i]), (rep, i, status[i], out_len[i])
                assert h_out[i][:out_len[i]].tobytes() == plains[i], (rep, i)
        # a truncated big stream in page-locked memory: same answer as the oracle (the pieces must not hide the error path)
        cut = len(encs[0]) - 3000
        in_len[0] = cut
        rc = native.lib().b2f_decode_batch(ctx._h, native.FMT_GZIP, n, in_ptrs, in_len, out_ptrs, out_cap, out_len, used, status)
        assert rc == 0
        orc_rc, orc_out, _, _ = orc.decode(orc.FMT_GZIP, encs[0][:cut], cap=len(plains[0]) + 64)
        assert status[0] == orc_rc != 0 and status[1] == 0 and h_out[1][:out_len[1]].tobytes() == plains[1]
    finally:
        for a in h_in + h_out:
            native.host_free(a)


def test_decode_output_too_small(ctx):
    d = b"abcabcabc" * 5000
    enc = orc.encode(orc.FMT_ZLIB, d)
    st, out, _, need = ctx.decode(1, enc, cap=1000)
    assert st == -3 and out == d[:1000] and need == len(d)


# ------------------------------------------------------------------------------------------ streaming handles (Read/Write surface)
def test_streaming_encoder_decoder_handles(ctx):
    from libflate_b200.deflate import Encoder, Decoder
    from libflate_b200 import gzip as bgzip, zlib as bzlib
    rng = random.Random(50)
    d = _text(rng, 300000)
    e = Encoder(ctx)                                                                                 # src/deflate/encode.rs:132-258
    for k in range(0, len(d), 50000):
        assert e.write(d[k: k + 50000]) == len(d[k: k + 50000])
    e.flush()
    enc = e.finish()
    assert enc == orc.encode(0, d, [50000] * 6 + [-1])
    dec = Decoder(ctx, enc)                                                                          # src/deflate/decode.rs:8-165
    got = bytearray()
    while True:
        chunk = dec.read(7777)
        if not chunk:
            break
        got += chunk
    assert bytes(got) == d
    ge = bgzip.Encoder(ctx, mtime=9)
    ge.write_all(d)
    genc = ge.finish()
    assert genc == orc.encode(2, d, [len(d)], mtime=9)
    assert bgzip.Decoder(ctx, genc).read_to_end() == d
    assert bgzip.MultiDecoder(ctx, genc * 3).read_to_end() == d * 3
    ze = bzlib.Encoder(ctx)
    ze.write_all(d)
    zenc = ze.finish()
    assert pyzlib.decompress(zenc) == d and bzlib.Decoder(ctx, zenc).read_to_end() == d
    # error surface: io::ErrorKind::InvalidData -> exception, partial data via unread_decoded_data()
    bad = bzlib.Decoder(ctx, bytes(G["zlib_issue71"]["encoded"]))
    with pytest.raises(Exception):
        bad.read_to_end()
    assert list(bad.unread_decoded_data()) == G["zlib_issue71"]["partial"]


def test_decoder_read_hands_out_complete_blocks_before_the_error(ctx):
    """Decoder::read (src/deflate/decode.rs:136-164) decodes block by block: the blocks before a corrupt one are served by read(),
    the error surfaces when the corrupt block is reached and its partial bytes stay in unread_decoded_data()"""
    from libflate_b200.deflate import Decoder
    rng = random.Random(51)
    d = _text(rng, 300000)
    enc = orc.encode(0, d, [100000] * 3, block_size=100000)        # three 100000-byte blocks + the empty final block
    enc = enc[: len(enc) - 2000]                                   # cut inside the third block: UnexpectedEof there
    rc, partial, _, _ = orc.decode(0, bytes(enc))
    assert rc != 0 and 200000 <= len(partial) < 300000
    dec = Decoder(ctx, bytes(enc))
    got = bytearray()
    with pytest.raises(Exception):
        while True:
            c = dec.read(70000)
            assert c
            got += c
    assert bytes(got) == d[:200000]                                # exactly the two complete blocks
    assert bytes(got) + dec.unread_decoded_data() == partial       # the failing block's partial output


# ------------------------------------------------------------------------------------------ BASELINE-shaped cases (reduced sizes; full sizes in bench.py)
def test_config2_shape_many_streams_raw_deflate(ctx):
    from libflate_b200 import titles
    datas = [titles.segment(1000 + i, 1 << 20).ljust(1 << 20, b"\n") for i in range(6)]
    encs = ctx.encode_batch(0, datas)
    for i, (d, e) in enumerate(zip(datas, encs)):
        if i < 2:
            assert e == orc.encode(0, d)
        assert pyzlib.decompress(e, -15) == d
    res = ctx.decode_batch(0, encs)
    assert all(st == 0 and out == d for (st, out, _, _), d in zip(res, datas))


def test_config3_shape_gzip_roundtrip_schedule_a(ctx):
    from libflate_b200 import titles
    d = titles.generate(5 * (1 << 20) + 12345, seed=42, workers=1).tobytes()
    sched = [8192] * (len(d) // 8192 + 1)
    enc = ctx.encode(2, d, sched, mtime=0)
    assert enc == orc.encode(2, d, sched, mtime=0)
    st, out, used, _ = ctx.decode(2, enc)
    assert st == 0 and out == d and used == len(enc)
    assert pygzip.decompress(enc) == d


def test_config4_shape_zlib_streams(ctx):
    from libflate_b200 import titles
    datas = [titles.segment(2000 + i, 1 << 18) for i in range(16)]
    encs = ctx.encode_batch(1, datas)
    for d, e in zip(datas, encs):
        assert e == orc.encode(1, d)
        assert e[-4:] == pyzlib.adler32(d).to_bytes(4, "big")


@pytest.mark.gpu
def test_cli_mirror_of_examples_flate(tmp_path):
    """examples/flate.rs:84-111: gzip-encode / gzip-decode / gzip-decode-multi / zlib-encode / zlib-decode over files.  The
    encoders are fed by an 8 KiB copy loop, so the bytes equal the oracle's with schedule [8192]*k."""
    from libflate_b200 import flate, titles
    d = titles.generate(700_001, seed=77).tobytes()
    src = tmp_path / "in.txt"
    src.write_bytes(d)
    sched = [8192] * (len(d) // 8192) + ([len(d) % 8192] if len(d) % 8192 else [])
    gz, zz, back = tmp_path / "o.gz", tmp_path / "o.z", tmp_path / "back"
    assert flate.main(["-i", str(src), "-o", str(gz), "--mtime", "0", "gzip-encode"]) == 0
    assert gz.read_bytes() == orc.encode(orc.FMT_GZIP, d, sched, mtime=0)
    assert flate.main(["-i", str(src), "-o", str(zz), "zlib-encode"]) == 0
    assert zz.read_bytes() == orc.encode(orc.FMT_ZLIB, d, sched)
    assert pyzlib.decompress(zz.read_bytes()) == d
    assert flate.main(["-i", str(gz), "-o", str(back), "-v", "gzip-decode"]) == 0 and back.read_bytes() == d
    assert flate.main(["-i", str(zz), "-o", str(back), "zlib-decode"]) == 0 and back.read_bytes() == d
    two = tmp_path / "two.gz"
    two.write_bytes(gz.read_bytes() + pygzip.compress(b"second member"))
    assert flate.main(["-i", str(two), "-o", str(back), "gzip-decode-multi"]) == 0 and back.read_bytes() == d + b"second member"
    assert flate.main(["-i", str(two), "-o", str(back), "gzip-decode"]) == 0 and back.read_bytes() == d


@pytest.mark.gpu
def test_stored_mode_is_assembled_on_the_device(ctx):
    """EncodeOptions::no_compression (encode.rs:354-383): block headers, payload, sync markers and trailers are laid out by the
    host and copied by one kernel -- host API and device API, all three containers, schedules with flushes, too-small output"""
    import torch
    from libflate_b200 import native, titles
    d = titles.generate(700_001, seed=9).tobytes()
    sched = [100_000, 65_535, -1, 1, 200_000, -1, -1, 334_465]
    for fmt, kw in ((orc.FMT_DEFLATE, {}), (orc.FMT_ZLIB, dict(zlib_flush_sync=1)), (orc.FMT_GZIP, dict(mtime=7)), (orc.FMT_ZLIB, dict(block_size=1000))):
        want = orc.encode(fmt, d, sched, mode=orc.MODE_STORED, **kw)
        assert ctx.encode(fmt, d, sched, mode=native.MODE_STORED, **kw) == want, (fmt, kw)
        rc, out, used, _ = orc.decode(fmt, want)
        assert rc == 0 and out == d
    # device API: two streams at odd offsets, the second one with too little room
    datas = [d, d[:70_000]]
    want = [orc.encode(orc.FMT_GZIP, x, mode=orc.MODE_STORED, mtime=0) for x in datas]
    in_off = [1, len(d) + 6]
    d_in = torch.zeros(in_off[1] + len(datas[1]) + 64, dtype=torch.uint8, device="cuda")
    for o, x in zip(in_off, datas):
        d_in[o:o + len(x)] = torch.frombuffer(bytearray(x), dtype=torch.uint8).cuda()
    caps = [len(want[0]) + 10, len(want[1]) - 1]
    e_off = [3, 3 + caps[0] + 5]
    d_enc = torch.full((e_off[1] + caps[1] + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    ol, st = ctx.encode_device(native.FMT_GZIP, d_in.data_ptr(), in_off, [len(x) for x in datas], d_enc.data_ptr(), e_off, caps, None, mode=native.MODE_STORED, mtime=0)
    assert st == [0, -3] and ol == [len(want[0]), len(want[1])]
    host = d_enc.cpu().numpy()
    assert bytes(host[e_off[0]:e_off[0] + ol[0]]) == want[0] and host[e_off[0] - 1] == 0xEE and host[e_off[0] + ol[0]] == 0xEE
    assert (host[e_off[1]:e_off[1] + caps[1]] == 0xEE).all()                  # the stream that does not fit is not written at all


@pytest.mark.gpu
def test_device_api_unaligned_offsets(ctx):
    """b2f_encode_device / b2f_decode_device with odd byte offsets on both sides (the resolve kernel writes aligned words only
    where a word is wholly owned by one unit)."""
    import torch
    from libflate_b200 import native, titles
    datas = [titles.generate(n, seed=s).tobytes() for n, s in ((1_500_001, 5), (1_200_003, 6), (70_001, 7))]
    scheds = [[8192] * (len(d) // 8192 + 1) for d in datas]
    want = [orc.encode(orc.FMT_ZLIB, d, sc) for d, sc in zip(datas, scheds)]
    in_off, pos = [], 1
    for d in datas:
        in_off.append(pos); pos += len(d) + 3
    d_in = torch.zeros(pos + 64, dtype=torch.uint8, device="cuda")
    for o, d in zip(in_off, datas):
        d_in[o:o + len(d)] = torch.frombuffer(bytearray(d), dtype=torch.uint8).cuda()
    caps = [len(d) + len(d) // 8 + 4096 for d in datas]
    e_off, pos = [], 3
    for c in caps:
        e_off.append(pos); pos += c + 1
    d_enc = torch.zeros(pos + 64, dtype=torch.uint8, device="cuda")
    ol, st = ctx.encode_device(native.FMT_ZLIB, d_in.data_ptr(), in_off, [len(d) for d in datas], d_enc.data_ptr(), e_off, caps, scheds)
    assert st == [0, 0, 0]
    for o, n, w in zip(e_off, ol, want):
        assert bytes(d_enc[o:o + n].cpu().numpy()) == w
    o_off, pos = [], 1
    for d in datas:
        o_off.append(pos); pos += len(d) + 65 + 2
    d_out = torch.full((pos + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    dl, used, st = ctx.decode_device(native.FMT_ZLIB, d_enc.data_ptr() + 0, e_off, ol, d_out.data_ptr() + 1, [o - 1 for o in o_off], [len(d) + 65 for d in datas])
    assert st == [0, 0, 0] and used == ol and dl == [len(d) for d in datas]
    host = d_out.cpu().numpy()
    for o, d in zip(o_off, datas):
        assert bytes(host[o:o + len(d)]) == d
        assert host[o - 1] == 0xEE and host[o + len(d)] == 0xEE          # nothing written outside the streams

"""Pins the CPU oracle (oracle/flate_oracle.c) against every golden vector sile/libflate's own tests hold for
the DEFLATE hot path.  Fixtures: tests/golden/ (made by tests/golden/make_goldens.py)."""
import gzip as pygzip
import json
import os
import random
import zlib as pyzlib

import pytest

from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "goldens.json")))


def gbin(name):
    return open(os.path.join(HERE, "golden", name), "rb").read()


# ---------------------------------------------------------------- encode goldens (exact bytes)
def test_deflate_hello_dynamic():       # src/deflate/encode.rs:152-154
    g = G["deflate_hello_dynamic"]
    assert list(orc.encode(orc.FMT_DEFLATE, g["plain"].encode())) == g["bytes"]


def test_deflate_hello_stored():        # src/deflate/encode.rs:178-180
    g = G["deflate_hello_stored"]
    assert list(orc.encode(orc.FMT_DEFLATE, g["plain"].encode(), mode=orc.MODE_STORED)) == g["bytes"]


def test_zlib_hello_default():          # src/zlib.rs:547-549
    g = G["zlib_hello_default"]
    assert list(orc.encode(orc.FMT_ZLIB, g["plain"].encode())) == g["bytes"]


def test_zlib_raw():                    # src/zlib.rs:750-764
    g = G["zlib_raw_encode"]
    assert list(orc.encode(orc.FMT_ZLIB, g["plain"].encode(), mode=orc.MODE_STORED)) == g["bytes"]


def test_gzip_stored_mtime():           # src/gzip.rs:800-802
    g = G["gzip_stored_mtime123"]
    assert list(orc.encode(orc.FMT_GZIP, g["plain"].encode(), mode=orc.MODE_STORED, mtime=g["mtime"])) == g["bytes"]


@pytest.mark.parametrize("sync", [False, True])
def test_zlib_issue27(sync):            # src/zlib.rs:840-902 : 3 writes + flush, twice
    g = G["zlib_issue27_sync" if sync else "zlib_issue27_none"]
    writes = [w.encode() for w in g["writes"]]
    data = b"".join(writes) * 2
    sched = ([len(w) for w in writes] + [orc.FLUSH]) * 2
    enc = orc.encode(orc.FMT_ZLIB, data, sched, zlib_flush_sync=sync)
    assert list(enc) == g["bytes"]
    rc, out, used, _ = orc.decode(orc.FMT_ZLIB, enc)
    assert rc == 0 and out == data and used == len(enc)
    assert pyzlib.decompress(enc) == data


def test_lz77_issue21():                # src/lz77.rs:16-31
    codes = orc.lz77_default(b"aaaaa")
    assert list(codes) == [97, 0x80000000 | (4 << 16) | 1]


def test_issue52_size_bound():          # src/deflate/encode.rs:435-457
    data = gbin("issue52_input.bin")
    for lim in (16031, 16032):
        enc = orc.encode(orc.FMT_DEFLATE, data[:lim])
        assert len(enc) < lim
        assert pyzlib.decompress(enc, -15) == data[:lim]


def test_checksum_kats():               # src/checksum.rs:45-56
    g = G["checksum_kat"]
    assert orc.crc32(g["input"].encode()) == g["crc32"]
    assert orc.adler32(g["input"].encode()) == g["adler32"]


# ---------------------------------------------------------------- decode goldens
def test_deflate_fixed_hello():         # src/deflate/decode.rs:28-33
    g = G["deflate_fixed_hello"]
    rc, out, used, _ = orc.decode(orc.FMT_DEFLATE, bytes(g["bytes"]))
    assert rc == 0 and out == g["plain"].encode() and used == len(g["bytes"])


def test_zlib_decode_works():           # src/zlib.rs:708-730
    g = G["zlib_decode_works"]
    rc, out, _, _ = orc.decode(orc.FMT_ZLIB, bytes(g["bytes"]))
    assert rc == 0 and out == g["plain"].encode()


def test_gzip_multi_member():           # src/gzip.rs:1217-1226
    one = orc.encode(orc.FMT_GZIP, b"Hello World!")
    rc, out, used, _ = orc.decode(orc.FMT_GZIP, one * 2)
    assert rc == 0 and out == b"Hello World!" and used == len(one)
    rc, out, used, _ = orc.decode(orc.FMT_GZIP_MULTI, one * 2)
    assert rc == 0 and out == b"Hello World!Hello World!" and used == 2 * len(one)


def test_offset_gz():                   # src/non_blocking/gzip.rs:178-183 (stored-block alignment)
    rc, out, _, _ = orc.decode(orc.FMT_GZIP, gbin("offset.gz"))
    assert rc == 0 and out == gbin("offset.bin")


def test_roundtrip_1mib_counter():      # src/deflate/mod.rs:49-64
    plain = bytes(i & 255 for i in range(32768 * 32))
    enc = orc.encode(orc.FMT_DEFLATE, plain)
    rc, out, used, _ = orc.decode(orc.FMT_DEFLATE, enc)
    assert rc == 0 and out == plain and used == len(enc)
    assert pyzlib.decompress(enc, -15) == plain


def test_issue2_roundtrips():           # src/zlib.rs:778-796
    for arr in G["zlib_issue2_inputs"]["inputs"]:
        enc = orc.encode(orc.FMT_ZLIB, bytes(arr))
        rc, out, _, _ = orc.decode(orc.FMT_ZLIB, enc)
        assert rc == 0 and out == bytes(arr)
        assert pyzlib.decompress(enc) == bytes(arr)


def test_test_n_roundtrip():            # src/non_blocking/deflate/decode.rs:273-286
    plain = "".join(f"test {i}" for i in range(10000)).encode()
    enc = orc.encode(orc.FMT_DEFLATE, plain)
    rc, out, _, _ = orc.decode(orc.FMT_DEFLATE, enc)
    assert rc == 0 and out == plain


# ---------------------------------------------------------------- error goldens
def test_issues_16_hdist():             # src/zlib.rs:799-837
    for name in sorted(os.listdir(os.path.join(HERE, "golden"))):
        if name.startswith("issue16_crash-"):
            rc, _, _, msg = orc.decode(orc.FMT_ZLIB, gbin(name))
            assert rc == orc.INVALID_DATA
            assert msg.startswith("The value of HDIST is too big: max=30")


def test_issue3_header_loads():         # src/deflate/decode.rs:176-190
    assert orc.dynamic_header_loads(bytes(G["deflate_issue3_header"]["encoded"])) == 0


def test_it_works_too_long_backref():   # src/deflate/decode.rs:194-212
    rc, _, _, msg = orc.decode(orc.FMT_DEFLATE, bytes(G["deflate_it_works_too_long"]["encoded"]))
    assert rc == orc.INVALID_DATA and msg.startswith("Too long backword reference")


def test_issue64_errors():              # src/deflate/decode.rs:216-220
    rc, _, _, _ = orc.decode(orc.FMT_DEFLATE, bytes(G["deflate_issue64"]["encoded"]))
    assert rc != 0


@pytest.mark.parametrize("k", [1, 2, 3])
def test_gzip_issue15(k):               # src/gzip.rs:1230-1247
    rc, _, _, _ = orc.decode(orc.FMT_GZIP, bytes(G[f"gzip_issue15_{k}"]["encoded"]))
    assert rc != 0


def test_zlib_issue71_partial_output(): # src/zlib.rs:917-934
    g = G["zlib_issue71"]
    rc, out, _, _ = orc.decode(orc.FMT_ZLIB, bytes(g["encoded"]))
    assert rc != 0
    assert list(out) == g["partial"]


def test_zlib_method0():                # src/zlib.rs:938-943
    rc, _, _, msg = orc.decode(orc.FMT_ZLIB, bytes([0, 0]))
    assert rc == orc.INVALID_DATA and "method=0" in msg


# ---------------------------------------------------------------- independent cross-checks (python zlib as 2nd inflate)
def _rand_text(rng, n):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    out = bytearray()
    while len(out) < n:
        out += rng.choice(words) + rng.choice([b" ", b"_", b"\n"])
    return bytes(out[:n])


@pytest.mark.parametrize("seed", range(6))
def test_random_roundtrip_vs_zlib(seed):
    rng = random.Random(seed)
    n = rng.choice([0, 1, 2, 3, 4, 5, 100, 4096, 32767, 32768, 32769, 70000, 300000])
    data = _rand_text(rng, n) if seed % 2 == 0 else bytes(rng.getrandbits(8) >> rng.choice([0, 4, 6]) for _ in range(n))
    for sched in (None, [7] * (n // 7 + 1), [8192] * (n // 8192 + 1)):
        for fmt, undo in ((orc.FMT_DEFLATE, lambda b: pyzlib.decompress(b, -15)), (orc.FMT_ZLIB, pyzlib.decompress),
                          (orc.FMT_GZIP, pygzip.decompress)):
            enc = orc.encode(fmt, data, sched)
            assert undo(enc) == data
            rc, out, used, _ = orc.decode(fmt, enc)
            assert rc == 0 and out == data and used == len(enc)


def test_decode_foreign_zlib_streams():
    rng = random.Random(99)
    data = _rand_text(rng, 200000)
    for level in (0, 1, 6, 9):
        enc = pyzlib.compress(data, level)
        rc, out, used, _ = orc.decode(orc.FMT_ZLIB, enc)
        assert rc == 0 and out == data and used == len(enc)
    co = pyzlib.compressobj(6, pyzlib.DEFLATED, -15, 8, pyzlib.Z_FIXED)
    enc = co.compress(data) + co.flush()
    rc, out, _, _ = orc.decode(orc.FMT_DEFLATE, enc)
    assert rc == 0 and out == data


def test_truncated_inputs_report_eof():
    enc = orc.encode(orc.FMT_GZIP, b"hello hello hello hello" * 50)
    for cut in (0, 5, 10, 15, len(enc) - 9, len(enc) - 1):
        rc, _, _, _ = orc.decode(orc.FMT_GZIP, enc[:cut])
        assert rc == orc.UNEXPECTED_EOF, cut


def test_fixed_and_window_options_roundtrip():
    rng = random.Random(5)
    data = _rand_text(rng, 50000)
    enc = orc.encode(orc.FMT_DEFLATE, data, mode=orc.MODE_FIXED)
    assert pyzlib.decompress(enc, -15) == data
    enc = orc.encode(orc.FMT_ZLIB, data, window_size=1024, max_length=32, block_size=5000)
    assert enc[:2] == bytes([0x28, 0x95]) or (enc[0] >> 4) == 2
    assert pyzlib.decompress(enc) == data

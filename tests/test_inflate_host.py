"""CPU checks of the product's inflate core (libflate_b200/csrc/inflate_core.cuh compiled for the host by
tests/native) against the oracle: valid streams, the reference's error goldens, truncations and bit flips."""
import json
import os
import random
import zlib as pyzlib

import pytest

import hostcheck as hc
from oracle import oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "goldens.json")))


def _text(rng, n):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz_") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b"\n"
    return bytes(b[:n])


def _same_as_oracle(stream, check_consumed=True):
    rc_o, out_o, used_o, msg = orc.decode(orc.FMT_DEFLATE, stream)
    rc_p, out_p, used_p, _ = hc.inflate(stream)
    assert rc_p == rc_o, (rc_p, rc_o, msg)
    assert out_p == out_o, (len(out_p), len(out_o), msg)
    if rc_o == 0 and check_consumed:
        assert used_p == used_o


def test_len_dist_tables():                       # symbol.rs:22-87
    LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
    LEN_EXTRA = [0] * 8 + [1] * 4 + [2] * 4 + [3] * 4 + [4] * 4 + [5] * 4 + [0]
    DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                 4097, 6145, 8193, 12289, 16385, 24577]
    L = hc.lib()
    for k in range(29):
        v = L.hc_len_base(k)
        assert (v & 0xFFFF, v >> 16) == (LEN_BASE[k], LEN_EXTRA[k])
    for k in range(30):
        v = L.hc_dist_base(k)
        assert (v & 0xFFFF, v >> 16) == (DIST_BASE[k], max(0, (k - 2) // 2))


def test_valid_streams_from_oracle_and_zlib():
    rng = random.Random(3)
    datas = [b"", b"a", b"Hello World!", _text(rng, 1000), _text(rng, 300000), bytes(rng.getrandbits(8) for _ in range(70000)),
             b"\x00" * 200000, bytes(i & 255 for i in range(1 << 20))]
    for d in datas:
        for enc in (orc.encode(orc.FMT_DEFLATE, d), orc.encode(orc.FMT_DEFLATE, d, mode=orc.MODE_FIXED),
                    orc.encode(orc.FMT_DEFLATE, d, mode=orc.MODE_STORED), orc.encode(orc.FMT_DEFLATE, d, [4096] * (len(d) // 4096 + 1)),
                    orc.encode(orc.FMT_DEFLATE, d, block_size=10000)):
            rc, out, used, _ = hc.inflate(enc)
            assert rc == 0 and out == d and used == len(enc)
        for level in (1, 6, 9):
            co = pyzlib.compressobj(level, pyzlib.DEFLATED, -15)
            enc = co.compress(d) + co.flush()
            rc, out, used, _ = hc.inflate(enc)
            assert rc == 0 and out == d and used == len(enc)
            _same_as_oracle(enc)


def test_reference_error_goldens():
    _same_as_oracle(bytes(G["deflate_it_works_too_long"]["encoded"]))     # src/deflate/decode.rs:194-212
    _same_as_oracle(bytes(G["deflate_issue64"]["encoded"]))               # :216-220
    _same_as_oracle(bytes(G["deflate_fixed_hello"]["bytes"]))
    for name in sorted(os.listdir(os.path.join(HERE, "golden"))):          # src/zlib.rs:799-837 (skip the 2-byte zlib header)
        if name.startswith("issue16_crash-"):
            _same_as_oracle(open(os.path.join(HERE, "golden", name), "rb").read()[2:])
    _same_as_oracle(bytes(G["zlib_issue71"]["encoded"])[2:])               # src/zlib.rs:917-934 (partial output)
    for k in (1, 2, 3):                                                    # src/gzip.rs:1230-1247 (10-byte gzip header, FLG has no fields)
        _same_as_oracle(bytes(G[f"gzip_issue15_{k}"]["encoded"])[10:])


def test_trailing_bytes_are_not_consumed():
    enc = orc.encode(orc.FMT_DEFLATE, b"hello hello hello")
    rc, out, used, _ = hc.inflate(enc + b"\x01\x02\x03\x04TRAILER")
    assert rc == 0 and out == b"hello hello hello" and used == len(enc)


@pytest.mark.parametrize("seed", range(4))
def test_truncations_match_oracle(seed):
    rng = random.Random(100 + seed)
    d = _text(rng, 5000)
    encs = [orc.encode(orc.FMT_DEFLATE, d), orc.encode(orc.FMT_DEFLATE, d, mode=orc.MODE_FIXED),
            orc.encode(orc.FMT_DEFLATE, d, mode=orc.MODE_STORED), pyzlib.compress(d, 6)[2:-4]]
    for enc in encs:
        cuts = set(range(0, min(len(enc), 80))) | {rng.randrange(len(enc)) for _ in range(60)} | {len(enc) - 1, len(enc) - 2}
        for cut in sorted(c for c in cuts if 0 <= c < len(enc)):
            _same_as_oracle(enc[:cut])


@pytest.mark.parametrize("seed", range(4))
def test_bitflips_match_oracle(seed):
    rng = random.Random(200 + seed)
    d = _text(rng, 3000)
    encs = [orc.encode(orc.FMT_DEFLATE, d), pyzlib.compress(d, 6)[2:-4], orc.encode(orc.FMT_DEFLATE, d, mode=orc.MODE_FIXED)]
    for enc in encs:
        for _ in range(250):
            b = bytearray(enc)
            for _ in range(rng.choice([1, 1, 2, 5])):
                i = rng.randrange(min(len(b), 200) if rng.random() < 0.7 else len(b))
                b[i] ^= 1 << rng.randrange(8)
            _same_as_oracle(bytes(b), check_consumed=False)


def test_random_garbage_matches_oracle():
    rng = random.Random(77)
    for _ in range(600):
        n = rng.randrange(1, 120)
        b = bytearray(rng.getrandbits(8) for _ in range(n))
        b[0] = (b[0] & ~7) | rng.choice([0b100, 0b101, 0b010, 0b011, 0b000, 0b001, 0b110])   # steer BTYPE
        _same_as_oracle(bytes(b), check_consumed=False)


def test_output_capacity():
    d = b"abcabcabcabc" * 1000
    enc = orc.encode(orc.FMT_DEFLATE, d)
    rc, out, _, _ = hc.inflate(enc, cap=100)
    assert rc == -3 and out == d[:100]

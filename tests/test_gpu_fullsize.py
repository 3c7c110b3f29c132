"""BASELINE.json configurations at (or near) full size, checked through size-independent properties:
encode -> decode round trips, independent inflate by Python's zlib, checksums, and spot comparisons with the oracle."""
import zlib as pyzlib

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
CACHE = "/tmp/b2f_test_cache"


@pytest.fixture(scope="module")
def ctx():
    from libflate_b200 import native
    c = native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def titles256():
    from libflate_b200 import titles
    return titles.generate(256 << 20, seed=1000, cache_dir=CACHE)


def test_config2_64_streams_of_4mib_raw_deflate(ctx, titles256):
    """64 x 4 MiB, one write_all each => one 4 MiB LZ77 chunk (32 chain segments, 4096 parse tiles) + one block + empty final block"""
    streams = [titles256[i << 22:(i + 1) << 22] for i in range(64)]
    encs = ctx.encode_batch(0, streams)
    assert encs[0] == orc.encode(0, streams[0].tobytes())                     # bit-exact against the oracle
    assert encs[63] == orc.encode(0, streams[63].tobytes())
    for i in (1, 17, 40):
        assert pyzlib.decompress(encs[i], -15) == streams[i].tobytes()        # independent inflate
    before = ctx.stats()
    res = ctx.decode_batch(0, encs, caps=[(4 << 20) + 64] * 64)
    after = ctx.stats()
    assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 64
    for (st, out, used, _), s, e in zip(res, streams, encs):
        assert st == 0 and used == len(e) and out == s.tobytes()


def test_config3_full_size_gzip_round_trip(ctx):
    from libflate_b200 import native, titles
    n = 277_303_937
    d = titles.generate(n, seed=42, cache_dir=CACHE)
    sched = np.asarray([8192] * (n // 8192 + 1), dtype=np.int64)
    enc = np.empty(native.lib().b2f_encode_bound(n, len(sched), None), dtype=np.uint8)
    m = ctx.encode_into(native.FMT_GZIP, d, enc, sched, mtime=0)
    # container framing + checksum (CRC-32 LE, ISIZE LE mod 2^32)
    assert bytes(enc[:10]) == bytes([31, 139, 8, 0, 0, 0, 0, 0, 0, 3])
    assert int.from_bytes(bytes(enc[m - 8:m - 4]), "little") == pyzlib.crc32(d) and int.from_bytes(bytes(enc[m - 4:m]), "little") == n
    # first 1 MiB block is bit-exact against the oracle (blocks are independent: same bytes up to the end of block 0)
    want0 = orc.encode(orc.FMT_GZIP, d[: 1 << 20].tobytes(), [8192] * 128, mtime=0)
    k = len(want0) - 8 - 12 - 2                                              # minus trailer, minus the oracle's own final empty block (bit-shifted tail)
    assert bytes(enc[:k]) == want0[:k]
    dec = np.empty(n + 64, dtype=np.uint8)
    before = ctx.stats()
    dl, used, st = ctx.decode_into(native.FMT_GZIP, enc, m, dec)
    after = ctx.stats()
    assert st == 0 and dl == n and used == m and np.array_equal(dec[:n], d)
    assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 1     # the 265-block stream took the parallel path
    assert pyzlib.crc32(pyzlib.decompress(bytes(enc[:m]), 31)[: 1 << 20]) == pyzlib.crc32(d[: 1 << 20])


def test_config4_zlib_many_streams_adler(ctx, titles256):
    streams = [titles256[i << 20:(i + 1) << 20] for i in range(256)]         # 256 x 1 MiB (BASELINE: 1024 across 8 GPUs = 128 per GPU)
    encs = ctx.encode_batch(1, streams)
    for i, (s, e) in enumerate(zip(streams, encs)):
        assert e[:2] == b"\x78\x9c" and e[-4:] == pyzlib.adler32(s).to_bytes(4, "big"), i
    assert encs[5] == orc.encode(1, streams[5].tobytes())
    assert pyzlib.decompress(encs[200]) == streams[200].tobytes()
    res = ctx.decode_batch(1, encs, caps=[(1 << 20) + 64] * 256)
    assert all(st == 0 and out == s.tobytes() and used == len(e) for (st, out, used, _), s, e in zip(res, streams, encs))
    assert ctx.adler32(streams[:32]) == [pyzlib.adler32(s) for s in streams[:32]]


def test_config5_many_block_gzip_decode(ctx, titles256):
    """decode-only shape: ONE gzip member with 1024 dynamic blocks (1 GiB; BASELINE uses 8 GiB = 8192 blocks)"""
    from libflate_b200 import native
    d = np.tile(titles256, 4)
    n = d.size
    sched = np.asarray([8192] * (n // 8192), dtype=np.int64)
    enc = np.empty(native.lib().b2f_encode_bound(n, len(sched), None), dtype=np.uint8)
    m = ctx.encode_into(native.FMT_GZIP, d, enc, sched, mtime=0)
    assert int.from_bytes(bytes(enc[m - 4:m]), "little") == n % (1 << 32)
    dec = np.empty(n + 64, dtype=np.uint8)
    before = ctx.stats()
    dl, used, st = ctx.decode_into(native.FMT_GZIP, enc, m, dec)
    after = ctx.stats()
    assert st == 0 and dl == n and used == m
    assert np.array_equal(dec[:n], d)
    assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 1     # block-parallel path, not the in-order kernel
    assert dict(after["stages"]).get("lz_resolve", 0) > 0

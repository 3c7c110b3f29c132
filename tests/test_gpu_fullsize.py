"""BASELINE.json configurations at full size: every compressed byte of configs 2 and 3 against the oracle, config 4 with
32 streams against the oracle and every Adler-32 trailer against zlib, config 5 at 4 GiB (8 GiB needs ~100 GB of encode
scratch in one call) as a round trip; plus independent inflate by Python's zlib and checksums."""
import json
import os
import zlib as pyzlib
from concurrent.futures import ThreadPoolExecutor

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
    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:              # all 64 streams bit-exact against the oracle (ctypes releases the GIL)
        want = list(ex.map(lambda a: orc.encode(0, a.tobytes()), streams))
    for i in range(64):
        assert encs[i] == want[i], i
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
    # every one of the 80 029 560 compressed bytes (265 blocks) against the oracle, and the oracle against the committed golden
    # that bench.py checks its own output with (tests/golden/config_goldens.json, made by make_config_goldens.py)
    want = orc.encode(orc.FMT_GZIP, d.tobytes(), sched.tolist(), mtime=0)
    assert m == len(want)
    bad = np.nonzero(enc[:m] != np.frombuffer(want, dtype=np.uint8))[0]
    assert bad.size == 0, int(bad[0])
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_goldens.json")))["config3"]
    assert (gold["size"], gold["enc_len"], gold["enc_crc32"], gold["plain_crc32"]) == (n, m, pyzlib.crc32(want), pyzlib.crc32(d))
    dec = np.empty(n + 64, dtype=np.uint8)
    before = ctx.stats()
    dl, used, st = ctx.decode_into(native.FMT_GZIP, enc, m, dec)
    after = ctx.stats()
    assert st == 0 and dl == n and used == m and np.array_equal(dec[:n], d)
    assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 1     # the 265-block stream took the parallel path
    assert pyzlib.crc32(pyzlib.decompress(bytes(enc[:m]), 31)[: 1 << 20]) == pyzlib.crc32(d[: 1 << 20])


def test_config4_zlib_many_streams_adler(ctx):
    """the BASELINE shape on ONE GPU: 1024 zlib streams of 1 MiB, one write_all each (at 8 GPUs each rank takes 128 of them)"""
    from libflate_b200 import titles
    big = titles.generate(1 << 30, seed=2000, cache_dir=CACHE)
    streams = [big[i << 20:(i + 1) << 20] for i in range(1024)]
    encs = ctx.encode_batch(1, streams)
    for i, (s, e) in enumerate(zip(streams, encs)):
        assert e[:2] == b"\x78\x9c" and e[-4:] == pyzlib.adler32(s).to_bytes(4, "big"), i     # every Adler-32 trailer
    pick = list(range(0, 1024, 32))                                           # 32 streams byte for byte against the oracle
    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        want = list(ex.map(lambda i: orc.encode(1, streams[i].tobytes()), pick))
    for i, w in zip(pick, want):
        assert encs[i] == w, i
    assert pyzlib.decompress(encs[200]) == streams[200].tobytes()
    res = ctx.decode_batch(1, encs, caps=[(1 << 20) + 64] * 1024)
    assert all(st == 0 and out == s.tobytes() and used == len(e) for (st, out, used, _), s, e in zip(res, streams, encs))
    assert ctx.adler32(streams[:64]) == [pyzlib.adler32(s) for s in streams[:64]]


def test_config5_many_block_gzip_decode(ctx, titles256):
    """decode-only shape: ONE gzip member with 4096 dynamic blocks (4 GiB; BASELINE uses 8 GiB = 8192 blocks, which needs
    ~100 GB of encode scratch in a single call -- bench.py --workload config5 runs the 8 GiB case)"""
    from libflate_b200 import native
    d = np.tile(titles256, 16)
    n = d.size
    sched = np.asarray([8192] * (n // 8192), dtype=np.int64)
    enc = np.empty(native.lib().b2f_encode_bound(n, len(sched), None), dtype=np.uint8)
    m = ctx.encode_into(native.FMT_GZIP, d, enc, sched, mtime=0)
    assert int.from_bytes(bytes(enc[m - 4:m]), "little") == n % (1 << 32)
    crc1 = pyzlib.crc32(titles256)                                            # CRC of the 16 repetitions from the generator's 256 MiB
    crc = 0
    for _ in range(16):
        crc = pyzlib.crc32(titles256, crc)
    assert int.from_bytes(bytes(enc[m - 8:m - 4]), "little") == crc and crc1 != crc
    dec = np.empty(n + 64, dtype=np.uint8)
    before = ctx.stats()
    dl, used, st = ctx.decode_into(native.FMT_GZIP, enc, m, dec)
    after = ctx.stats()
    assert st == 0 and dl == n and used == m
    assert np.array_equal(dec[:n], d)
    assert after["decode_parallel_streams"] - before["decode_parallel_streams"] == 1     # block-parallel path, not the in-order kernel
    assert dict(after["stages"]).get("lz_resolve", 0) > 0

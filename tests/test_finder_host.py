"""CPU checks of the block-boundary finder's header test (libflate_b200/csrc/finder_core.cuh): every dynamic block the
oracle (libflate restatement) or zlib writes must be accepted, and false positives on real / random data must be rare."""
import random
import zlib as pyzlib

import hostcheck as hc
from oracle import oracle as orc


def _text(rng, n):
    words = [bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz_") for _ in range(rng.randint(2, 9))) for _ in range(300)]
    b = bytearray()
    while len(b) < n:
        b += rng.choice(words) + b"\n"
    return bytes(b[:n])


def test_all_true_dynamic_blocks_are_candidates():
    rng = random.Random(5)
    d = _text(rng, 400000)
    streams = [orc.encode(orc.FMT_DEFLATE, d, block_size=30000), orc.encode(orc.FMT_DEFLATE, b""), orc.encode(orc.FMT_DEFLATE, b"a"),
               orc.encode(orc.FMT_DEFLATE, bytes(rng.getrandbits(8) for _ in range(100000)), block_size=20000),
               orc.encode(orc.FMT_DEFLATE, b"\x00" * 300000, block_size=50000)]
    for lvl in (6, 9):
        co = pyzlib.compressobj(lvl, pyzlib.DEFLATED, -15)
        streams.append(co.compress(d) + co.flush())
    total_fp = 0
    for s in streams:
        starts = hc.block_starts(s)
        n, cands = hc.find_candidates(s)
        dyn = [q for q in starts if ((s[q >> 3] | (s[(q >> 3) + 1] << 8 if (q >> 3) + 1 < len(s) else 0)) >> (q & 7) >> 1) & 3 == 2]
        assert set(dyn) <= set(cands), (len(starts), len(cands))
        total_fp += len(set(cands) - set(starts))
    assert total_fp <= 3, total_fp


def test_false_positive_rate_on_random_bytes():
    rng = random.Random(6)
    blob = bytes(rng.getrandbits(8) for _ in range(300000))
    n, _ = hc.find_candidates(blob)
    assert n <= 2, n


def test_fast_header_tests_equal_their_written_out_forms():
    """table-driven Kraft test and 32-offsets-at-once precheck (what k_find_blocks runs) against the field-by-field forms"""
    import ctypes as C
    L = hc.lib()
    L.hc_finder_selftest.restype = C.c_uint32
    for seed in (1, 2, 3):
        assert L.hc_finder_selftest(C.c_uint64(seed), C.c_uint32(400000)) == 0

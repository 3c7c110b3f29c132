"""Host arithmetic of the single-stream split, no GPU: checksum combination against zlib, bit placement against
big-integer arithmetic, block cuts against the library's own segmentation plan (b2f_plan_from_writes)."""
import random
import zlib

import numpy as np

from libflate_b200 import native as nv, split


def test_checksum_combine_matches_zlib():
    rng = random.Random(1)
    L = nv.lib()
    for _ in range(200):
        a = bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 7, 100, 5000])))
        b = bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 3, 777, 70000])))
        assert L.b2f_crc32_combine(zlib.crc32(a), zlib.crc32(b), len(b)) == zlib.crc32(a + b)
        assert L.b2f_adler32_combine(zlib.adler32(a), zlib.adler32(b), len(b)) == zlib.adler32(a + b)
    # chains of many parts, long second operands (exponent well above 2^32 bits)
    crc, adler, blob = 0, 1, b""
    for k in range(6):
        part = bytes((k * 37 + i) & 255 for i in range(100000 + k))
        crc, adler, blob = L.b2f_crc32_combine(crc, zlib.crc32(part), len(part)), L.b2f_adler32_combine(adler, zlib.adler32(part), len(part)), blob + part
    assert (crc, adler) == (zlib.crc32(blob), zlib.adler32(blob))
    assert L.b2f_crc32_combine(0x12345678, 0, 1 << 33) == L.b2f_crc32_combine(L.b2f_crc32_combine(0x12345678, 0, 1 << 32), 0, 1 << 32)


def test_bit_placement_matches_bigint():
    rng = random.Random(2)
    for _ in range(50):
        parts = [(rng.getrandbits(n), n) for n in (rng.randint(1, 300) for _ in range(rng.randint(1, 6)))]
        start = rng.randint(0, 40)
        want, pos = 0, start
        out = np.zeros((start + sum(n for _, n in parts) + 7) // 8 + 4, dtype=np.uint8)
        for v, n in parts:
            want |= v << pos
            sh = pos & 7                                              # what b2f_bits_shift_device does on the part's GPU
            data = np.frombuffer((v << sh).to_bytes((sh + n + 7) // 8 + 1, "little"), dtype=np.uint8)
            split.or_bits(out, pos, data, n)
            pos += n
        assert int.from_bytes(out.tobytes(), "little") == want


def test_block_cuts_match_the_library_plan():
    rng = random.Random(3)
    for _ in range(30):
        n = rng.randint(1, 6 << 20)
        sched, left = [], n
        while left > 0:
            w = min(left, rng.choice([8192, 70001, 1 << 20, 300000, 1]))
            sched.append(w); left -= w
            if rng.random() < 0.05:
                sched.append(-1)
        bs = rng.choice([1 << 20, 100000, 65536])
        _, block_ends, _, _ = nv.plan_from_writes(sched, n, block_size=bs)
        cuts = split.block_cuts(sched, n, bs)
        assert [c[0] for c in cuts] == block_ends[:-1]              # the plan's last entry is finish()'s block
        for nparts in (2, 3, 8):
            pl = split.plan_parts(sched, n, nparts, bs)
            assert pl[0][0] == 0 and pl[-1][1] == n and pl[-1][3] == len(sched)
            assert all(a[1] == b[0] and a[3] == b[2] for a, b in zip(pl, pl[1:]))
            assert all(p[0] in (0, *[c[0] for c in cuts]) for p in pl)

"""ONE stream encoded as several parts by several contexts (on the test box all contexts share cuda:0, in bench.py
--workload config3-split every rank owns one): byte-identical to the single-call encode and to the oracle."""
import random

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def test_split_encode_equals_single_call_and_oracle():
    from libflate_b200 import native, split, titles
    ctxs = [native.Context(0) for _ in range(4)]
    try:
        d = titles.generate(9 * (1 << 20) + 12345, seed=21)
        scheds = {"8k": [8192] * (d.size // 8192 + 1), "odd": [700001] * (d.size // 700001 + 1), "flush": [3 << 20, -1, 1 << 20, -1, -1] + [65536] * 100}
        for name, sched in scheds.items():
            for fmt in (0, 1, 2):
                want = orc.encode(fmt, d.tobytes(), sched, mtime=5)
                for nparts in (2, 3, 4):
                    got = split.encode_split(ctxs[:nparts], fmt, d, sched, mtime=5)
                    assert got == want, (name, fmt, nparts, len(got), len(want))
        got = split.encode_split(ctxs[:3], 1, d, scheds["8k"], block_size=300000)
        assert got == orc.encode(1, d.tobytes(), scheds["8k"], block_size=300000)
    finally:
        for c in ctxs:
            c.close()

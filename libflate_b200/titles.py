"""Synthetic "enwiki-titles-shaped" text (SURVEY.md section 8d / BASELINE.md section 3): Zipf(1.05) over a 50 000 word
vocabulary, 1-5 words joined by '_', occasional `_(word)` / `_YYYY` / `N_` decorations, generated as independently
sorted 2 MiB segments (seed = base + segment index), newline terminated.  Deterministic for a given (seed, size)."""
import math
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np

SEGMENT = 2 << 20
_VOCAB = None


def _vocab():
    global _VOCAB
    if _VOCAB is None:
        rng = np.random.default_rng(20240925)
        n = 50000
        lens = np.clip(np.rint(rng.normal(6.5, 2.5, n)).astype(int), 2, 16)
        cap = rng.random(n) < 0.75
        letters = rng.integers(0, 26, size=(n, 16))
        words = []
        for i in range(n):
            w = bytes((97 + letters[i, : lens[i]]).astype(np.uint8))
            if cap[i]:
                w = w[:1].upper() + w[1:]
            words.append(w)
        p = 1.0 / np.arange(1, n + 1) ** 1.05
        cdf = np.cumsum(p / p.sum())
        _VOCAB = (words, cdf)
    return _VOCAB


def segment(seed, size=SEGMENT):
    words, cdf = _vocab()
    rng = np.random.default_rng(seed)
    est = int(size / 11.5) + 64
    titles = []
    total = 0
    while total < size:
        k = 1 + np.minimum(4, np.floor(rng.exponential(1 / 0.9, est)).astype(int))
        ids = np.searchsorted(cdf, rng.random(int(k.sum()) + 2 * est))
        deco = rng.random(est)
        years = rng.integers(1000, 2026, est)
        nums = rng.integers(1, 100, est)
        j = 0
        for t in range(est):
            kk = k[t]
            s = b"_".join(words[i] for i in ids[j: j + kk])
            j += kk
            d = deco[t]
            if d < 0.08:
                s += b"_(" + words[ids[j]] + b")"
                j += 1
            elif d < 0.12:
                s += b"_%d" % years[t]
            elif d < 0.14:
                s = b"%d_" % nums[t] + s
            titles.append(s)
            total += len(s) + 1
            if total >= size:
                break
    titles.sort()
    out = b"\n".join(titles) + b"\n"
    return out[:size]


def _seg_job(args):
    return segment(*args)


def generate(size, seed=42, workers=None, cache_dir=None):
    """`size` bytes of titles-shaped text as a numpy uint8 array."""
    if cache_dir:
        path = os.path.join(cache_dir, f"titles_{seed}_{size}.bin")
        if os.path.exists(path) and os.path.getsize(path) == size:
            return np.fromfile(path, dtype=np.uint8)
    nseg = max(1, math.ceil(size / SEGMENT))
    jobs = [(seed + i, min(SEGMENT, size - i * SEGMENT)) for i in range(nseg)]
    if workers is None:
        workers = min(len(jobs), os.cpu_count() or 1)
    if workers > 1 and nseg > 1:
        with ProcessPoolExecutor(workers) as ex:
            parts = list(ex.map(_seg_job, jobs, chunksize=1))
    else:
        parts = [_seg_job(j) for j in jobs]
    arr = np.frombuffer(b"".join(parts), dtype=np.uint8)
    assert arr.size == size
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        arr.tofile(path)
    return arr

"""CPU model of the skip links of k_lz_find (csrc/encode_kernels.cu, DESIGN.md "Skip links").

Claim: if link[j] of a chain node j is replaced by ANY link that only skips nodes whose trigram equals j's own trigram, every walk
still finds the most recent earlier occurrence of its trigram -- so the kernel may upgrade links without any ordering between
warps (a walker sees the plain link or an upgraded one).  The model applies upgrades in random order, partially, repeatedly, and
checks every position's walk against brute force (libflate's prefix table: `libflate_lz77/src/default.rs:69-129`)."""
import random

import numpy as np

WINDOW = 300          # small window and few buckets: long chains, many collisions, links that fall out of the window
BUCKETS = 16


def _trigrams(d):
    return d[:-2].astype(np.uint32) | (d[1:-1].astype(np.uint32) << 8) | (d[2:].astype(np.uint32) << 16)


def _plain_links(tri):
    head = {}
    link = np.zeros(len(tri), dtype=np.int64)
    for i, t in enumerate(tri):
        b = int(t * 0x9E3779B1 & 0xFFFFFFFF) >> 28 & (BUCKETS - 1)
        p = head.get(b)
        if p is not None and i - p <= WINDOW:
            link[i] = i - p
        head[b] = i
    return link


def _brute(tri):
    last = {}
    want = np.zeros(len(tri), dtype=np.int64)
    for i, t in enumerate(tri):
        p = last.get(int(t))
        if p is not None and i - p <= WINDOW:
            want[i] = i - p
        last[int(t)] = i
    return want


def _walk(i, tri, first, link):
    """the kernel's walk: first hop = the plain link of i (the atomic's return value), later hops = whatever link[] holds"""
    d, total, j = int(first[i]), 0, i
    while d:
        total += d
        if total > WINDOW:
            return 0
        j -= d
        if tri[j] == tri[i]:
            return total
        d = int(link[j])
    return 0


def _upgrade(i, tri, plain, link):
    """what a warp does for position i at some point in time: candidate c = i - plain[i]; same trigram -> link past it"""
    d = int(plain[i])
    if d and tri[i - d] == tri[i]:
        dn = int(link[i - d])                       # the candidate's link as it is right now (plain or already upgraded)
        up = d + dn
        link[i] = up if dn and up <= WINDOW else 0


def test_unordered_skip_link_upgrades_keep_every_walk_exact():
    rng = random.Random(5)
    for case in range(6):
        n = 3000
        if case % 3 == 0:
            d = np.array([rng.choice(b"ab") for _ in range(n)], dtype=np.uint8)                       # runs and tiny alphabet
        elif case % 3 == 1:
            words = [bytes(rng.choice(b"abcdefgh") for _ in range(rng.randint(2, 6))) for _ in range(12)]
            d = np.frombuffer(b" ".join(rng.choice(words) for _ in range(n))[:n], dtype=np.uint8).copy()
        else:
            d = np.array([rng.randrange(256) if rng.random() < 0.2 else 0x41 for _ in range(n)], dtype=np.uint8)   # one hot trigram
        tri = _trigrams(d)
        plain = _plain_links(tri)
        want = _brute(tri)
        link = plain.copy()
        order = list(range(len(tri)))
        for rnd in range(4):                        # unordered, partial, repeated: every intermediate state must be exact
            rng.shuffle(order)
            for i in order[: len(order) * (rnd + 1) // 4]:
                _upgrade(i, tri, plain, link)
            got = np.array([_walk(i, tri, plain, link) for i in range(len(tri))])
            assert np.array_equal(got, want), (case, rnd, int(np.flatnonzero(got != want)[0]))
        # upgrades done in position order collapse whole runs: no walk needs more hops than there are DISTINCT other trigrams
        link = plain.copy()
        for i in range(len(tri)):
            _upgrade(i, tri, plain, link)
        assert np.array_equal(np.array([_walk(i, tri, plain, link) for i in range(len(tri))]), want)

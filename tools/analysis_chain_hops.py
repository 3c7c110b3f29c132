"""Why k_lz_find looks the way it does: hash-chain hop statistics of titles-shaped text for 13..16 hash bits.
For every position: hops along the bucket chain until the same trigram (or the end of the 32 KiB window); per warp step (32
consecutive positions) the maximum is what a SIMT walk pays.  Printed on the first 256 KiB of the bench input:
14 bits -> 1.24 hops per position on average but 9.6 per warp step, p99 116 (runs of identical lines flood a bucket)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libflate_b200 import titles

d = titles.generate(2 << 20, seed=42)[:262144].astype(np.uint32)
tri = d[:-2] | (d[1:-1] << 8) | (d[2:] << 16)
W = 32768
for bits in (13, 14, 15, 16):
    h = ((tri.astype(np.uint64) * 0x9E3779B1) & 0xFFFFFFFF) >> (32 - bits)
    head = {}
    link = np.full(len(tri), -1, dtype=np.int64)
    for i in range(len(tri)):
        hh = int(h[i]); j = head.get(hh, -1)
        if j >= 0 and i - j <= W: link[i] = j
        head[hh] = i
    hops = np.zeros(len(tri), dtype=np.int32); found = np.zeros(len(tri), dtype=bool)
    for i in range(len(tri)):
        j = link[i]; c = 0
        while j >= 0 and i - j <= W:
            c += 1
            if tri[j] == tri[i]: found[i] = True; break
            j = link[j]
        hops[i] = c
    m = hops[: len(hops) // 32 * 32].reshape(-1, 32).max(axis=1)
    print(f"{bits} bits: mean hops {hops.mean():.2f}  found {found.mean():.3f}  mean warp-step max {m.mean():.2f}  p99 {np.percentile(m, 99):.0f}  "
          f"positions with more than 16 / 176 hops: {(hops > 16).mean() * 100:.3f} % / {(hops > 176).mean() * 100:.4f} %")

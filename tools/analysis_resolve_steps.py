"""Dependency structure of the decode's LZ77 resolution (k_spec_resolve) on titles-shaped text, per step of 32 tokens:
how many matches read only data from before the step ('free'), how many can be redirected there through matches of the same
step (out[p] == out[p - dist_j] for every byte p of match j's output: a source range that lies inside ONE earlier token of the
step can be shifted by that token's distance), and how many really depend on the step's own output through several tokens."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libflate_b200 import titles
from oracle import oracle as orc

d = titles.generate(2 << 20, seed=42)[:262144]
codes = np.asarray(orc.lz77_default(d.tobytes()), dtype=np.uint64)
is_m = (codes & 0x80000000) != 0
ln = np.where(is_m, (codes >> 16) & 0x1FF, 1).astype(np.int64)
dist = np.where(is_m, codes & 0xFFFF, 0).astype(np.int64)
n = len(codes)
tot = {"matches": 0, "free": 0, "redirected": 0, "unresolved": 0, "steps": 0, "depth_sum": 0, "max_inorder_after": 0}
inorder_before = []; inorder_after = []
for s0 in range(0, n - 31, 32):
    L = ln[s0:s0 + 32]; D = dist[s0:s0 + 32]; M = is_m[s0:s0 + 32]
    off = np.concatenate([[0], np.cumsum(L)[:-1]])
    nb = na = 0
    for l in range(32):
        if not M[l]: continue
        tot["matches"] += 1
        s = off[l] - D[l]
        if s + L[l] <= 0: tot["free"] += 1; continue
        nb += 1
        depth = 0; res = None
        while True:
            if s + L[l] <= 0: res = "redirected"; break
            if s < 0: res = "unresolved"; break                      # straddles the step start
            j = int(np.searchsorted(off, s, side="right") - 1)
            if not M[j] or s + L[l] > off[j] + L[j] or j >= l: res = "unresolved"; break
            s -= D[j]; depth += 1
            if depth > 16: res = "unresolved"; break
        tot[res] += 1; tot["depth_sum"] += depth
        if res == "unresolved": na += 1
    inorder_before.append(nb); inorder_after.append(na); tot["steps"] += 1
print(tot)
print("in-order matches per step: before", np.mean(inorder_before), "after redirect", np.mean(inorder_after))

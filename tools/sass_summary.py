"""Writes a per-kernel SASS mnemonic histogram of libb2f.so (cuobjdump -sass): which kernels exist in the shipped sm_100a
cubins, how large they are and which hardware features they use (UBLKCP = TMA bulk copy, SYNCS = mbarrier, ATOMS/RED = atomics,
MATCH/REDUX/VOTE = warp collectives, IDP = dp4a).  usage: python tools/sass_summary.py > profiles/r02_sass_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "libflate_b200", "libb2f.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, arch = None, None
hist = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0]] += 1
print(f"# {os.path.relpath(so, ROOT)}: arch {arch}; mnemonic counts per kernel (static SASS instructions)")
feat = ["UBLKCP", "SYNCS", "ATOMS", "ATOMG", "RED", "MATCH", "REDUX", "VOTE", "SHFL", "IDP", "LDGSTS", "BAR", "LDS", "STS", "LDG", "STG"]
print(f"{'kernel':58s} {'insts':>6s} " + " ".join(f"{f:>6s}" for f in feat))
for k, h in hist.items():
    print(f"{k[:58]:58s} {sum(h.values()):6d} " + " ".join(f"{h.get(f, 0):6d}" for f in feat))
print()
for k, h in hist.items():
    print(f"{k}: " + ", ".join(f"{m} {c}" for m, c in h.most_common(12)))

"""Encode-only stage times (no output check): for timing experiments with variant builds (B2F_LIB=...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libflate_b200 import native, titles
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 265
ctx = native.Context(0)
d = titles.generate(mib << 20, seed=42)
sched = [8192] * (d.size // 8192 + 1)
ctx.set_overlap(False)
for it in range(3):
    enc = ctx.encode(native.FMT_GZIP, d, sched, mtime=0)
    se = ctx.stats()
print(os.environ.get("B2F_LIB", "default"), " ".join(f"{n}={ms:.3f}" for n, ms in se["stages"] if n in ("lz_find", "lz_fixup", "parse_exits")), len(enc))

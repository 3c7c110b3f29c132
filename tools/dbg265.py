import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B2F_DEBUG"] = "1"
from libflate_b200 import native, titles
ctx = native.Context(0)
d = titles.generate(265 << 20, seed=42)
enc = ctx.encode(native.FMT_DEFLATE, d, [8192] * (d.size // 8192 + 1))
st, out, used, _ = ctx.decode(native.FMT_DEFLATE, enc, cap=d.size + 64)
print("status", st, out == d.tobytes(), ctx.stats())

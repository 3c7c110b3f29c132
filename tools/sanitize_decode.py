"""Small decode workload for `compute-sanitizer` (memcheck / racecheck / initcheck are 10-50x slower than a plain run):
libflate-style and foreign streams through the speculative path, compared with the plain bytes."""
import os
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libflate_b200 import native, titles

ctx = native.Context(0)
text = titles.generate(3 << 20, seed=9).tobytes()
own = ctx.encode(native.FMT_ZLIB, text, [8192] * (len(text) // 8192 + 1))
for name, enc in (("own", own), ("zlib6", zlib.compress(text, 6)), ("zeros", zlib.compress(b"\0" * (3 << 20) + text[:300000], 6))):
    plain = text if name != "zeros" else b"\0" * (3 << 20) + text[:300000]
    st, out, used, _ = ctx.decode(native.FMT_ZLIB, enc, cap=len(plain) + 64)
    assert st == 0 and out == plain and used == len(enc), (name, st)
    print(name, "ok", ctx.stats()["decode_parallel_streams"], ctx.stats()["decode_inorder_streams"])
assert ctx.crc32([text])[0] == zlib.crc32(text) and ctx.adler32([text])[0] == zlib.adler32(text)
ctx.close()
print("sanitize workload ok")

"""Prints per-stage device times of one encode and one decode (through the C ABI, host buffers) for a titles-shaped input."""
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libflate_b200 import native, titles

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "A"
ctx = native.Context(0)
d = titles.generate(mib << 20, seed=42)
sched = [8192] * (d.size // 8192 + 1) if mode in ("A", "P") else None
if mode == "P":                       # profiling pass (ncu): every kernel once, serialised, whole-input launches
    ctx.set_overlap(False)
for it in range(1 if mode == "P" else 3):
    t = time.time(); enc = ctx.encode(native.FMT_GZIP, d, sched, mtime=0); te = time.time() - t
    se = ctx.stats()
    t = time.time(); st, out, used, _ = ctx.decode(native.FMT_GZIP, enc, cap=d.size + 64); td = time.time() - t
    sd = ctx.stats()
assert st == 0 and out == d.tobytes()
print(f"input {mib} MiB  ratio {len(enc)/d.size:.4f}  encode wall {te*1e3:.1f} ms  decode wall {td*1e3:.1f} ms")
ctx.set_overlap(False)
enc2 = ctx.encode(native.FMT_GZIP, d, sched, mtime=0); se2 = ctx.stats()
assert enc2 == enc
print("encode (overlap) :", " ".join(f"{n}={ms:.3f}" for n, ms in se["stages"]), f"| device {se['device_ms']:.3f}")
se = se2
print("encode stages (ms):", " ".join(f"{n}={ms:.3f}" for n, ms in se["stages"]), f"| device {se['device_ms']:.3f}")
print("decode stages (ms):", " ".join(f"{n}={ms:.3f}" for n, ms in sd["stages"]), f"| device {sd['device_ms']:.3f}")

"""Wall-clock split of the end-to-end round trip (C ABI, host buffers): encode and decode separately, page-locked vs pageable
buffers, with the library's per-stage device times of the overlapped run.  usage: e2e_times.py [MiB]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libflate_b200 import native, titles

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 265
n = 277_303_937 if mib == 265 else mib << 20
ctx = native.Context(0)
d = titles.generate(n, seed=42, cache_dir="/tmp/b2f_bench_cache")
sched = np.full(n // 8192 + 1, 8192, dtype=np.int64)
cap = native.lib().b2f_encode_bound(n, len(sched), None)
for kind in ("pinned", "pageable"):
    if kind == "pinned":
        h_in, h_enc, h_dec = native.host_alloc(n), native.host_alloc(cap), native.host_alloc(n + 64)
        h_in[:] = d
    else:
        h_in, h_enc, h_dec = d, np.empty(cap, dtype=np.uint8), np.empty(n + 64, dtype=np.uint8)
        h_enc[:] = 0; h_dec[:] = 0                      # touch the pages once (a first-touch page fault is not part of the path)
    te, td = [], []
    for it in range(6):
        t = time.perf_counter(); m = ctx.encode_into(native.FMT_GZIP, h_in, h_enc, sched, mtime=0); te.append(time.perf_counter() - t)
        se = ctx.stats()
        t = time.perf_counter(); dl, used, st = ctx.decode_into(native.FMT_GZIP, h_enc, m, h_dec); td.append(time.perf_counter() - t)
        sd = ctx.stats()
        assert st == 0 and dl == n
    assert np.array_equal(h_dec[:n], d)
    e, dd = min(te[2:]) * 1e3, min(td[2:]) * 1e3
    print(f"{kind:9s} encode {e:7.2f} ms  decode {dd:7.2f} ms  round trip {e + dd:7.2f} ms = {n / (e + dd) * 1e3 / 2**30:6.2f} GiB/s")
    print("   encode stages:", " ".join(f"{a}={b:.2f}" for a, b in se["stages"]), f"| device {se['device_ms']:.2f}")
    print("   decode stages:", " ".join(f"{a}={b:.2f}" for a, b in sd["stages"]), f"| device {sd['device_ms']:.2f}")

# decode pipeline depth (B2F_DECODE_PARTS is read when the context is created)
h_enc, h_dec = native.host_alloc(cap), native.host_alloc(n + 64)
m = ctx.encode_into(native.FMT_GZIP, d, h_enc, sched, mtime=0)
ctx.close()
for parts in (1, 2, 3, 4, 6, 8):
    os.environ["B2F_DECODE_PARTS"] = str(parts)
    c2 = native.Context(0)
    td = []
    for it in range(6):
        t = time.perf_counter(); dl, used, st = c2.decode_into(native.FMT_GZIP, h_enc, m, h_dec); td.append(time.perf_counter() - t)
    sd = c2.stats()
    print(f"decode parts={parts}: {min(td[2:]) * 1e3:6.2f} ms wall |", " ".join(f"{a}={b:.2f}" for a, b in sd["stages"]))
    c2.close()

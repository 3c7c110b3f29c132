"""Resident (device buffers) encode and decode times of the headline workload, separately: wall clock around each call (the calls
synchronise) and the library's own device time; overlap on / off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libflate_b200 import native, titles
size = 277303937
ctx = native.Context(0)
L = native.lib()
data = titles.generate(size, seed=42)
sched = np.asarray([8192] * (size // 8192) + ([size % 8192] if size % 8192 else []), dtype=np.int64)   # an array: no per-call list conversion in the binding
bound = L.b2f_encode_bound(size, len(sched), None)
d_in = torch.from_numpy(data).cuda()
d_enc = torch.empty(bound + 256, dtype=torch.uint8, device="cuda")
d_dec = torch.empty(size + 256, dtype=torch.uint8, device="cuda")
for overlap in (True, False):
    ctx.set_overlap(overlap)
    te, td, de, dd = [], [], [], []
    for it in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ol, st = ctx.encode_device(native.FMT_GZIP, d_in.data_ptr(), [0], [size], d_enc.data_ptr(), [0], [bound], [sched], mtime=0)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        se = ctx.stats()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        dl, used, st = ctx.decode_device(native.FMT_GZIP, d_enc.data_ptr(), [0], [ol[0]], d_dec.data_ptr(), [0], [size + 64])
        torch.cuda.synchronize(); t3 = time.perf_counter()
        sd = ctx.stats()
        if it >= 2:
            te.append(t1 - t0); td.append(t3 - t2); de.append(se["device_ms"]); dd.append(sd["device_ms"])
    print(f"overlap={overlap}: encode wall {1e3*np.mean(te):.2f} ms (device {np.mean(de):.2f})   decode wall {1e3*np.mean(td):.2f} ms (device {np.mean(dd):.2f})")
    print("   enc:", " ".join(f"{n}={ms:.2f}" for n, ms in se["stages"]))
    print("   dec:", " ".join(f"{n}={ms:.2f}" for n, ms in sd["stages"]))

"""Aggregates the warp-stall samples of one kernel in an .ncu-rep by CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex library.so object_name   (object_name e.g. spec_kernels)
The SASS page of the report gives samples per instruction (in order); `nvdisasm -g` of the same build gives the source line of
every instruction (in the same order)."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, kern, so, obj = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Address"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
samples = [int(r[ci["# Samples"]] or 0) for r in rows[1:] if len(r) > ci["# Samples"] and r[ci["# Samples"]].isdigit()]
execs = [int(r[ci["Instructions Executed"]] or 0) for r in rows[1:] if len(r) > ci["# Samples"] and r[ci["# Samples"]].isdigit()]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = next(f for f in os.listdir(tmp) if f.startswith(obj + "."))
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's section
sec = next(i for i, l in enumerate(dis) if l.startswith("//---") and re.search(kern, l))
cur, per_inst = None, []
for l in dis[sec + 1:]:
    if l.startswith("//---"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per_inst.append(cur)
assert len(per_inst) == len(samples), (len(per_inst), len(samples))
agg = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter()
for ln, s, e in zip(per_inst, samples, execs):
    agg[ln] += s; cnt[ln] += 1; ex[ln] += e
tot = sum(samples)
src_cache = {}
def src(fn, n):
    if fn not in src_cache:
        for root in ("libflate_b200/csrc", "."):
            p = os.path.join(root, fn)
            if os.path.exists(p):
                src_cache[fn] = open(p).read().splitlines(); break
        else:
            src_cache[fn] = []
    L = src_cache[fn]
    return L[n - 1].strip()[:110] if 0 < n <= len(L) else ""
print("total samples", tot, " instructions", len(samples), " warp-instructions executed", sum(execs))
for ln, s in agg.most_common(top):
    print(f"{s:7d} {100*s/tot:5.1f}%  n_inst={cnt[ln]:3d} exec={ex[ln]:>10}  {ln[0]}:{ln[1]:<4d} {src(*ln)}")

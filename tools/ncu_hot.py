"""Prints the hottest SASS/source lines (by warp-stall samples) of one kernel from an .ncu-rep (source page CSV)."""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first block only (SASS view): header at line index 1
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Address"')), len(lines))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:end]))))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in rows[1:] if len(r) > ci["# Samples"] and r[ci["# Samples"]].isdigit())
data = []
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[1:]:
    if len(r) <= ci["# Samples"] or not r[ci["# Samples"]].isdigit():
        continue
    s = int(r[ci["# Samples"]])
    top_st = sorted(((int(r[ci[h]] or 0), h) for h in stalls), reverse=True)[:2]
    data.append((s, r[ci["Source"]], r[ci["Instructions Executed"]], top_st))
print("total samples", tot)
for s, src, ie, st in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{s:7d} {100*s/max(tot,1):5.1f}%  exec={ie:>10}  {src[:90]:90s} {st}")

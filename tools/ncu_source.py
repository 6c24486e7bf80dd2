#!/usr/bin/env python
"""Per-instruction stall summary of one kernel from an .ncu-rep (source page): cumulative sample curve,
stall-reason mix and the hottest SASS lines.   usage: ncu_source.py <rep> <kernel regex> [top N]"""
import csv, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ix = {n: hdr.index(n) for n in names}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[0] == "Address":
        break
    try:
        data.append((int(r[isamp]), int(r[iex]), r[ia].strip(), {n: int(r[ix[n]]) for n in names}))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("kernel", rows[hi - 1][1][:80] if hi else "")
print("SASS lines", len(data), "samples", tot, "warp-instructions per warp %.0f" % (sum(d[1] for d in data) / max(data[0][1], 1)))
agg = {n: sum(d[3][n] for d in data) for n in names}
print("stall mix:", {k: round(v / tot, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v / tot > 0.005})
c = 0
for i, d in enumerate(data):
    c += d[0]
    if i % 100 == 0:
        print("  pos %4d cum %.3f  %s" % (i, c / tot, d[2][:60]))
for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:topn]:
    print("  hot %4d samples %5d  %-60s %s" % (i, data[i][0], data[i][2][:60], max(data[i][3].items(), key=lambda kv: kv[1])))

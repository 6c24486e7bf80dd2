#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines
(ncu source page gives SASS in program order; nvdisasm -g of the same cubin gives the line of every SASS instruction).
usage: ncu_lines.py <rep> <kernel regex> <mangled-name substring> [top N] [lib.so]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, pat, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lib = sys.argv[5] if len(sys.argv) > 5 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "h-slam_b200", "libhslam_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = None
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if mangled in t:
            dis = t.splitlines()
            break
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and mangled in l][0]
lines, cur, curf = [], None, None
for l in dis[start + 1:]:
    if l.startswith(".text.") and lines:
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        curf, cur = os.path.basename(m.group(1)), int(m.group(2))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((curf, cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if len(r) >= len(hdr):
        try:
            data.append((int(r[iex]), int(r[isamp])))
        except ValueError:
            pass
data = data[:len(lines)]
warps = max(data[0][0], 1)
by = collections.defaultdict(lambda: [0, 0])
for (f, ln, ins), (ex, sm) in zip(lines, data):
    by[(f, ln)][0] += ex
    by[(f, ln)][1] += sm
tot, ts = sum(v[0] for v in by.values()), max(sum(v[1] for v in by.values()), 1)
print("SASS", len(lines), "instr/warp %.0f" % (tot / warps))
srcs = {}
for (f, ln), v in sorted(by.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    for d in ("h-slam_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f or "")
        if f and os.path.exists(p):
            srcs.setdefault(p, open(p).read().splitlines())
            if ln and ln <= len(srcs[p]):
                text = srcs[p][ln - 1].strip()[:100]
            break
    print("%-18s %5s  exec/warp %7.1f (%.3f)  samples %.3f   %s" % (f, ln, v[0] / warps, v[0] / tot, v[1] / ts, text))

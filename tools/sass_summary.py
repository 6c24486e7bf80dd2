#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (cuobjdump -sass): the evidence that the sm_100a-specific
instructions are really in the binary.  usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "h-slam_b200", "libhslam_b200.so")
WATCH = [("UTMALDG", "tensor-map TMA load (cp.async.bulk.tensor)"), ("UBLKCP", "bulk copy (cp.async.bulk)"),
         ("SYNCS", "mbarrier arrive / try_wait"), ("UCGABAR", "cluster barrier"), ("LDG.E.ENL2.256", "256-bit global load"),
         ("LDG.E.256", "256-bit global load"), ("STG.E.ENL2.256", "256-bit global store"), ("STG.E.256", "256-bit global store"),
         ("FFMA2", "packed fp32 FMA"), ("FMUL2", "packed fp32 mul"), ("FADD2", "packed fp32 add"), ("DMMA", "fp64 tensor-core MMA"),
         ("DFMA", "fp64 FMA"), ("MUFU", "SFU"), ("LDS", "shared load"), ("STS", "shared store"), ("LDL", "local load (spill)"),
         ("STL", "local store (spill)"), ("BAR.SYNC", "block barrier"), ("SHFL", "warp shuffle"), ("ATOM", "atomic (global)"),
         ("RED", "reduction (global)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    counts = collections.OrderedDict()
    total = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kern = re.sub(r"\(.*", "", kern).replace("hs::", "")
            counts[kern] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_all"] += 1
            for w, _ in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    counts[kern][w] += 1
    print("# cuobjdump -sass h-slam_b200/libhslam_b200.so  (sm_100a) — instruction counts per kernel")
    print("# " + "; ".join(f"{w}: {d}" for w, d in WATCH))
    for k, c in counts.items():
        items = [f"{w}={c[w]}" for w, _ in WATCH if c[w]]
        print(f"{k}: instructions={c['_all']}  " + " ".join(items))
        for w, _ in WATCH:
            total[w] += c[w]
    print("TOTAL: " + " ".join(f"{w}={total[w]}" for w, _ in WATCH if total[w]))


if __name__ == "__main__":
    sys.exit(main())

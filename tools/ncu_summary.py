#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.
   launches <csv>            : per-kernel count / avg / total / share from a `--metrics gpu__time_duration.sum` CSV log
   raw <ncu-rep>             : key metrics per captured launch from `ncu -i rep --page raw --csv`
"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in data:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':40s} {'n':>5s} {'avg_us':>10s} {'total_us':>11s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:40s} {len(v):5d} {sum(v)/len(v):10.2f} {sum(v):11.1f} {sum(v)/tot:7.3f}")


WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_act_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit"),
    ("lts__t_sectors.sum", "l2_sectors"),
    ("LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefront_pct"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible_warps"),
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    for r in data:
        print("==", r[ki].split("(")[0].replace("void ", ""))
        for m, short in WANT:
            if m in hdr:
                i = hdr.index(m)
                print(f"   {short:22s} {r[i]:>16s} {units[i]}")
        # achieved bandwidths (what north_star asks ncu to report): bytes / duration of this (cold-cache) launch
        def val(m, to):
            i = hdr.index(m)
            v, u = float(r[i].replace(",", "")), units[i]
            scale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                     "sector": 32.0}.get(u, 1.0)
            return v * scale
        try:
            t = val("gpu__time_duration.sum", "s")
            hbm = (val("dram__bytes_read.sum", "B") + val("dram__bytes_write.sum", "B")) / t / 1e9
            l2 = val("lts__t_sectors.sum", "B") / t / 1e9
            print(f"   {'hbm_GBps_achieved':22s} {hbm:16.1f} GB/s  ({hbm / 6555.2:.2f} of the measured 6555 GB/s copy peak)")
            print(f"   {'l2_GBps_achieved':22s} {l2:16.1f} GB/s")
        except (ValueError, KeyError):
            pass


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])

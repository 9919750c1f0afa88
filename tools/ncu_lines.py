#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample shares from an ncu report captured with
--import-source on:  python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, v = rows[0], rows[-1]
    for key in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
                "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"):
        for i, n in enumerate(h):
            if n == key:
                print("%-70s %s %s" % (n, v[i], rows[1][i] if len(rows) > 2 else ""))
    stall = {}
    for i, n in enumerate(h):
        if n.startswith("smsp__average_warp_latency_issue_stalled_") or n.startswith("smsp__average_warps_issue_stalled_"):
            if n.endswith(".ratio") or n.endswith("_per_issue_active.ratio"):
                try:
                    stall[n.split("stalled_")[1].split(".")[0]] = float(v[i])
                except ValueError:
                    pass
    tot = sum(x for k, x in stall.items() if "not_issued" not in k) or 1.0
    print("stalls:", ", ".join("%s %.1f%%" % (k.replace("_per_warp_active", ""), 100 * x / tot)
                               for k, x in sorted(stall.items(), key=lambda kv: -kv[1])[:9] if "not_issued" not in k))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, agg = None, collections.defaultdict(lambda: [0, 0, ""])
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Line No", "Function Name"):
            continue
        if len(r) > 8 and r[2] == "-" and r[0].isdigit():
            a = agg[(cur, int(r[0]))]
            a[0] += int(r[7]) if r[7].isdigit() else 0
            a[1] += int(r[6]) if r[6].isdigit() else 0
            a[2] = r[1]
    ti = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    byf = collections.defaultdict(lambda: [0, 0])
    for (f, l), a in agg.items():
        byf[f][0] += a[0]
        byf[f][1] += a[1]
    for f, a in sorted(byf.items(), key=lambda kv: -kv[1][0]):
        print("%-28s %5.1f%% instr %5.1f%% samples" % (f, 100.0 * a[0] / ti, 100.0 * a[1] / ts))
    for (f, l), a in sorted(agg.items()):
        if 100.0 * a[0] / ti >= thr or 100.0 * a[1] / ts >= thr:
            print("%-24s %4d %5.2f%%i %5.2f%%s  %s" % (f[:24], l, 100.0 * a[0] / ti, 100.0 * a[1] / ts, a[2].strip()[:96]))


if __name__ == "__main__":
    main()

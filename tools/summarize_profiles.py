#!/usr/bin/env python3
"""Condense ncu outputs from gpurun_out/ into small text files under profiles/ (the .ncu-rep itself is too big for git).
usage: summarize_profiles.py <tag> [--rep gpurun_out/x.ncu-rep] [--launches gpurun_out/y.csv]"""
import argparse
import collections
import csv
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_alu.sum",
           "sm__inst_executed_pipe_alu.sum", "sm__sass_inst_executed_op_integer_pred_on.sum"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    popc = [i for i, h in enumerate(hdr) if "popc" in h.lower()]
    with open(out, "w") as f:
        f.write("kernel," + ",".join(f"{m} [{units[i]}]" for m, i in cols) + "\n")
        for r in rows[2:]:
            f.write(r[hdr.index("Kernel Name")].split("(")[0] + "," + ",".join(r[i].replace(",", "") for _, i in cols) + "\n")
    print("wrote", out, "extra popc-like columns:", [hdr[i] for i in popc][:5])


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# per-kernel totals of gpu__time_duration.sum over the whole bench.py run under ncu (cold-cache, serialised:\n"
                "# compare SHARES, not absolutes)\nkernel,launches,total_us,avg_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{t / 1e3:.1f},{t / 1e3 / n:.1f},{100 * t / tot:.2f}\n")
    print(open(out).read())


ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--rep")
ap.add_argument("--launches")
a = ap.parse_args()
(ROOT / "profiles").mkdir(exist_ok=True)
if a.rep:
    full(a.rep, ROOT / "profiles" / f"{a.tag}_ncu_full_summary.csv")
if a.launches:
    launches(a.launches, ROOT / "profiles" / f"{a.tag}_launches_summary.csv")

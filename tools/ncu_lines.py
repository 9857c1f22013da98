#!/usr/bin/env python
"""Attribute the executed warp instructions of one kernel of an ncu report (captured with --import-source on, built with
-lineinfo) to CUDA source lines:

    python tools/ncu_lines.py gpurun_out/prof_r2_a.ncu-rep k_fast_cell [--min-pct 0.5] [--csv out.csv]

Reads `ncu --page source --print-source cuda,sass --csv`: every SASS instruction carries its own "Instructions Executed";
the view lists, per source file, each source line (with the totals of its instructions) followed by the SASS lines that came
from it.  The table is what profiles/*_lines.csv hold.
"""
import argparse
import csv
import io
import os
import subprocess
from collections import defaultdict


def source_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", os.path.abspath(rep), "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                          f"::regex:{kernel}:"], capture_output=True, text=True, cwd="/tmp").stdout
    return list(csv.reader(io.StringIO(out)))


def per_line(rows):
    """-> {(file, line): [warp_instr, thread_instr, samples, text]}"""
    acc = defaultdict(lambda: [0, 0, 0, ""])
    fname, hdr, cur = None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            hdr = None
            continue
        if r[0] == "Line No":
            hdr = r
            ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or fname is None:
            continue
        if r[0].strip().isdigit():                      # a source line carries the totals of the SASS rows listed below it
            cur = (fname, int(r[0]))
            acc[cur][3] = r[1].strip()
            try:
                acc[cur][0] += int(r[ci]); acc[cur][1] += int(r[ti]); acc[cur][2] += int(r[si] or 0)
            except (ValueError, IndexError):
                pass
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel")
    ap.add_argument("--min-pct", type=float, default=0.5)
    ap.add_argument("--csv")
    a = ap.parse_args()
    acc = per_line(source_rows(a.rep, a.kernel))
    tot = sum(v[0] for v in acc.values()) or 1
    tots = sum(v[2] for v in acc.values()) or 1
    items = sorted(acc.items(), key=lambda kv: (kv[0][0], kv[0][1]))
    print(f"# kernel {a.kernel}: {tot/1e6:.2f} M warp instructions, {tots} stall samples")
    lines = ["file,line,warp_instr,pct,thread_instr,samples_pct,source"]
    for (f, ln), (wi, ti, sm, txt) in items:
        if wi == 0:
            continue
        lines.append(f'{f},{ln},{wi},{100*wi/tot:.2f},{ti},{100*sm/tots:.2f},"{txt[:140].replace(chr(34), chr(39))}"')
        if 100 * wi / tot >= a.min_pct:
            print(f"{f}:{ln:<5d} {wi/1e6:8.2f}M {100*wi/tot:5.1f}%  avg_thr {ti/max(wi,1):4.1f}  stall {100*sm/tots:4.1f}%  {txt[:100]}")
    if a.csv:
        open(a.csv, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep (source page, cuda+sass view).
usage: ncu_lines.py <rep> <kernel-regex> [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, out, hdr = "", [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif len(r) > 7 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 7 and r[0].isdigit():
        try:
            out.append((cur_file, int(r[0]), r[1].strip()[:90], int(r[4] or 0), int(r[7] or 0)))
        except ValueError:
            pass
tot_i = sum(o[4] for o in out) or 1
tot_s = sum(o[3] for o in out) or 1
print(f"total inst {tot_i}  samples {tot_s}")
for f, ln, src, s, i in sorted(out, key=lambda o: -o[4])[:top]:
    print(f"{100 * i / tot_i:5.1f}% inst {100 * s / tot_s:5.1f}% stall  {f}:{ln}  {src}")

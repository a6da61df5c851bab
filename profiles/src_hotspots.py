#!/usr/bin/env python
"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv
--print-source cuda,sass) by CUDA source line: stall samples and executed
instructions.  Usage: python profiles/src_hotspots.py report.ncu-rep [min_pct]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
only = sys.argv[3] if len(sys.argv) > 3 else None      # substring of the kernel name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file, last_line, cur_fn = None, None, ""
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != "":
        last_line = (cur_file, int(r[0]), r[1].strip()[:72])
    if only is not None and only not in cur_fn:
        continue
    if len(r) > 2 and r[2].startswith("0x") and last_line is not None:
        i_s = hdr.index("Warp Stall Sampling (All Samples)")
        i_e = hdr.index("Instructions Executed")
        a = agg.setdefault(last_line, [0, 0, 0])
        a[0] += int(r[i_s] or 0)
        a[1] += int(r[i_e] or 0)
        a[2] += 1
tot = sum(a[0] for a in agg.values()) or 1
totex = sum(a[1] for a in agg.values()) or 1
print("stall samples", tot, " warp instructions executed", totex)
for k, a in agg.items():
    if a[0] > tot * minpct / 100 or a[1] > totex * minpct / 100:
        print("%-18s %4d %5.1f%% smp %5.1f%% instr (%3d sass) | %s" % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / totex, a[2], k[2]))

#!/usr/bin/env python
"""Summarise ncu outputs into tracked text files under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r01.csv  > profiles/launches_r01.md   (r01b = end of round 1)
  python profiles/summarize.py full gpurun_out/prof_r01.ncu-rep       > profiles/kernels_r01.md  (also writes traffic.json)
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name):
    import re
    m = re.search(r"strided_sweep(_tma)?<([^>]*)>", name)
    if m:
        args = [a.replace("(int)", "").replace("(bool)", "").strip() for a in m.group(2).split(",")]
        final = args[2] in ("1", "true")
        if m.group(1):
            how = "TMA prefetch" if args[3] in ("1", "true") else "cp.async prefetch"
            return "strided_sweep_tma (%s, persistent, %s)" % ("z" if final else "y", how)
        return "strided_sweep (%s, register loads)" % ("z" if final else "y")
    for key, lab in (("sweep_xw2_kernel", "sweep_xw2_kernel (sweep_x: warp per line pair, two lines per lane, TMA boxes)"),
                     ("sweep_xw_kernel", "sweep_xw_kernel (sweep_x: warp per line, TMA boxes)"),
                     ("sweep_xt_kernel", "sweep_xt_kernel (sweep_x: TMA-fed patches)"),
                     ("sweep_xf_kernel", "sweep_xf_kernel (sweep_x: 5-point RHS + folded solve)"),
                     ("sweep_x_kernel", "sweep_x_kernel (x: RHS + solve)"),
                     ("z_forward", "z_forward (slab)"), ("z_backward", "z_backward (slab)"),
                     ("rhs_kernel", "rhs_kernel (fallback)"), ("thomas_kernel", "thomas_kernel (fallback)")):
        if key in name:
            return lab
    return name.split("(")[0][-60:]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of `python bench.py --steps 4 --warmup 3 --no-cpu-baseline` (kernels of this library only: `-k regex:sweep|thomas_kernel|rhs_kernel|z_forward|z_backward`)\n")
    print("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.3f | %.1f | %.1f %% |" % (k, a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))


WANT = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->L1 read sectors"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.per_cycle_active", "issue active / cycle"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp instructions")]


def full(path, cells=512 ** 3, write_traffic=True):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    print("# ncu --set full (--clock-control none) of the three sweep kernels, %d cells, one launch each\n" % cells)
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        print("## %s\n" % name)
        print("| metric | value |\n|---|---|")
        vals = {}
        for key, lab in WANT:
            if key in hdr:
                i = hdr.index(key)
                vals[key] = r[i]
                print("| %s | %s %s |" % (lab, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.25:
                    stalls.append((v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        print("| top stalls (warps per issue) | %s |" % ", ".join("%s %.2f" % (n, v) for v, n in sorted(stalls, reverse=True)))
        print()

        def gb(key):
            i = hdr.index(key)
            v = float(r[i])
            u = units[i].lower()
            return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        axis = "x" if "sweep_x" in name else ("y" if "(y" in name else "z")
        traffic[axis] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
    if write_traffic:
        traffic["cells"] = cells
        traffic["source"] = "gpurun_out/%s, summarised in profiles/kernels_r02.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)" % os.path.basename(path)
        json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print("DRAM bytes per launch (read+write)%s:" % (" -> profiles/traffic.json" if write_traffic else ""), traffic)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:       # full REPORT [cells] [notraffic]
        full(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 512 ** 3, "notraffic" not in sys.argv)

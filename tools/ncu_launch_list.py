#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` log into the launch list kept under profiles/:
one line per (kernel, grid) of this library's kernels with launch count, mean duration and share of the
library's GPU time.      python tools/ncu_launch_list.py gpurun_out/launches.csv > profiles/rNN_launch_list.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, vi, mi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum" or "<unnamed>" not in r[ki] or "at::" in r[ki]:
        continue
    agg.setdefault((r[ki][:64], r[gi]), []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("%-66s %-16s %5s %10s %7s" % ("kernel", "grid", "n", "mean us", "share"))
for (k, g), v in agg.items():
    print("%-66s %-16s %5d %10.2f %7.3f" % (k, g, len(v), sum(v) / len(v) / 1e3, sum(v) / tot))

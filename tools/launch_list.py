#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) per kernel.
usage: python tools/launch_list.py gpurun_out/launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ki].replace("void ", "").split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total [us] | per launch [us] | share |")
print("|---|---|---|---|---|")
for k, (n, t) in agg.items():
    print(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {t / tot:.3f} |")

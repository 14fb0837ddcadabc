#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches / total / mean
and share of the listed device time.  usage: summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", "")), r["Grid Size"], r["Block Size"]))
agg = OrderedDict()
for name, ns, grid, block in rows:
    short = name.split("(")[0][-90:]
    k = (short, grid, block)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += ns
tot = sum(a[1] for a in agg.values()) or 1.0
print("| kernel | grid | block | launches | total us | mean us | share |")
print("|---|---|---|---|---|---|---|")
for (short, grid, block), (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %s | %s | %d | %.1f | %.1f | %.1f%% |" % (short, grid, block, n, ns / 1e3, ns / n / 1e3, 100 * ns / tot))

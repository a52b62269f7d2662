#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).  usage: launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hd = next(r for r in rows if "Kernel Name" in r)
ki, vi, ui = hd.index("Kernel Name"), hd.index("Metric Value"), hd.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hd or len(r) <= vi or r[ki] == "Kernel Name":
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^.*::", "", name)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values()) or 1.0
print(f"{'kernel':44s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:44]:44s} {a[0]:8d} {a[1]:12.1f} {100 * a[1] / tot:6.1f}%")

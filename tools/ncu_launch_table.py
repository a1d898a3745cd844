#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/ncu_launch_table.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"]
    k = k.replace("void ", "").replace("unnamed>::", "").replace("vu::", "")[:64] + " grid=" + row["Grid Size"].replace(" ", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'launches':>8s} {'total us':>10s} {'us/launch':>10s} {'share':>6s}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:10.1f} {100 * a[1] / tot:5.1f}%  {k}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")

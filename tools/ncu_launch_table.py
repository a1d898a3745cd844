#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (``--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--csv``):  python tools/ncu_launch_table.py launches.csv [--skip-at] [--pixels P --frames N]

With the dram metrics the table also carries the bytes each kernel moved; ``--pixels`` / ``--frames`` print them in units
of P (pixels per frame) per frame.  ``--skip-at`` drops torch's own kernels (data generation of the benches)."""
import argparse
import collections
import csv

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--skip-at", action="store_true")
ap.add_argument("--pixels", type=float, default=0)
ap.add_argument("--frames", type=float, default=0)
ap.add_argument("--only", default="", help="substring filter on kernel names (comma list)")
args = ap.parse_args()

lines = [l for l in open(args.csv) if not l.startswith("==")]
launch = collections.OrderedDict()   # ID -> [name, us, rd, wr]
for row in csv.DictReader(lines):
    k = row["Kernel Name"]
    if args.skip_at and ("at::" in k or "at_cuda" in k or "cub::" in k or "elementwise" in k):
        continue
    name = k.replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("vu::", "")[:60] + " grid=" + row["Grid Size"].replace(" ", "")
    if args.only and not any(s in name for s in args.only.split(",")):
        continue
    a = launch.setdefault(row["ID"], [name, 0.0, 0.0, 0.0])
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        a[1] = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    elif m.startswith("dram__bytes"):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        a[2 if "read" in m else 3] += v * mult
agg = collections.OrderedDict()
for name, us, rd, wr in launch.values():
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += us
    a[2] += rd
    a[3] += wr
tot = sum(a[1] for a in agg.values())
have_dram = any(a[2] or a[3] for a in agg.values())
unit = args.pixels * args.frames
hdr = f"{'launches':>8s} {'total us':>10s} {'us/launch':>10s} {'share':>6s}"
if have_dram:
    hdr += f" {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s}" + (f" {'P/frame':>8s}" if unit else "")
print(hdr + "  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    line = f"{a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:10.1f} {100 * a[1] / tot:5.1f}%"
    if have_dram:
        line += f" {a[2] / 1e6:9.1f} {a[3] / 1e6:9.1f} {(a[2] + a[3]) / a[1] / 1e3:7.0f}" + (f" {(a[2] + a[3]) / unit:8.2f}" if unit else "")
    print(line + "  " + k)
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches" +
      (f"; dram {sum(a[2] for a in agg.values()) / 1e6:.1f} MB read + {sum(a[3] for a in agg.values()) / 1e6:.1f} MB written" if have_dram else "") +
      (f" = {sum(a[2] + a[3] for a in agg.values()) / unit:.2f} P per frame" if have_dram and unit else ""))

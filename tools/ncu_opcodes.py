#!/usr/bin/env python
"""Opcode histogram (by executed warp instructions) of an `ncu --page source --csv` dump.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python tools/ncu_opcodes.py src.csv [warps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
cnt, samples = collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    op = r[ci["Source"]].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = ".".join(o.split(".")[:2]) if o.startswith(("VABSDIFF", "SHFL", "LDG", "IMAD")) else o.split(".")[0]
    cnt[o] += int(r[ci["Instructions Executed"]])
    samples[o] += int(r[ci["# Samples"]])
tot, stot = sum(cnt.values()), sum(samples.values())
print(f"{'opcode':16s} {'warp instr':>12s} {'share':>6s} {'samples':>7s}" + ("  per warp" if warps else ""))
for o, n in cnt.most_common(30):
    print(f"{o:16s} {n:12d} {100 * n / tot:5.1f}% {100 * samples[o] / max(stot, 1):6.1f}%" + (f" {n / warps:9.1f}" if warps else ""))
print("total", tot)

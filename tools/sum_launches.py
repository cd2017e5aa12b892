#!/usr/bin/env python
"""Per-kernel totals of the second half of an ncu launch list (one steady step of tools/trace_step.py --steps 2)."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
data = rows[1:]
half = len(data) // 2
tot = collections.OrderedDict()
for r in data[half:]:
    key = r[ix["Kernel Name"]].split("(")[0].replace("void ", "") + r[ix["Grid Size"]]
    t = float(r[ix["Metric Value"]]) / 1000
    tot.setdefault(key, [0, 0])
    tot[key][0] += t
    tot[key][1] += 1
for k, v in tot.items():
    print(f"{k:62s} n={v[1]:3d} total={v[0]:8.1f} us  each={v[0] / v[1]:7.1f}")
print("sum %.1f us, launches %d" % (sum(v[0] for v in tot.values()), len(data) - half))

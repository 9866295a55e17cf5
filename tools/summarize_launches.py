#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (+ grid)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
detail = len(sys.argv) > 2
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    name = re.sub(r"^void |dgtd::", "", name)[:70]
    if detail:
        name += " grid=" + row.get("Grid Size", "")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    e = tot.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += v
s = sum(v[1] for v in tot.values())
print(f"total {s:.1f} us over {sum(v[0] for v in tot.values())} launches")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1]/v[0]:.1f} | {100*v[1]/s:.1f}% |")

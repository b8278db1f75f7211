#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total, avg, share.
usage: scripts/launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit.startswith("ns") else v * 1e3 if unit.startswith("ms") else v
    k = r[ix["Kernel Name"]]
    agg[k][0] += 1
    agg[k][1] += us
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':90s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:90]:90s} {n:8d} {t:10.1f} {t / n:8.2f} {100 * t / tot:6.1f}%")

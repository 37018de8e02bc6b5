#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of total, launches, average us.

usage: launch_summary.py <launches.csv> [title]"""
import collections, csv, sys

path = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else path
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg[r[ix["Kernel Name"]]]
    a[0] += us; a[1] += 1
tot = sum(a[0] for a in agg.values())
n = sum(a[1] for a in agg.values())
print(f"# {title}\n# total {tot / 1e3:.1f} ms over {n} launches (cold-cache, serialised: compare SHARES)\n")
print("| share | total us | launches | avg us | kernel |\n|---|---|---|---|---|")
for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {100 * us / tot:.2f}% | {us:.1f} | {c} | {us / c:.1f} | `{k[:110]}` |")

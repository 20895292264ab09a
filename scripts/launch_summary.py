#!/usr/bin/env python
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: scripts/launch_summary.py gpurun_out/launches.csv > profiles/r1_launches.txt"""
import csv
import sys
from collections import OrderedDict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
d = OrderedDict()
for r in rows[start:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        v_us = v / 1000.0 if u.startswith("ns") else (v if u.startswith("us") else v * 1000.0)
        d.setdefault(r[ki].split("(")[0], []).append(v_us)
tot = sum(sum(v) for v in d.values())
print(f"# {sys.argv[1]}: {sum(len(v) for v in d.values())} launches, {tot/1000:.2f} ms of kernel time "
      f"(cold-cache, serialised under ncu: compare shares, not absolutes)")
print(f"{'kernel':44s} {'n':>5s} {'mean_us':>10s} {'min_us':>10s} {'max_us':>10s} {'share':>7s}")
for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:44s} {len(v):5d} {sum(v)/len(v):10.1f} {min(v):10.1f} {max(v):10.1f} {sum(v)/tot:7.3f}")

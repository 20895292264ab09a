#!/usr/bin/env python
"""Per-source-line hot spots of a kernel from an `ncu --set full --import-source on` capture (built with -lineinfo).
usage: scripts/ncu_lines.py prof.ncu-rep [kernel-name-substring] [top N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, hdr, agg, fn, seen_fn = None, None, {}, "", set()
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        fn = r[1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        hdr = {k: i for i, k in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or want not in fn:
        continue
    if r[0] == "" or not r[0].isdigit():
        continue
    try:
        samples = int(r[hdr["# Samples"]] or 0)
        inst = int(r[hdr["Instructions Executed"]] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur_file, int(r[0])), [0, 0, r[1].strip()[:100]])
    a[0] += samples
    a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f"# {rep} {want}: {tot_s} stall samples, {tot_i} warp instructions (all captured launches of the kernel)")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{k[0]}:{k[1]:<4d} samples {100 * a[0] / tot_s:5.1f}%  inst {100 * a[1] / tot_i:5.1f}%  {a[2]}")

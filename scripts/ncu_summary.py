#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the LAST
step (from the last occurrence of --first-kernel to the end), per-kernel time and share."""
import argparse
import csv

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--first-kernel", default="pack_ref")
a = ap.parse_args()
lines = [l for l in open(a.csv) if not l.startswith("==")]
rows = [(r["Kernel Name"], float(r["Metric Value"]) / 1e3, r.get("Grid Size", ""), r.get("Block Size", ""))
        for r in csv.DictReader(lines)]
starts = [i for i, r in enumerate(rows) if a.first_kernel in r[0]]
seg = rows[starts[-1]:] if starts else rows
tot = sum(r[1] for r in seg)
print(f"# {len(seg)} launches in the last step, {tot:.1f} us of kernel time (cold-cache, serialised under ncu)")
print(f"{'us':>9} {'share':>6}  {'grid':>16} {'block':>14}  kernel")
for n, t, g, b in seg:
    print(f"{t:9.2f} {100 * t / tot:5.1f}%  {g:>16} {b:>14}  {n[:100]}")

#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv --kernel-name regex:X`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


data = [(num(r[i_s]), num(r[i_ex]), n, r[i_src].strip()) for n, r in enumerate(rows[hi + 1:]) if len(r) > i_ex]
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "SASS instrs", len(data), "warp-instrs executed", sum(d[1] for d in data))
for s, e, n, src in sorted(data, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  ex={e:8d}  #{n:4d} {src[:100]}")

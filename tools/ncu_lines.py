#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
instructions executed and stall samples (first kernel instance in the file only)."""
import csv
import sys
import collections

path, kfile = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.reader(open(path)))
cur_file = None
agg = collections.defaultdict(lambda: [0, 0, ""])
hdr = None
seen_funcs = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed")
        iS = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    if r[2] == "-":  # cuda source line header row (no SASS address): holds the aggregated numbers
        key = (cur_file.split("/")[-1], ln)
        agg[key][0] += int(r[iI] or 0)
        agg[key][1] += int(r[iS] or 0)
        agg[key][2] = r[1].strip()[:110]
tot_i = sum(v[0] for v in agg.values()) or 1
tot_s = sum(v[1] for v in agg.values()) or 1
print(f"total inst {tot_i}  samples {tot_s}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    if kfile and kfile not in f:
        continue
    print(f"{f}:{ln:4d} inst {100*v[0]/tot_i:5.1f}%  stall {100*v[1]/tot_s:5.1f}%  {v[2]}")

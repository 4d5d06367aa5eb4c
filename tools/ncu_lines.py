#!/usr/bin/env python
"""Per CUDA source line totals (instructions executed, stall samples) of one kernel:
ncu -i rep --page source --csv --print-source cuda,sass -k regex:<kernel> > f.csv; python tools/ncu_lines.py f.csv [min_pct]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
i_s, i_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
lines = [(r[0], r[1], num(r[i_e]), num(r[i_s])) for r in rows[hi + 1:] if len(r) > i_e and r[0].isdigit()]
tot, ts = sum(l[2] for l in lines), sum(l[3] for l in lines)
print("warp instructions", tot, "samples", ts)
for no, src, e, s in lines:
    if 100.0 * e / tot >= minpct or 100.0 * s / max(ts, 1) >= minpct:
        print("%5s %5.1f%% inst %5.1f%% smp  %s" % (no, 100.0 * e / tot, 100.0 * s / max(ts, 1), src.strip()[:110]))

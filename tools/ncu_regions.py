#!/usr/bin/env python
"""Groups the SASS of one kernel (ncu --page source --csv) into runs of equal execution count: where instructions go."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1]))]
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # n-th kernel in the file
hi = heads[which]
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
print("kernel:", rows[hi - 1][:2])
hdr = rows[hi]
body = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
i_src, i_ex, i_s = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[i_ex] or 0) for r in body)
ts = sum(int(r[i_s] or 0) for r in body)
print("sass lines", len(body), "warp instructions", tot, "samples", ts)
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
last, start, acc, sacc = None, 0, 0, 0
def flush(k):
    if last is not None and (100.0 * acc / tot >= minpct or 100.0 * sacc / max(ts, 1) >= minpct):
        print("%4d-%4d n=%4d exec=%8d total=%9d (%4.1f%% inst, %4.1f%% samples)  %s" % (
            start, k - 1, k - start, last, acc, 100.0 * acc / tot, 100.0 * sacc / max(ts, 1), body[start][i_src].strip()[:60]))
for k, r in enumerate(body):
    ex = int(r[i_ex] or 0)
    if last is None or abs(ex - last) > 0.02 * max(ex, last, 1):
        flush(k)
        start, acc, sacc = k, 0, 0
    last = ex
    acc += ex
    sacc += int(r[i_s] or 0)
flush(len(body))

#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from `ncu -i rep --page source --csv -k <kernel>` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
heads = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # n-th kernel section of the file
print("kernel:", rows[heads[which] - 1][1][:80])
end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
rows = [None, rows[heads[which]]] + [r for r in rows[heads[which] + 1:end] if len(r) == len(rows[heads[which]])]
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
tot = sum(int(r[i_s] or 0) for r in body)
print("total samples", tot, "instructions", sum(int(r[i_ex] or 0) for r in body))
top = sorted(range(len(body)), key=lambda k: -int(body[k][i_s] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for k in sorted(top):
    r = body[k]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    print("%4d %6s %5.1f%% %-70s %s" % (k, r[i_ex], 100.0 * int(r[i_s] or 0) / max(tot, 1), r[i_src].strip()[:70], st))

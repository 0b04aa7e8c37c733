#!/usr/bin/env python
"""Print the SASS lines with the most warp-stall samples from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu -i rep --page source --csv --launch-skip N --launch-count 1 > src.csv; python tools/ncu_top_stalls.py src.csv [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
iS, iSrc, iEx = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[iS]) for r in data if r[iS].isdigit())
print("kernel:", rows[0][1] if len(rows[0]) > 1 else "?", "| total samples", tot, "| sass lines", len(data))
agg = {}
for r in data:
    for i in stall_cols:
        if r[i].isdigit():
            agg[h[i]] = agg.get(h[i], 0) + int(r[i])
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][iS]) if t[1][iS].isdigit() else 0)[:n_top]:
    st = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print("%5d %6s %5.1f%% ex=%-8s %-64s %s" % (idx, r[iS], 100 * int(r[iS]) / max(tot, 1), r[iEx], r[iSrc].strip()[:64], st))

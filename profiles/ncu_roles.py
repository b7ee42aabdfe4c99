"""Aggregate warp-stall samples of an `ncu --page source --csv` dump by instruction-index range (= warp role).
usage: python profiles/ncu_roles.py src.csv b0 b1 b2 ...   (range boundaries; prints samples and stall mix per range)"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
bounds = [int(v) for v in sys.argv[2:]]
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    k = sum(1 for b in bounds if n >= b)
    a = agg.setdefault(k, [0, Counter(), 0])
    a[0] += int(r[idx["# Samples"]] or 0)
    a[2] += int(r[idx["Instructions Executed"]] or 0)
    for c in stall_cols:
        a[1][c[6:]] += int(r[idx[c]] or 0)
for k in sorted(agg):
    lo = bounds[k - 1] if k else 0
    hi = bounds[k] if k < len(bounds) else len(rows) - 2
    s, c, ex = agg[k]
    print("#%5d-%5d samples %6d  warp-insts %9d  %s" % (lo, hi, s, ex, " ".join("%s:%d" % kv for kv in c.most_common(6) if kv[1])))

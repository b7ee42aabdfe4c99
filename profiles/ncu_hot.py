"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`.
usage: python profiles/ncu_hot.py src.csv [topN]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
total = 0
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr):
        continue
    s = int(r[idx["# Samples"]] or 0)
    total += s
    stalls = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    data.append((s, n, r[idx["Source"]].strip(), r[idx["Instructions Executed"]], stalls))
print("total samples", total)
for s, n, src, ex, stalls in sorted(data, reverse=True)[:top]:
    print("%6d (%4.1f%%) #%-5d %-58s exec=%-9s %s" % (s, 100.0 * s / total, n, src[:58], ex, " ".join("%s:%d" % (c, v) for v, c in stalls if v)))

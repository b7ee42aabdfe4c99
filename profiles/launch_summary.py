"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) per kernel.
usage: python profiles/launch_summary.py launches.csv [steps_in_capture]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]]
    name = re.sub(r"\(.*$", "", name).replace("void ", "")
    ns = float(r[ix["Metric Value"]].replace(",", ""))
    agg[name][0] += 1
    agg[name][1] += ns
    total += ns
print("%-70s %8s %10s %8s %6s" % ("kernel", "launches", "total_us", "avg_us", "share"))
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s %8d %10.1f %8.2f %5.1f%%" % (name[:70], n, ns / 1e3, ns / 1e3 / n, 100 * ns / total))
print("total %.1f us over %d launches (%.1f us/step at %g steps)" % (total / 1e3, sum(v[0] for v in agg.values()), total / 1e3 / steps, steps))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised
durations: compare SHARES, not absolutes).  usage: python profiles/summarize_launches.py <launches.csv>"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000.0 if row["Metric Unit"] == "ns" else (v * 1000.0 if row["Metric Unit"] == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
        tot += v
    print("total %.1f us over %d launches" % (tot, sum(a[0] for a in agg.values())))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-74s n=%3d sum=%9.1f us (%4.1f%%) max=%8.1f" % (k, a[0], a[1], 100 * a[1] / tot, a[2]))


if __name__ == "__main__":
    main(sys.argv[1])

"""Summarise an ncu launch list per kernel.  Input: `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]
--clock-control none --csv` log.  usage: python profiles/launch_summary.py launches.csv [steps_in_capture]
Per kernel: launches, summed / average duration, share of the summed kernel time, and (when captured) DRAM bytes per launch.
Per family (tc_conv_kernel, tc_wgrad_kernel + reduce, BatchNorm / gather glue, loss, optimiser): time and DRAM traffic per step."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
per_id = defaultdict(dict)
names = {}
for r in rd:
    kid = r[ix["ID"]]
    names[kid] = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).replace("void ", "").replace("hpfg::", "")
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]] if "Metric Unit" in ix else ""
    m = r[ix["Metric Name"]]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    elif m.startswith("gpu__time"):
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
    per_id[kid][m] = v
agg = defaultdict(lambda: [0, 0.0, 0.0])
for kid, m in per_id.items():
    a = agg[names[kid]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
total = sum(a[1] for a in agg.values())
have_dram = any(a[2] > 0 for a in agg.values())
print("%-72s %8s %10s %8s %6s %12s" % ("kernel", "launches", "total_us", "avg_us", "share", "dram MB/launch" if have_dram else ""))
for name, (n, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %8d %10.1f %8.2f %5.1f%% %12s" % (name[:72], n, ns / 1e3, ns / 1e3 / n, 100 * ns / total, ("%.2f" % (by / n / 1e6)) if have_dram else ""))
print("total %.1f us over %d launches (%.1f us/step at %g steps)" % (total / 1e3, sum(v[0] for v in agg.values()), total / 1e3 / steps, steps))


def family(name):
    if name.startswith("tc_conv_kernel"):
        return "tc_conv_kernel (fprop / dgrad / 1x1 / logits)"
    if name.startswith("tc_wgrad"):
        return "tc_wgrad_kernel + tc_wgrad_reduce"
    if name.startswith("bn_") or "pool" in name or name.startswith("up") or "pad_to" in name or "dropout" in name or "nchw" in name:
        return "BatchNorm / pool / upsample / pad glue"
    if name.startswith("loss") or "argmax" in name or "softmax_mse" in name or "ict_mix" in name or "dice" in name:
        return "loss"
    if name.startswith("sgd") or name.startswith("ema") or name.startswith("tc_pack"):
        return "optimiser + weight packing"
    return "other (" + name[:30] + ")"


fam = defaultdict(lambda: [0, 0.0, 0.0])
for name, (n, ns, by) in agg.items():
    f = fam[family(name)]
    f[0] += n
    f[1] += ns
    f[2] += by
print("\nper family, per step:")
for name, (n, ns, by) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print("  %-48s %6.1f launches %9.1f us %5.1f%% %s" % (name, n / steps, ns / 1e3 / steps, 100 * ns / total, ("%9.1f MB DRAM" % (by / 1e6 / steps)) if have_dram else ""))
if have_dram:
    print("  %-48s %26s %9.1f MB DRAM" % ("all kernels", "", sum(f[2] for f in fam.values()) / 1e6 / steps))

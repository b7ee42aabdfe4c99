"""Print selected metrics from `ncu -i X.ncu-rep --page raw --csv` output.  usage: python profiles/ncu_raw.py raw.csv [substr ...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
want = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
                        "sm__pipe_tensor", "sm__warps_active.avg.pct", "launch__registers_per_thread", "bank_conflicts",
                        "lts__t_bytes.sum", "sm__throughput.avg.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                        "smsp__average_warp", "issue_stalled", "launch__grid_size", "shared_mem", "sm__cycles_active.avg",
                        "smsp__inst_executed.sum", "tma", "smsp__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want):
        print("%-110s %s %s" % (h, v, u))

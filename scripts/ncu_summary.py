"""Summarise one `ncu --set full` report:  ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for k in KEYS:
    for i, h in enumerate(hdr):
        if h == k:
            print(f"{h} [{units[i]}] = {vals[i]}")
for i, h in enumerate(hdr):
    if h in KEYS:
        continue
    if ("pipe_tensor" in h and "pct" in h) or ("pipe_xu" in h and "pct" in h) or ("pipe_fma" in h and "pct_of_peak_sustained_active" in h) \
            or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(vals[i] or 0) > 0.15):
        print(f"{h} [{units[i]}] = {vals[i]}")

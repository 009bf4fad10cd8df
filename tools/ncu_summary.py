#!/usr/bin/env python
"""Summarise an ncu --page raw --csv export: python tools/ncu_summary.py prof_raw.csv [pattern...]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
pats = sys.argv[2:] or [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct",
    "launch__registers_per_thread", "launch__occupancy_limit", "launch__waves",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    "sm__inst_executed_pipe_", "sm__pipe_", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg ",
    "smsp__average_warps_issue_stalled", "smsp__warp_issue_stalled", "smsp__average_warp_latency",
    "lts__t_bytes.sum ", "l1tex__throughput", "lts__throughput", "sm__cycles_active.avg",
    "smsp__cycles_active.avg", "clocks", "SM Frequency", "DRAM Frequency",
]
for d in data:
    name_i = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    print("==", d[name_i] if name_i is not None else "")
    for i, k in enumerate(hdr):
        if any(p.strip() in k for p in pats) and d[i] not in ("", "n/a"):
            print(f"  {k:85s} {d[i]:>20s} {units[i]}")

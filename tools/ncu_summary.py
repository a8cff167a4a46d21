#!/usr/bin/env python3
"""Key per-kernel numbers from `ncu -i X.ncu-rep --page raw --csv` (duration, DRAM bytes, pipes, stalls).
Usage: ncu -i rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("==", r[col["Kernel Name"]][:70])
    for k in KEYS:
        if k in col:
            print(f"   {k:75s} {r[col[k]]:>16s} {units[col[k]]}")
    st = sorted(((float(r[i].replace(",", "")), h) for h, i in col.items()
                 if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i].replace(".", "").replace(",", "").isdigit()), reverse=True)[:5]
    print("   stalls/issue:", ", ".join(f"{h.split('issue_stalled_')[1].split('_per')[0]}={v:.2f}" for v, h in st))

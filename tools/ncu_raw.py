#!/usr/bin/env python
"""Key metrics of every kernel in an ncu report: python tools/ncu_raw.py X.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "lts__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum"]
idx = [(h, hdr.index(h)) for h in want if h in hdr]
for r in rows[2:]:
    print("-" * 100)
    for h, i in idx:
        print("%-70s %-12s %s" % (h, units[i], r[i][:90]))

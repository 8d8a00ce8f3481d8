#!/usr/bin/env python
"""Print selected raw metrics of the first kernel in an .ncu-rep: python profiles/ncu_pick.py report.ncu-rep [substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
want = sys.argv[2:] or ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                        "launch__occupancy_limit", "smsp__thread_inst_executed_per_inst_executed.ratio", "_per_issue_active.ratio", "launch__grid_size", "launch__block_size",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                        "smsp__inst_executed_pipe_", "local_load", "local_store"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want) and "per_second" not in h and "elapsed" not in h and not h.endswith((".min", ".max")) and ".max." not in h and ".min." not in h:
        print(h, u, v)

#!/usr/bin/env python
"""Dump the metrics we judge kernels by from an .ncu-rep (read here with `ncu -i`; no GPU needed).
usage: python profiles/ncu_summary.py gpurun_out/prof_shade.ncu-rep > profiles/rNN_ncu_shade.txt"""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
STALLS = ["long_scoreboard", "short_scoreboard", "wait", "no_instruction", "branch_resolving", "math_pipe_throttle", "not_selected", "lg_throttle", "mio_throttle", "barrier", "dispatch_stall", "imc_miss", "selected"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=== launch:", r[hdr.index("Kernel Name")][:70], "| id", r[hdr.index("ID")])
    for w in WANT:
        if w in hdr: print("  %-75s %14s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    print("  -- warps stalled per issue-active cycle (smsp__average_warps_issue_stalled_*_per_issue_active.ratio)")
    for s in STALLS:
        k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
        if k in hdr: print("     %-22s %s" % (s, r[hdr.index(k)]))

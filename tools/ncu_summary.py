#!/usr/bin/env python3
"""Summarise one kernel of an .ncu-rep (ncu --set full) into the metric,unit,value CSV kept under profiles/.
usage: ncu_summary.py report.ncu-rep "<comment line>" > profiles/NAME.csv"""
import csv, io, subprocess, sys
KEEP = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum launch__registers_per_thread launch__grid_size launch__block_size
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__warps_eligible.avg.per_cycle_active smsp__inst_executed.sum
sm__cycles_elapsed.avg sm__cycles_elapsed.avg.per_second
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio""".split()
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print("metric,unit,value")
for c in sys.argv[2:]:
    print("# " + c)
print("# kernel: " + vals[col["Kernel Name"]])
for m in KEEP:
    if m in col:
        print(f"{m},{units[col[m]]},{vals[col[m]]}")

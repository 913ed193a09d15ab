#!/usr/bin/env python3
"""Condenses an .ncu-rep into the handful of metrics the design decisions rest on
(issue-slot utilisation, active lanes per instruction, stall reasons, cache hit rates,
DRAM bytes).  Usage: ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (per SM)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "L1 LSU writeback busy %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 LSU data-stage wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "  of which shared memory"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global load requests (warp level)"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global load sectors"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU pipe warp instructions"),
    ("sm__inst_executed_pipe_xu.sum", "XU pipe warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction (i-cache)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc_miss"),
]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
print("# %s" % rep)
for k, r in enumerate(rows[2:]):
    print("\n## launch %d: %s" % (k, r[col["Kernel Name"]][:90]))
    print("grid %s block %s" % (r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)]))
    for key, label in WANT:
        if key in col:
            print("%-42s %14s %s" % (label, r[col[key]], units[col[key]]))

"""Markdown table of the headline metrics of an `ncu --page raw --csv` export.  usage: python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
ci = {n: i for i, n in enumerate(hdr)}
print("Kernel:", vals[ci["Kernel Name"]] if "Kernel Name" in ci else "?")
print("\n| metric | unit | value |\n|---|---|---|")
for w in WANT:
    if w in ci:
        print(f"| {w} | {units[ci[w]]} | {vals[ci[w]]} |")

#!/usr/bin/env python
"""Key counters of one kernel from an ncu report: python scripts/ncu_summary.py report.ncu-rep [rays]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
rays = float(sys.argv[2]) if len(sys.argv) > 2 else 2073600.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head, units, vals = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
for v in vals:
    d = dict(zip(head, v))
    u = dict(zip(head, units))
    for k in want:
        if k in d:
            print("%-90s %s %s" % (k, d[k], u.get(k, "")))
    try:
        wi = float(d["smsp__inst_executed.sum"].replace(",", "")); ti = float(d["smsp__thread_inst_executed.sum"].replace(",", ""))
        print("warp instructions per ray %.1f, lanes per instruction %.2f" % (wi / rays, ti / wi))
    except Exception as e:
        print("derived: ", e)
    print()

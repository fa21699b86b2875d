#!/usr/bin/env python
"""Append the headline metrics of the stage-kernel launches in an `ncu --page raw --csv` dump to
profiles/stage_kernel_ncu_raw_summary.json under a version tag.   usage: ncu_raw_summary.py raw.csv tag"""
import csv, json, sys
from pathlib import Path
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
out = []
for r in rows[2:]:
    if "stage" not in r[ki]:
        continue
    e = {"kernel": r[ki][:80]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            e[k] = ("%s %s" % (r[i], units[i])).strip()
    out.append(e)
p = Path(__file__).resolve().parents[1] / "profiles" / "stage_kernel_ncu_raw_summary.json"
d = json.loads(p.read_text()) if p.exists() else {}
d[sys.argv[2]] = out
p.write_text(json.dumps(d, indent=1))
for e in out:
    print(sys.argv[2], e.get("gpu__time_duration.sum"), e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum"), e.get("smsp__issue_active.avg.pct_of_peak_sustained_active"), e.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"))

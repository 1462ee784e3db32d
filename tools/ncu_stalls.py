#!/usr/bin/env python3
"""Warp-state (stall) breakdown and pipe utilisation per kernel of an ncu --set full report.
usage: tools/ncu_stalls.py REPORT.ncu-rep [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0]
    if filt not in name:
        continue
    print(f"## {name}")
    vals = []
    for i, m in enumerate(hdr):
        if "smsp__average_warp" in m and "per_issue_active" in m or m.startswith("smsp__average_warps_issue_stalled") :
            try:
                vals.append((float(r[i].replace(",", "")), m))
            except ValueError:
                pass
    for v, m in sorted(vals, reverse=True)[:12]:
        print(f"  {v:8.3f}  {m}")
    for key in ("gpu__time_duration.sum", "sm__inst_executed_pipe_fmaheavy", "sm__inst_executed_pipe_fmalite", "sm__pipe_fmaheavy", "sm__pipe_fmalite",
                "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles", "sm__inst_executed_pipe_alu", "sm__pipe_alu", "sm__inst_executed_pipe_lsu",
                "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit",
                "launch__registers_per_thread", "sm__maximum_warps_per_active_cycle_pct", "smsp__thread_inst_executed_per_inst_executed",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
                "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_active.avg "):
        for i, m in enumerate(hdr):
            if m.startswith(key):
                print(f"  {m} = {r[i]} {units[i]}")
    print()

#!/usr/bin/env python3
"""Summarise an ncu report (--set full) and an ncu launch list (gpu__time_duration.sum CSV) as markdown.

usage: tools/ncu_summary.py REPORT.ncu-rep [LAUNCHES.csv] > profiles/rNN_ncu_summary.md
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__inst_executed_op_global_red.sum", "lts__t_requests_srcunit_tex_op_red.sum",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {r[hdr.index('Kernel Name')].split('(')[0]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"| {m} | {r[i]} | {units[i]} |")
        print()
    if len(sys.argv) > 2:
        lr = [r for r in csv.reader(open(sys.argv[2])) if len(r) > 5]
        h = lr[0]
        ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
        agg = collections.OrderedDict()
        for r in lr[1:]:
            v = float(r[iv].replace(",", ""))
            v = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
            a = agg.setdefault(r[ik].split("(")[0][:70], [0, 0.0])
            a[0] += 1
            a[1] += v
        print("## Launch list\n\n| kernel | launches | total ms | avg ms |\n|---|---|---|---|")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            print(f"| {k} | {n} | {t:.2f} | {t / n:.3f} |")


if __name__ == "__main__":
    main()

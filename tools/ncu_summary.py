#!/usr/bin/env python
"""Summarise ncu outputs into the small text files committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv          # per-kernel time + share
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep [kernel-regex]  # key counters of a --set full capture
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
    "lts__t_bytes.sum", "lts__t_bytes.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__thread_inst_executed_per_inst_executed.pct", "smsp__thread_inst_executed_pred_on_per_inst_executed.ratio",
    "smsp__inst_issued.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "local_load", "derived__memory_l2_theoretical_sectors_global_excessive",
]


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    total = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(name, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1; a[1] += us; total += us
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {total / 1e3:.3f} ms total (ncu-serialised, cold cache: compare shares)")
    print(f"{'kernel':70s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}  grid block")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {a[0]:5d} {a[1]:12.1f} {a[1] / a[0]:10.1f} {100 * a[1] / total:6.1f}%  {a[2]} {a[3]}")


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr = i
            break
    names, units = rows[hdr], rows[hdr + 1]
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        if pattern and not re.search(pattern, d["Kernel Name"]):
            continue
        print(f"## {d['Kernel Name'][:120]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            for n, u in zip(names, units):
                if n == k or (k in ("local_load",) and k in n):
                    print(f"  {n:90s} {d[n]:>18s} {u}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

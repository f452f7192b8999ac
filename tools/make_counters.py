#!/usr/bin/env python
"""profiles/counters.json + profiles/traffic.json + the per-kernel text summaries from the ncu captures of tools/r02_profiles.sh.

    python tools/make_counters.py [gpurun_out]        # reads gpurun_out/r02_*.ncu-rep, rewrites the numeric fields; the prose stays
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out")
# config key -> (capture, summary file under profiles/, source-lines file or None)
CAPTURES = {
    "c2_phantom": ("r02_c2_pool.ncu-rep", "r02_trace_pool_kernel_c2_ncu_full.txt", "r02_trace_pool_kernel_c2_source_lines.txt"),
    "c5_phantom": ("r02_c5_pool.ncu-rep", "r02_trace_pool_kernel_c5_ncu_full.txt", "r02_trace_pool_kernel_c5_source_lines.txt"),
    "c3_lss": ("r02_c3_lss.ncu-rep", "r02_trace_kernel_lss_c3_ncu_full.txt", "r02_trace_kernel_lss_c3_source_lines.txt"),
    "c4_dots": ("r02_c4_dots.ncu-rep", "r02_trace_kernel_dots_c4_ncu_full.txt", "r02_trace_kernel_dots_c4_source_lines.txt"),
    "c1_phantom": ("r02_c1_lane.ncu-rep", "r02_trace_kernel_lane_bound_c1_ncu_full.txt", None),
}


def raw_row(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            d[h] = (float(v.replace(",", "")), u)
        except ValueError:
            d[h] = (v, u)
    return d


def to_ms(v, u):
    return v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else (v if u in ("ms", "msecond") else v * 1e3))


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def main():
    cpath = os.path.join(ROOT, "profiles", "counters.json")
    counters = json.load(open(cpath))
    for key, (rep, summary, lines) in CAPTURES.items():
        path = os.path.join(SRC, rep)
        if not os.path.exists(path):
            print("missing", path)
            continue
        d = raw_row(path)
        c = counters.setdefault(key, {})
        c["kernel_ms_under_ncu"] = to_ms(*d["gpu__time_duration.sum"])
        c["dram_bytes_per_launch"] = to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
        c["issue_active"] = d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0] / 100.0
        c["lanes_per_inst"] = round(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0], 2)
        c["regs"] = int(d["launch__registers_per_thread"][0])
        c["warps_active"] = d["sm__warps_active.avg.pct_of_peak_sustained_active"][0] / 100.0
        c["l2_hit_rate"] = d["lts__t_sector_hit_rate.pct"][0] / 100.0
        c["l1_hit_rate"] = d["l1tex__t_sector_hit_rate.pct"][0] / 100.0
        c["warp_instructions"] = d["smsp__inst_executed.sum"][0]
        with open(os.path.join(ROOT, "profiles", summary), "w") as f:
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "raw", path], capture_output=True, text=True).stdout)
        if lines:
            with open(os.path.join(ROOT, "profiles", lines), "w") as f:
                f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), path, "60"], capture_output=True, text=True).stdout)
        print(key, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in c.items() if k not in ("binding", "source", "kernel")})
    json.dump(counters, open(cpath, "w"), indent=1)
    traffic = {k + "_bytes_per_launch": v["dram_bytes_per_launch"] for k, v in counters.items()}
    traffic["source"] = "profiles/counters.json (round 2; dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, one ncu --set full capture per workload)"
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    lcsv = os.path.join(SRC, "r02_launches.csv")
    if os.path.exists(lcsv):
        with open(os.path.join(ROOT, "profiles", "r02_launches.txt"), "w") as f:
            f.write(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches", lcsv], capture_output=True, text=True).stdout)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries kept under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv          > profiles/r1_launches.txt
  python profiles/summarize.py kernels  gpurun_out/prof.ncu-rep          > profiles/r1_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__registers_per_thread",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(r[ui], 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; source: {path}")
    print(f"# {'kernel':58s} {'launches':>8s} {'total_us':>12s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {v[0]:8d} {v[1]:12.1f} {v[1] / tot:7.3f}")


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on ; source: {path}")
    for r in rows[2:]:
        print("\n== " + r[hdr.index("Kernel Name")][:110])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"   {m:86s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](sys.argv[2])

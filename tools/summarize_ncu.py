#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few lines the roofline argument needs.

    python tools/summarize_ncu.py gpurun_out/prof_stream.ncu-rep [--json profiles/ncu_traffic.json --elems N]
"""
import argparse
import csv
import hashlib
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL_SOURCES = ("fabe13_kernels.cu", "fabe13_core.h", "fabe13_constants.h", "fabe13_ph_remainders.h", "fabe13_internal.h")  # = bench.py's list


def kernel_source_hash():
    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "fabe_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()

KEYS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_elapsed.avg.per_second",
    "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--json")
    ap.add_argument("--elems", type=int, default=0)
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("report:", args.report)
    for n, r in enumerate(data):
        print(f"launch {n}: {r[ki][:150]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:84s} {units[i]:16s} " + " | ".join(r[i][:24] for r in data))
    if args.json and args.elems:
        i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = [float(r[i_r]) * scale[units[i_r]] + float(r[i_w]) * scale[units[i_w]] for r in data]
        per = sum(tot) / len(tot) / args.elems
        with open(args.json, "w") as f:
            # stamped with the kernel sources the capture was taken on (run this right after the GPU call, before
            # touching them): bench.py carries the figure into roofline.traffic only while the stamp matches
            json.dump({"dram_bytes_per_elem": per, "elements_per_launch": args.elems, "launches": len(tot),
                       "report": os.path.basename(args.report), "kernel_source_sha256": kernel_source_hash(),
                       "metric": "dram__bytes_read.sum + dram__bytes_write.sum (ncu --set full)"}, f, indent=1)
        print(f"dram bytes per element: {per:.3f} (algorithmic 24)")


if __name__ == "__main__":
    main()

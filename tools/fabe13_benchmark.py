#!/usr/bin/env python
"""The reference's benchmark table (fabe13/benchmark_fabe13.c:153-155 header, :169 size ladder, :192 input range,
:265-268 row format, :279-299 accuracy check) with GPU columns appended, so scripts that parse the reference's stdout
keep working: the first eight columns keep their meaning and order; device-resident and pinned-host columns follow.

    python tools/fabe13_benchmark.py [--max-n 1000000000] [--reps 5]

Column 2 ("FABE13 (sec)") is the reference's own call: fabe13_sincos(in, sin, cos, n) on ordinary pageable host
arrays, wall clock.  "Libm" is numpy's sin+cos over the same array (one thread), as the reference's libm loop is.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def fmt_n(n):
    return f"{n:,}"


def best_of(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-n", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--libm-max-n", type=int, default=100_000_000, help="skip the libm column above this size")
    args = ap.parse_args()
    import torch

    dev = torch.device("cuda:0")
    print("\n=== FABE13 Benchmark (B200 build) ===")
    print(f"FABE13 Active Implementation: {fb.implementation_name()} (SIMD Width: {fb.simd_width()})")
    print("Benchmarking sincos(x) for various array sizes (N):")
    line = "-" * 191
    print(line)
    print("%-15s | %-15s | %-15s | %-15s | %-15s | %-10s | %-15s | %-15s | %-17s | %-17s | %-17s" % (
        "N", "FABE13 (sec)", "Libm (sec)", "FABE13 (M/sec)", "Libm (M/sec)", "Speedup", "FABE13 CPU(%)", "Libm CPU(%)",
        "Device (M/sec)", "Pinned (M/sec)", "Device GB/s@24B"))
    print(line)
    rng = np.random.default_rng(13)
    n = 10
    worst = {}
    while n <= args.max_n:
        x = (2.0 * rng.random(n) - 1.0) * 100000.0 * np.pi  # benchmark_fabe13.c:192
        s = np.empty_like(x)
        c = np.empty_like(x)
        fb.sincos(x[: min(n, 1000)].copy())  # warm-up as the reference does (:195-199)
        cpu0 = time.process_time()
        w0 = time.perf_counter()
        t_f = best_of(lambda: fb.sincos(x, out_sin=s, out_cos=c), args.reps)
        cpu_pct = 100.0 * (time.process_time() - cpu0) / max(time.perf_counter() - w0, 1e-12)
        if n <= args.libm_max_n:
            cpu0 = time.process_time()
            w0 = time.perf_counter()
            ls = np.empty_like(x)
            lc = np.empty_like(x)
            t_l = best_of(lambda: (np.sin(x, out=ls), np.cos(x, out=lc)), max(1, args.reps if n <= 10_000_000 else 1))
            lcpu = 100.0 * (time.process_time() - cpu0) / max(time.perf_counter() - w0, 1e-12)
            if n <= 1_000_000:  # accuracy check as benchmark_fabe13.c:279-296
                ok = np.isfinite(ls) & np.isfinite(s)
                worst[n] = (float(np.abs(s - ls)[ok].max()), float(np.abs(c - lc)[ok].max()))
            libm_cols = ("%.6f" % t_l, "%.2f" % (n / t_l / 1e6), "%.2fx" % (t_l / t_f), "%.1f" % lcpu)
        else:
            libm_cols = ("skipped", "-", "-", "-")
        # device-resident
        xd = torch.from_numpy(x).to(dev)
        sd = torch.empty_like(xd)
        cd = torch.empty_like(xd)
        for _ in range(3):
            fb.sincos(xd, out_sin=sd, out_cos=cd)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        reps_d = max(args.reps, 5)
        e0.record()
        for _ in range(reps_d):
            fb.sincos(xd, out_sin=sd, out_cos=cd)
        e1.record()
        torch.cuda.synchronize()
        t_d = e0.elapsed_time(e1) * 1e-3 / reps_d
        del xd, sd, cd
        # pinned host
        px, ps, pc = fb.PinnedArray(n), fb.PinnedArray(n), fb.PinnedArray(n)
        px.array[:] = x
        fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
        t_p = best_of(lambda: fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array), args.reps)
        for p in (px, ps, pc):
            p.free()
        print("%-15s | %-15.6f | %-15s | %-15.2f | %-15s | %-10s | %-15.1f | %-15s | %-17.2f | %-17.2f | %-17.1f" % (
            fmt_n(n), t_f, libm_cols[0], n / t_f / 1e6, libm_cols[1], libm_cols[2], cpu_pct, libm_cols[3],
            n / t_d / 1e6, n / t_p / 1e6, 24.0 * n / t_d / 1e9), flush=True)
        n *= 10
    print(line)
    print("\nAccuracy (max abs diff vs numpy sin/cos, inputs in +-1e5*pi, N <= 1e6):")
    for k in sorted(worst):
        print("  N=%-10s  sin %.3e   cos %.3e" % (fmt_n(k), worst[k][0], worst[k][1]))
    print("\nReference README (NEON, same input range): 1.224e-11 / 1.225e-11.")


if __name__ == "__main__":
    main()

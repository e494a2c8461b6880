#!/usr/bin/env python
"""Option host_assist on one GPU: a page-locked N-element array through fabe13_sincos_host with k host threads helping
(k = 0: the GPU alone), wall clock, and how the array was split.    python tools/host_assist_sweep.py [N]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
    threads = len(os.sched_getaffinity(0))
    px, ps, pc = fb.PinnedArray(n), fb.PinnedArray(n), fb.PinnedArray(n)
    x = torch.empty(n, dtype=torch.float64, device="cuda:0")
    fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
    torch.from_numpy(px.array).copy_(x)
    del x
    fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
    want = (ps.array[:4096].copy(), ps.array[-4096:].copy())
    print(f"# host_assist sweep, N = {n:.3g} pinned, {threads} host threads, {fb.implementation_name()}")
    print("| host_assist | Gelem/s | GPU share | GPU Gelem/s | host Gelem/s |")
    print("|---|---|---|---|---|")
    for k in [0, 2, 4, 8, 12, threads - 2, threads - 1, threads]:
        fb.set_option("host_assist", k)
        fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
        g0, h0 = fb.get_option("assist_gpu_elems"), fb.get_option("assist_host_elems")
        ps.array[:4096] = 0
        ps.array[-4096:] = 0
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
        dt = time.perf_counter() - t0
        g, h = fb.get_option("assist_gpu_elems") - g0, fb.get_option("assist_host_elems") - h0
        ok = np.array_equal(ps.array[:4096], want[0]) and np.array_equal(ps.array[-4096:], want[1])
        share = g / (g + h) if k else 1.0
        print(f"| {k} | {n * reps / dt / 1e9:.2f} | {share:.3f} | {(g if k else n * reps) / dt / 1e9:.2f} | {h / dt / 1e9:.2f} |" + ("" if ok else " RESULTS DIFFER"), flush=True)
    fb.set_option("host_assist", 0)
    for p in (px, ps, pc):
        p.free()


if __name__ == "__main__":
    main()

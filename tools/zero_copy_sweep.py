#!/usr/bin/env python
"""Pinned host arrays: copy-engine pipeline vs one zero-copy kernel reading/writing host memory over PCIe, by n."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fabe_b200 as fb
nmax = 100_000_000
px, ps, pc = fb.PinnedArray(nmax), fb.PinnedArray(nmax), fb.PinnedArray(nmax)
px.array[:] = np.linspace(-1e4, 1e4, nmax)
print(f"{'n':>10s} {'pipeline us':>12s} {'zero-copy us':>13s}   Gelem/s pipeline / zero-copy")
for n in (10_000, 32_768, 100_000, 300_000, 1_000_000, 3_000_000, 10_000_000, 30_000_000, 100_000_000):
    res = []
    for zc in (0, 1 << 30):
        fb.set_option("zero_copy_max", zc)
        x, s, c = px.array[:n], ps.array[:n], pc.array[:n]
        fb.sincos(x, out_sin=s, out_cos=c)
        ts = []
        for _ in range(7 if n <= 10_000_000 else 3):
            t0 = time.perf_counter(); fb.sincos(x, out_sin=s, out_cos=c); ts.append(time.perf_counter() - t0)
        res.append(min(ts))
    print(f"{n:10d} {res[0] * 1e6:12.1f} {res[1] * 1e6:13.1f}   {n / res[0] / 1e9:7.3f} / {n / res[1] / 1e9:7.3f}", flush=True)

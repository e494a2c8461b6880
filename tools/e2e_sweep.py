#!/usr/bin/env python
"""Host-resident (pinned) end-to-end throughput vs pipeline chunk size."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fabe_b200 as fb
n = 250_000_000
px, ps, pc = fb.PinnedArray(n), fb.PinnedArray(n), fb.PinnedArray(n)
px.array[:] = np.linspace(-1e4, 1e4, n)
for chunk in (1 << 19, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24, 1 << 25):
    fb.set_option("chunk_elems", chunk)
    fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array); ts.append(time.perf_counter() - t0)
    print(f"chunk_elems={chunk:>9d}: best {n / min(ts) / 1e9:.3f} Gelem/s  avg {n / (sum(ts) / len(ts)) / 1e9:.3f}  ({16 * n / min(ts) / 1e9:.1f} GB/s D2H)", flush=True)
for name, fn in (("sin only", fb.sin_array), ("cos only", fb.cos_array)):
    fb.set_option("chunk_elems", 1 << 22)
    fn(px.array, out=ps.array)
    t0 = time.perf_counter(); fn(px.array, out=ps.array); t = time.perf_counter() - t0
    print(f"{name}: {n / t / 1e9:.3f} Gelem/s", flush=True)

#!/usr/bin/env python
"""Host-resident (pinned) end-to-end throughput of fabe13_sincos_host vs array size and pipeline chunk size, one GPU.
   python tools/e2e_sweep.py [sizes...]      (default 250e6 500e6 1e9)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fabe_b200 as fb

sizes = [int(float(a)) for a in sys.argv[1:]] or [250_000_000, 500_000_000, 1_000_000_000]
nmax = max(sizes)
px, ps, pc = fb.PinnedArray(nmax), fb.PinnedArray(nmax), fb.PinnedArray(nmax)
px.array[:] = np.linspace(-1e4, 1e4, nmax)
for n in sizes:
    x, s, c = px.array[:n], ps.array[:n], pc.array[:n]
    for chunk in (1 << 21, 1 << 22, 1 << 23, 1 << 24):
        fb.set_option("chunk_elems", chunk)
        fb.sincos(x, out_sin=s, out_cos=c)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); fb.sincos(x, out_sin=s, out_cos=c); ts.append(time.perf_counter() - t0)
        print(f"n={n:>10d} chunk_elems={chunk:>9d}: best {n / min(ts) / 1e9:.3f} Gelem/s  avg {n / (sum(ts) / len(ts)) / 1e9:.3f}  "
              f"({16 * n / min(ts) / 1e9:.1f} GB/s D2H)", flush=True)
fb.set_option("chunk_elems", 1 << 22)
n = sizes[0]
for name, fn in (("sin only", fb.sin_array), ("cos only", fb.cos_array)):
    fn(px.array[:n], out=ps.array[:n])
    t0 = time.perf_counter(); fn(px.array[:n], out=ps.array[:n]); t = time.perf_counter() - t0
    print(f"{name} (n={n}): {n / t / 1e9:.3f} Gelem/s", flush=True)

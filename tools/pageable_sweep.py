#!/usr/bin/env python
"""Pageable (malloc'ed) host arrays through fabe13_sincos_host: wall-clock throughput vs staging threads, next to the
same arrays page-locked once with fabe13_cuda_host_register, and the host's plain memcpy rate for scale."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402

n = 100_000_000
x = np.random.default_rng(1).uniform(-1e4, 1e4, n)
s = np.zeros(n)
c = np.zeros(n)
t0 = time.perf_counter()
s[:] = x
t = time.perf_counter() - t0
print(f"host memcpy, 1 thread: {8 * n / t / 1e9:.1f} GB/s read + the same written")
fb.sincos(x[:1_000_000], out_sin=s[:1_000_000], out_cos=c[:1_000_000])
for threads, nt in ((8, 1), (8, 0), (0, 1), (0, 0), (16, 1), (16, 0), (8, 1), (8, 0), (0, 1), (0, 0)):
    fb.set_option("stage_threads", threads)
    fb.set_option("stream_copy", nt)
    fb.sincos(x, out_sin=s, out_cos=c)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        fb.sincos(x, out_sin=s, out_cos=c)
        ts.append(time.perf_counter() - t0)
    print(f"pageable, stage_threads={threads}, stream_copy={nt}: best {n / min(ts) / 1e9:.3f} Gelem/s  avg {n / (sum(ts) / len(ts)) / 1e9:.3f}", flush=True)
lib = fb._lib.load()
t0 = time.perf_counter()
for a in (x, s, c):
    assert lib.fabe13_cuda_host_register(a.ctypes.data, a.nbytes) == 0
t_reg = time.perf_counter() - t0
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    fb.sincos(x, out_sin=s, out_cos=c)
    ts.append(time.perf_counter() - t0)
print(f"same arrays after fabe13_cuda_host_register ({t_reg * 1e3:.0f} ms for {24 * n / 1e9:.1f} GB, once): best {n / min(ts) / 1e9:.3f} Gelem/s")
for a in (x, s, c):
    lib.fabe13_cuda_host_unregister(a.ctypes.data)

#!/usr/bin/env python
"""Short A/B run for kernel-variant builds (FABE13_LIBRARY=<variant .so> python tools/ab_quick.py <tag> [worklist])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from tools.sweep import report, time_steps  # noqa: E402

dev = torch.device("cuda:0")
tag = sys.argv[1] if len(sys.argv) > 1 else ""
if len(sys.argv) > 2:
    fb.set_option("worklist", int(sys.argv[2]))
n = 1_000_000_000
m = 100_000_000
x = torch.empty(n, dtype=torch.float64, device=dev)
s = torch.empty_like(x)
c = torch.empty_like(x)
fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
for rep in range(2):
    avg, best = time_steps(x, s, c, steps=20)
    report(f"{tag} c3 1e9 burst(20)", n, avg, best)
avg, best = time_steps(x, s, c, steps=300)
report(f"{tag} c3 1e9 sustained(300)", n, avg, best)
avg, best = time_steps(x[:m], s[:m], c[:m], steps=20)
report(f"{tag} c2 1e8", m, avg, best)
for k in (1_000_000, 4_000_000, 10_000_000):
    fb.set_option("small_n", 0)
    avg, best = time_steps(x[:k], s[:k], c[:k], steps=50)
    report(f"{tag} small_n=0 n={k}", k, avg, best)
fb.set_option("small_n", 1 << 22)
for pct in (0.01, 0.1, 1, 5, 25):
    fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
    step = max(1, int(round(100 / pct)))
    h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
    fb.fill_loguniform_device(h, 0xFABE1304)
    x[:m][::step] = h
    avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
    report(f"{tag} mix {100.0 / step:.2f}% extreme n=1e8", m, avg, best)
torch.manual_seed(1)
for pct in (0.1, 1, 3):  # randomly placed (Poisson counts per warp) instead of every k-th element
    fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
    mask = torch.rand(m, device=dev) < pct / 100.0
    h = torch.empty(int(mask.sum().item()), dtype=torch.float64, device=dev)
    fb.fill_loguniform_device(h, 0xFABE1304)
    x[:m][mask] = h
    del mask
    avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
    report(f"{tag} random mix {pct}% extreme n=1e8", m, avg, best)
fb.fill_loguniform_device(x[:m], 0xFABE1304)
avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
report(f"{tag} c4 dense loguniform 1e8", m, avg, best)
fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
x[:m][::7] = 0.0
avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
report(f"{tag} 1 zero in 7 n=1e8", m, avg, best)

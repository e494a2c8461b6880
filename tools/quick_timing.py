#!/usr/bin/env python
"""Quick A/B timing: c3 burst + sustained, sparse mixes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb
from tools.sweep import report, time_steps
dev = torch.device("cuda:0")
n = 1_000_000_000
x = torch.empty(n, dtype=torch.float64, device=dev); s = torch.empty_like(x); c = torch.empty_like(x)
fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
avg, best = time_steps(x, s, c, steps=10); report(f"{tag} c3 1e9 burst(10)", n, avg, best)
avg, best = time_steps(x, s, c, steps=300); report(f"{tag} c3 1e9 sustained(300)", n, avg, best)
m = 100_000_000
for pct in (0.1, 1, 5):
    step = int(round(100 / pct))
    h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
    fb.fill_loguniform_device(h, 0xFABE1304)
    x[:m][::step] = h
    avg, best = time_steps(x[:m], s[:m], c[:m], steps=5); report(f"{tag} mix {pct}% extreme n=1e8", m, avg, best)
    fb.fill_uniform_device(x[:m], 0xFABE1303, -1e4, 1e4)

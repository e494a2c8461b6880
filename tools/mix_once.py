#!/usr/bin/env python
"""One sparse-mix run (for ncu launch lists): n=1e8, 1 lane in STEP in the Payne-Hanek range."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb
step = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda:0")
m = 100_000_000
x = torch.empty(m, dtype=torch.float64, device=dev); s = torch.empty_like(x); c = torch.empty_like(x)
fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
fb.fill_loguniform_device(h, 0xFABE1304)
x[::step] = h
for _ in range(4):
    fb.sincos(x, out_sin=s, out_cos=c)
torch.cuda.synchronize()
print("ok", float(s[:10].abs().sum()))

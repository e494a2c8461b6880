#!/usr/bin/env python
"""C4 dense + one sparse mix, for A/B builds of the second pass."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb
from tools.sweep import report, time_steps
tag = sys.argv[1] if len(sys.argv) > 1 else ""
dev = torch.device("cuda:0")
m = 100_000_000
x = torch.empty(m, dtype=torch.float64, device=dev); s = torch.empty_like(x); c = torch.empty_like(x)
fb.fill_loguniform_device(x, 0xFABE1304)
for sp in (4, 8, 16):
    fb.set_option("second_pass_blocks_per_sm", sp)
    avg, best = time_steps(x, s, c, steps=5); report(f"{tag} c4 dense second_pass_bps={sp}", m, avg, best)
fb.set_option("second_pass_blocks_per_sm", 8)
fb.fill_uniform_device(x, 0xFABE1302, -1e4, 1e4)
h = torch.empty(m // 20, dtype=torch.float64, device=dev); fb.fill_loguniform_device(h, 0xFABE1304); x[::20] = h
avg, best = time_steps(x, s, c, steps=5); report(f"{tag} c4 mix 5%", m, avg, best)

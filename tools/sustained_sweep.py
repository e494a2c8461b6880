#!/usr/bin/env python
"""Burst (10 steps) vs sustained (300 steps, power-capped) throughput for a few launch shapes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb
from tools.sweep import report, time_steps
dev = torch.device("cuda:0")
n = 1_000_000_000
x = torch.empty(n, dtype=torch.float64, device=dev); s = torch.empty_like(x); c = torch.empty_like(x)
fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
for unroll, tpb, bps in ((1, 1, 0), (1, 2, 0), (1, 4, 0), (2, 1, 0), (2, 4, 0), (1, 1, 32), (1, 1, 1)):
    fb.set_option("unroll", unroll); fb.set_option("tiles_per_block", tpb); fb.set_option("blocks_per_sm", bps)
    avg, best = time_steps(x, s, c, steps=10); report(f"unroll={unroll} tiles_per_block={tpb} blocks_per_sm={bps} burst(10)", n, avg, best)
    avg, best = time_steps(x, s, c, steps=300); report(f"unroll={unroll} tiles_per_block={tpb} blocks_per_sm={bps} sustained(300)", n, avg, best)

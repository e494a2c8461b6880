#!/usr/bin/env python
"""Timing of the Payne-Hanek workloads (BASELINE config 4) on one GPU: dense log-uniform and sparse mixes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from tools.sweep import report, time_steps  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    m = 100_000_000
    x = torch.empty(m, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    fb.fill_loguniform_device(x, 0xFABE1304)
    for core in (0, 1):
        fb.set_option("core", core)
        for shift in (0, 3, 20):
            fb.set_option("worklist_shift", shift)
            for sp in (4, 8, 16):
                fb.set_option("second_pass_blocks_per_sm", sp)
                avg, best = time_steps(x, s, c, steps=5)
                report(f"c4 dense loguniform n={m} core={core} worklist_shift={shift} second_pass_bps={sp}", m, avg, best)
    fb.set_option("core", 0)
    fb.set_option("worklist_shift", 3)
    fb.set_option("second_pass_blocks_per_sm", 8)
    for pct in (0.01, 0.1, 1, 5, 12, 25, 50):
        fb.fill_uniform_device(x, 0xFABE1302, -1e4, 1e4)
        step = max(1, int(round(100 / pct)))
        h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
        fb.fill_loguniform_device(h, 0xFABE1304)
        x[::step] = h
        avg, best = time_steps(x, s, c, steps=5)
        report(f"c4 mix: 1 lane in {step} extreme ({100.0 / step:.2f}%) n={m}", m, avg, best)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B of the Payne-Hanek worklist designs on one GPU, alternating in one process (option worklist: 0 = by size,
1 = global list + second-pass kernel [the round-1 streaming kernel], 2 = warp-local lists, 3 = hybrid):
    python tools/ab_worklist.py <tag> [modes, default 0,1]  C3 burst and sustained, C2, C4 dense, sparse mixes, and
per-call GPU time for small/medium n."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from tools.sweep import report, time_steps  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    tag = sys.argv[1] if len(sys.argv) > 1 else ""
    n = 1_000_000_000
    x = torch.empty(n, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    m = 100_000_000
    pair = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (0, 1)
    modes = pair + pair
    fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
    for wl in modes:
        fb.set_option("worklist", wl)
        avg, best = time_steps(x, s, c, steps=20)
        report(f"{tag} worklist={wl} c3 1e9 burst(20)", n, avg, best)
    for wl in pair:
        fb.set_option("worklist", wl)
        avg, best = time_steps(x, s, c, steps=300)
        report(f"{tag} worklist={wl} c3 1e9 sustained(300)", n, avg, best)
    for wl in modes:
        fb.set_option("worklist", wl)
        avg, best = time_steps(x[:m], s[:m], c[:m], steps=20)
        report(f"{tag} worklist={wl} c2 1e8", m, avg, best)
    # small / medium calls: GPU time per call (back to back on one stream)
    for k in (10_000, 100_000, 1_000_000, 4_000_000, 10_000_000):
        for wl in pair:
            for small_n in (0, 1 << 22):
                fb.set_option("worklist", wl)
                fb.set_option("small_n", small_n)
                avg, best = time_steps(x[:k], s[:k], c[:k], steps=50)
                report(f"{tag} worklist={wl} small_n={small_n} n={k}", k, avg, best)
    fb.set_option("small_n", 1 << 22)
    fb.fill_loguniform_device(x[:m], 0xFABE1304)
    for wl in modes:
        fb.set_option("worklist", wl)
        avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
        report(f"{tag} worklist={wl} c4 dense loguniform 1e8", m, avg, best)
    for pct in (0.01, 0.1, 1, 5, 12, 25, 50):
        fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
        step = max(1, int(round(100 / pct)))
        h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
        fb.fill_loguniform_device(h, 0xFABE1304)
        x[:m][::step] = h
        for wl in pair:
            fb.set_option("worklist", wl)
            avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
            report(f"{tag} worklist={wl} mix {100.0 / step:.2f}% extreme n=1e8", m, avg, best)
    # zeros in the data (tiny lanes: not plain, but no Payne-Hanek work): the block vote fires on every tile
    fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
    x[:m][::7] = 0.0
    for wl in pair:
        fb.set_option("worklist", wl)
        avg, best = time_steps(x[:m], s[:m], c[:m], steps=10)
        report(f"{tag} worklist={wl} 1 zero in 7 n=1e8", m, avg, best)


if __name__ == "__main__":
    main()

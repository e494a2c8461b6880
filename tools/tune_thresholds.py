#!/usr/bin/env python
"""The two thresholds of the fused kernel's Payne-Hanek routing, swept in one process on mixes of plain and huge
arguments (n = 1e8, hybrid sequence forced): option local_min (a warp with at least this many huge lanes of its 128
reduces them itself, fewer go to the global worklist) and option dense_hint (first-element votes, of 32, that send a
warp down the dense route)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from tools.sweep import time_steps  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    m = 100_000_000
    x = torch.empty(m, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    fb.set_option("worklist", 3)
    torch.manual_seed(1)
    mixes = []
    for label, make in (("periodic 1%", lambda: 100), ("periodic 2%", lambda: 50), ("periodic 3.3%", lambda: 30),
                        ("random 1%", None), ("random 3%", None), ("random 10%", None), ("random 40%", None), ("random 70%", None)):
        mixes.append(label)
    print("| mix | " + " | ".join(f"local_min={v}" for v in (2, 3, 4, 6, 8)) + " | " + " | ".join(f"dense_hint={v}" for v in (12, 16, 20, 24, 28)) + " |")
    print("|---" * 11 + "|")
    for label in mixes:
        fb.fill_uniform_device(x, 0xFABE1302, -1e4, 1e4)
        if label.startswith("periodic"):
            step = {"periodic 1%": 100, "periodic 2%": 50, "periodic 3.3%": 30}[label]
            h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[::step] = h
        else:
            pct = float(label.split()[1].rstrip("%"))
            mask = torch.rand(m, device=dev) < pct / 100.0
            h = torch.empty(int(mask.sum().item()), dtype=torch.float64, device=dev)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[mask] = h
            del mask
        row = []
        fb.set_option("dense_hint", 16)
        for v in (2, 3, 4, 6, 8):
            fb.set_option("local_min", v)
            avg, _ = time_steps(x, s, c, steps=8)
            row.append(f"{m / avg / 1e6:.1f}")
        fb.set_option("local_min", 2)
        for v in (12, 16, 20, 24, 28):
            fb.set_option("dense_hint", v)
            avg, _ = time_steps(x, s, c, steps=8)
            row.append(f"{m / avg / 1e6:.1f}")
        print(f"| {label} | " + " | ".join(row) + " |", flush=True)


if __name__ == "__main__":
    main()

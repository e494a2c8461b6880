#!/usr/bin/env python
"""Option dense_hint (first-element votes, of 32, that send a warp's tile down the dense route) swept on random mixes of
plain and huge arguments with the default one-launch design, n = 1e8; Gelem/s, median of 5 groups of 8 calls."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from tools.tune_stragglers import rate  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    m = 100_000_000
    x = torch.empty(m, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    torch.manual_seed(1)
    hints = (4, 8, 10, 12, 14, 16, 20, 33)
    print("| mix | " + " | ".join(f"dense_hint={v}" + (" (never)" if v == 33 else "") for v in hints) + " |")
    print("|---" * (len(hints) + 1) + "|")
    for pct in (5, 10, 20, 30, 40, 50, 60, 80):
        fb.fill_uniform_device(x, 0xFABE1302, -1e4, 1e4)
        mask = torch.rand(m, device=dev) < pct / 100.0
        h = torch.empty(int(mask.sum().item()), dtype=torch.float64, device=dev)
        fb.fill_loguniform_device(h, 0xFABE1304)
        x[mask] = h
        del mask, h
        row = []
        for v in hints:
            fb.set_option("dense_hint", v)
            row.append(f"{rate(x, s, c):.1f}")
        print(f"| random {pct}% | " + " | ".join(row) + " |", flush=True)
    fb.set_option("dense_hint", 16)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where should a warp's few Payne-Hanek lanes be reduced?  Sweeps option direct_max (each lane reduces its own, up to
this many per warp) under the one-launch design (worklist=2) and the hybrid sequence (worklist=3, lone lanes to the
global list + second pass) on sparse mixes at n = 1e8, and measures the fixed cost of the hybrid sequence and of
programmatic dependent launch on plain data at the 8-GPU slice size (n = 1.25e8) and at n = 1e9.
Kernels alternate inside one process; Gelem/s, median of `reps` groups of `steps` back-to-back calls."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def rate(x, s, c, steps=8, reps=5):
    fb.sincos(x, out_sin=s, out_cos=c)
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fb.sincos(x, out_sin=s, out_cos=c)
        e1.record()
        torch.cuda.synchronize()
        out.append(x.numel() * steps / e0.elapsed_time(e1) / 1e6)
    out.sort()
    return out[len(out) // 2]


def main():
    dev = torch.device("cuda:0")
    m = 100_000_000
    big = torch.empty(1_000_000_000, dtype=torch.float64, device=dev)
    s = torch.empty_like(big)
    c = torch.empty_like(big)
    torch.manual_seed(1)
    cols = [(2, 0), (2, 1), (2, 2), (2, 4), (2, 8), (2, 16), (3, 0), (3, 4), (3, 8)]
    print("# sparse mixes, n = 1e8: Gelem/s by (worklist, direct_max); worklist 2 = one launch, 3 = hybrid sequence")
    print("| mix | " + " | ".join(f"wl={w} dm={d}" for w, d in cols) + " |")
    print("|---" * (len(cols) + 1) + "|")
    x = big[:m]
    for label in ("random 0.01%", "random 0.1%", "periodic 1%", "random 1%", "random 2%", "random 3%", "random 5%", "random 10%"):
        fb.fill_uniform_device(x, 0xFABE1302, -1e4, 1e4)
        if label.startswith("periodic"):
            h = torch.empty((m + 99) // 100, dtype=torch.float64, device=dev)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[::100] = h
        else:
            pct = float(label.split()[1].rstrip("%"))
            mask = torch.rand(m, device=dev) < pct / 100.0
            h = torch.empty(int(mask.sum().item()), dtype=torch.float64, device=dev)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[mask] = h
            del mask
        row = []
        for w, d in cols:
            fb.set_option("worklist", w)
            fb.set_option("direct_max", d)
            row.append(f"{rate(x, s[:m], c[:m]):.1f}")
        print(f"| {label} | " + " | ".join(row) + " |", flush=True)
    fb.set_option("direct_max", 4)
    print("\n# plain data (C3 inputs): the fixed cost of the launch sequence, Gelem/s")
    print("| n | one launch, pdl=1 | one launch, pdl=0 | hybrid, pdl=1 | hybrid, pdl=0 |")
    print("|---|---|---|---|---|")
    fb.fill_uniform_device(big, 0xFABE1303, -1e4, 1e4)
    for n in (125_000_000, 250_000_000, 1_000_000_000):
        row = []
        for w, p in ((2, 1), (2, 0), (3, 1), (3, 0)):
            fb.set_option("worklist", w)
            fb.set_option("pdl", p)
            row.append(f"{rate(big[:n], s[:n], c[:n], steps=20 if n < 1e9 else 6, reps=7):.1f}")
        print(f"| {n:.3g} | " + " | ".join(row) + " |", flush=True)
    fb.set_option("worklist", 0)
    fb.set_option("pdl", 1)


if __name__ == "__main__":
    main()

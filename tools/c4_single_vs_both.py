#!/usr/bin/env python
"""Is a single-output kernel slower in elements/s than the two-output one on the Payne-Hanek mix?  
The kernels alternate inside one process, round after round, so that they see the same clocks; per kernel: median and
best over the rounds, and the SM clock NVML reported during its steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def main():
    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    m = 100_000_000
    x = torch.empty(m, dtype=torch.float64, device=dev)
    a = torch.empty_like(x)
    b = torch.empty_like(x)
    fb.fill_loguniform_device(x, 0xFABE1304)
    kernels = {
        "sincos": lambda: fb.sincos(x, out_sin=a, out_cos=b),
        "sin": lambda: fb.sin_array(x, out=a),
        "cos": lambda: fb.cos_array(x, out=a),
        "tan": lambda: fb.tan_array(x, out=a),
    }
    rounds, steps = 25, 5
    times = {k: [] for k in kernels}
    clocks = {k: [] for k in kernels}
    for k in kernels.values():
        for _ in range(3):
            k()
    torch.cuda.synchronize()
    for r in range(rounds):
        order = list(kernels) if r % 2 == 0 else list(kernels)[::-1]
        for name in order:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                kernels[name]()
            e1.record()
            clocks[name].append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            torch.cuda.synchronize()
            times[name].append(e0.elapsed_time(e1) / steps)
    print(f"# C4 (N = 1e8, |x| log-uniform to 2^1024), kernels alternating, {rounds} rounds x {steps} steps")
    print("| kernel | Gelem/s median | best | worst | SM MHz median |")
    print("|---|---|---|---|---|")
    for name in kernels:
        t = sorted(times[name])
        c = sorted(clocks[name])
        print(f"| {name} | {m / t[len(t) // 2] / 1e6:.1f} | {m / t[0] / 1e6:.1f} | {m / t[-1] / 1e6:.1f} | {c[len(c) // 2]} |")


if __name__ == "__main__":
    main()

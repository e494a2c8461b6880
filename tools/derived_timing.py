#!/usr/bin/env python
"""Device-resident throughput of the single-output array functions (16 B/element: 8 read + 8 written):
sin, cos, tan, cot, sinc at N = 1e9 on C3 inputs, N = 1e8 on the Payne-Hanek mix."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402

PEAK = 6542.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, x, o, steps):
    for _ in range(3):
        fn(x, out=o)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn(x, out=o)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return ms[len(ms) // 2], ms[0]


def main():
    dev = torch.device("cuda:0")
    n = 1_000_000_000
    x = torch.empty(n, dtype=torch.float64, device=dev)
    o = torch.empty_like(x)
    print(f"# single-output array functions, 16 B/element, roofline denominator {PEAK:.1f} GB/s (measured copy peak)")
    print("| function | workload | N | Gelem/s (median) | GB/s | of measured | best GB/s |")
    print("|---|---|---|---|---|---|---|")
    fns = (("sin", fb.sin_array), ("cos", fb.cos_array), ("tan", fb.tan_array), ("cot", fb.cot_array), ("sinc", fb.sinc_array))
    fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
    for name, fn in fns:
        med, best = timed(fn, x, o, 20)
        print(f"| {name} | C3 uniform [-1e4, 1e4] | {n:.0e} | {n / med / 1e6:.1f} | {16 * n / med / 1e6:.0f} | "
              f"{16 * n / med / 1e6 / PEAK:.3f} | {16 * n / best / 1e6:.0f} |", flush=True)
    m = 100_000_000
    fb.fill_loguniform_device(x[:m], 0xFABE1304)
    # the two-output kernel on the same data, in the same process, before and after the single-output ones (the GPU
    # heats up over this script: a single-output kernel must not be slower in elements/s than the two-output one)
    o2 = torch.empty(m, dtype=torch.float64, device=dev)

    def both(t, out):
        fb.sincos(t, out_sin=out, out_cos=o2)

    med, best = timed(both, x[:m], o[:m], 10)
    print(f"| sincos (24 B/element) | C4 log-uniform to 2^1024 | {m:.0e} | {m / med / 1e6:.1f} | {24 * m / med / 1e6:.0f} | "
          f"{24 * m / med / 1e6 / PEAK:.3f} | {24 * m / best / 1e6:.0f} |", flush=True)
    for name, fn in fns + (("sincos (24 B/element), again", both),):
        if fn is both:
            med, best = timed(both, x[:m], o[:m], 10)
            print(f"| {name} | C4 log-uniform to 2^1024 | {m:.0e} | {m / med / 1e6:.1f} | {24 * m / med / 1e6:.0f} | "
                  f"{24 * m / med / 1e6 / PEAK:.3f} | {24 * m / best / 1e6:.0f} |", flush=True)
            continue
        med, best = timed(fn, x[:m], o[:m], 10)
        print(f"| {name} | C4 log-uniform to 2^1024 | {m:.0e} | {m / med / 1e6:.1f} | {16 * m / med / 1e6:.0f} | "
              f"{16 * m / med / 1e6 / PEAK:.3f} | {16 * m / best / 1e6:.0f} |", flush=True)


def main_f32():
    dev = torch.device("cuda:0")
    n = 2_000_000_000
    x = torch.empty(n, dtype=torch.float32, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    x.uniform_(-1e4, 1e4)
    print("\n# FP32 sincosf, 12 B/element")
    print("| function | workload | N | Gelem/s (median) | GB/s | of measured | best GB/s |")
    print("|---|---|---|---|---|---|---|")
    for _ in range(3):
        fb.sincosf(x, out_sin=s, out_cos=c)
    torch.cuda.synchronize()
    steps = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fb.sincosf(x, out_sin=s, out_cos=c)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    med, best = ms[len(ms) // 2], ms[0]
    print(f"| sincosf | uniform [-1e4, 1e4] | {n:.0e} | {n / med / 1e6:.1f} | {12 * n / med / 1e6:.0f} | "
          f"{12 * n / med / 1e6 / PEAK:.3f} | {12 * n / best / 1e6:.0f} |", flush=True)


if __name__ == "__main__":
    if "--f32-only" not in sys.argv:
        main()
    main_f32()

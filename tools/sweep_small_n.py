#!/usr/bin/env python
"""Device-resident small and medium arrays (BASELINE config 1 is N = 1e6): GPU time per call of the one-element-per-thread
kernel (n < small_n) against the fused streaming kernel (n >= small_n), to place option small_n.  Back-to-back calls on one
stream, CUDA events around groups of 50; microseconds per call and Gelem/s."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def per_call_us(x, s, c, calls=50, reps=7):
    fb.sincos(x, out_sin=s, out_cos=c)
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(calls):
            fb.sincos(x, out_sin=s, out_cos=c)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / calls)
    out.sort()
    return out[len(out) // 2]


def main():
    dev = torch.device("cuda:0")
    nmax = 1 << 23
    x = torch.empty(nmax, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    fb.fill_uniform_device(x, 0xFABE1301, -3.141592653589793, 3.141592653589793)
    saved = fb.get_option("small_n")
    print("| n | inline kernel us | fused kernel us | inline Gelem/s | fused Gelem/s |")
    print("|---|---|---|---|---|")
    for n in (1 << 12, 1 << 14, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1_000_000, 1 << 20, 1 << 21, 1 << 22, 1 << 23):
        fb.set_option("small_n", 1 << 30)
        t_inline = per_call_us(x[:n], s[:n], c[:n])
        fb.set_option("small_n", 0)
        t_fused = per_call_us(x[:n], s[:n], c[:n])
        print(f"| {n} | {t_inline:.2f} | {t_fused:.2f} | {n / t_inline / 1e3:.1f} | {n / t_fused / 1e3:.1f} |", flush=True)
    fb.set_option("small_n", saved)


if __name__ == "__main__":
    main()

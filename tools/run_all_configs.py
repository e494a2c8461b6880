#!/usr/bin/env python
"""All five BASELINE.json configs in one run on one GPU box: device-resident and host-resident throughput, fraction of
the HBM roofline at 24 B/element, the accuracy contract (value parity, k bit parity, specials) on a sample checked
against the oracle, and the reference's CPU library on the same inputs (1 thread and all threads).

    python tools/run_all_configs.py > profiles/r01_all_configs.md

Test-infrastructure use of oracle/ (as the checker and as the reported CPU baseline), like bench.py and tests/.
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fabe_b200 as fb  # noqa: E402
from oracle.loader import Oracle, Reference, cpu_model, ref_variant  # noqa: E402
from parity import check_against_oracle  # noqa: E402

PEAK = 6542.1
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
DEV = torch.device("cuda:0")
ORACLE = Oracle()


def device_time(x, steps):
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    for _ in range(3):
        fb.sincos(x, out_sin=s, out_cos=c)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fb.sincos(x, out_sin=s, out_cos=c)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
    return ms[len(ms) // 2] * 1e-3, ms[0] * 1e-3, s, c


def graph_time(x, calls=20, replays=20):
    """Small arrays are bound by the cost of issuing a launch from the host (Python + ctypes + one cudaLaunchKernelEx:
    ~10 us here, 4 us from C), not by the kernel: the same calls replayed from a CUDA graph show what the GPU side takes."""
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fb.sincos(x, out_sin=s, out_cos=c)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fb.sincos(x, out_sin=s, out_cos=c)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(replays + 1)]
    ev[0].record()
    for i in range(replays):
        g.replay()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(replays))
    return ms[len(ms) // 2] * 1e-3 / calls


def host_time(x_dev, n):
    px, ps, pc = fb.PinnedArray(n), fb.PinnedArray(n), fb.PinnedArray(n)
    px.array[:] = x_dev[:n].cpu().numpy()
    fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        fb.sincos(px.array, out_sin=ps.array, out_cos=pc.array)
        ts.append(time.perf_counter() - t0)
    for p in (px, ps, pc):
        p.free()
    return min(ts)


def cpu_reference(x_host, threads):
    path = ref_variant("buildsh")
    if path is None:
        return None, "oracle/_ref absent"
    ref = Reference(path=path)
    parts = np.array_split(x_host, threads)
    outs = [(np.zeros_like(p), np.zeros_like(p)) for p in parts]
    pool = ThreadPoolExecutor(max_workers=threads)
    run = lambda: list(pool.map(lambda i: ref.sincos_raw(parts[i], outs[i][0], outs[i][1]), range(threads)))
    run()
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        run()
        best = min(best, time.perf_counter() - t0)
    return x_host.size / best / 1e9, ref.name


def accuracy(x, s, c, with_k=True):
    """Contract check on a strided sample + both ends; returns (verdict, detail)."""
    n = x.numel()
    idx = torch.unique(torch.cat([torch.arange(0, min(n, 200_000), device=DEV), torch.arange(max(0, n - 200_000), n, device=DEV),
                                  torch.arange(0, n, max(1, n // 400_000), device=DEV)]))
    xs = x[idx].cpu().numpy()
    k = None
    if with_k:
        _, _, kk = fb.sincos_k(x[idx].contiguous())
        torch.cuda.synchronize()
        k = kk.cpu().numpy()
    try:
        check_against_oracle(xs, s[idx].cpu().numpy(), c[idx].cpu().numpy(), k, ORACLE, label="sample")
    except AssertionError as e:
        return "FAIL", str(e)
    ls, lc = ORACLE.libm(xs)
    fin = np.isfinite(xs)
    err = max(np.abs(s[idx].cpu().numpy() - ls)[fin].max(), np.abs(c[idx].cpu().numpy() - lc)[fin].max()) if fin.any() else 0.0
    return "pass", f"{idx.numel()} samples, max err vs libm {err:.2e}"


def edge_pool():
    z = np.load(os.path.join(ROOT, "tests", "golden", "fabe13_golden_v1.npz"))
    parts = [z[f"{name}/x"].view(np.float64) for name in ("specials", "pio2_multiples_small", "pio2_multiples_large", "pio4_odd_multiples")]
    parts.append(np.ldexp(1.0, -np.arange(1000, 1075)).astype(np.float64))
    parts.append(np.array([1e22, -1e300, 2.0 ** 45, 1.7976931348623157e308]))
    return np.concatenate(parts)


def main():
    threads = len(os.sched_getaffinity(0))
    print(f"# All BASELINE configs, one box: {fb.implementation_name()}; host {cpu_model()}, {threads} threads")
    print(f"\nRoofline denominator {PEAK:.1f} GB/s (measured copy peak), 24 B/element; median of the timed steps (best in parentheses).\n")
    print("| config | N | device-resident Gelem/s | GB/s | of measured | of 8 TB/s | host-resident (pinned) Gelem/s | reference CPU 1 thr / all thr Gelem/s | accuracy contract |")
    print("|---|---|---|---|---|---|---|---|---|")
    configs = [
        ("C1 uniform [-pi, pi]", 1_000_000, "uniform", (-np.pi, np.pi), 0xFABE1301, 50),
        ("C2 uniform [-1e4, 1e4]", 100_000_000, "uniform", (-1e4, 1e4), 0xFABE1302, 20),
        ("C3 uniform [-1e4, 1e4]", 1_000_000_000, "uniform", (-1e4, 1e4), 0xFABE1303, 20),
        ("C3b uniform [-1e5 pi, 1e5 pi] (reference benchmark range)", 1_000_000_000, "uniform", (-1e5 * np.pi, 1e5 * np.pi), 0xFABE1303, 10),
        ("C4 log-uniform to 2^1024 (95.6% Payne-Hanek)", 100_000_000, "loguniform", None, 0xFABE1304, 10),
        ("C4s C2 body, 1% lanes extreme", 100_000_000, "mix", (-1e4, 1e4), 0xFABE1302, 10),
        ("C5 edge/special cyclic pattern", 100_000_000, "edge", None, 0, 10),
    ]
    for name, n, gen, rng, seed, steps in configs:
        x = torch.empty(n, dtype=torch.float64, device=DEV)
        if gen == "uniform":
            fb.fill_uniform_device(x, seed, rng[0], rng[1])
        elif gen == "loguniform":
            fb.fill_loguniform_device(x, seed)
        elif gen == "mix":
            fb.fill_uniform_device(x, seed, rng[0], rng[1])
            h = torch.empty(n // 100, dtype=torch.float64, device=DEV)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[37::100] = h
        else:
            pool = torch.from_numpy(edge_pool()).to(DEV)
            x = pool.repeat(n // pool.numel() + 1)[:n].contiguous()
        med, best, s, c = device_time(x, steps)
        replay = graph_time(x) if n <= 4_000_000 else None
        verdict, detail = accuracy(x, s, c)
        n_host = min(n, 100_000_000)
        t_host = host_time(x, n_host)
        sample = x[: min(n, 20_000_000)].cpu().numpy()
        cpu1, isa = cpu_reference(sample, 1)
        cpun, _ = cpu_reference(x[: min(n, 20_000_000 * threads // 4)].cpu().numpy(), threads)
        gbs = 24 * n / med / 1e9
        dev_cell = f"{n / med / 1e9:.1f} ({n / best / 1e9:.1f})"
        if replay:
            dev_cell += f"; replayed from a CUDA graph {n / replay / 1e9:.1f} = {24 * n / replay / 1e9 / PEAK:.2f} of measured ({replay * 1e6:.1f} us per call)"
        print(f"| {name} | {n:.0e} | {dev_cell} | {gbs:.0f} | {gbs / PEAK:.3f} | {gbs / 8000:.3f} | "
              f"{n_host / t_host / 1e9:.2f} | {cpu1:.2f} / {cpun:.2f} ({isa}) | {verdict}: {detail} |", flush=True)
        del x, s, c
    print("\nSmall N (C5, launch-bound): see r01_abi_latency.txt.  Multi-GPU: r01_bench_c3_{2,4,8}gpu.json.")


if __name__ == "__main__":
    main()

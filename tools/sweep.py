#!/usr/bin/env python
"""Tuning sweep of the streaming kernel on one GPU (not part of the product): unroll x blocks_per_sm x core, and the
other BASELINE workloads.  Prints one line per configuration: Gelem/s, GB/s at 24 B/element, fraction of the measured
copy peak."""
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


def time_steps(x, s, c, steps=10, warm=3):
    for _ in range(warm):
        fb.sincos(x, out_sin=s, out_cos=c)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fb.sincos(x, out_sin=s, out_cos=c)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return sum(ms) / len(ms), min(ms)


def report(tag, n, avg, best):
    print(f"{tag:58s} avg {n / avg / 1e6:8.2f} Gelem/s {24 * n / avg / 1e6:8.1f} GB/s ({24 * n / avg / 1e6 / PEAK:5.3f})"
          f"  best {24 * n / best / 1e6:8.1f} GB/s ({24 * n / best / 1e6 / PEAK:5.3f})", flush=True)


def main():
    dev = torch.device("cuda:0")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    x = torch.empty(n, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    fb.fill_uniform_device(x, 0xFABE1303, -1e4, 1e4)
    for core in (0, 1):
        fb.set_option("core", core)
        for unroll in (1, 2):
            fb.set_option("unroll", unroll)
            for bps in (0, 4, 32):
                fb.set_option("blocks_per_sm", bps)
                for tpb in ((1, 2, 4, 8, 16) if bps == 0 else (1,)):
                    fb.set_option("tiles_per_block", tpb)
                    avg, best = time_steps(x, s, c)
                    report(f"c3 n={n} core={core} unroll={unroll} blocks_per_sm={bps} tiles_per_block={tpb}", n, avg, best)
    fb.set_option("tiles_per_block", 1)
    fb.set_option("core", 0)
    fb.set_option("unroll", 1)
    fb.set_option("blocks_per_sm", 0)
    # other sizes / workloads with the defaults
    for m in (10, 1000, 100_000, 1_000_000, 4_194_303, 4_194_304, 10_000_000, 100_000_000):
        avg, best = time_steps(x[:m], s[:m], c[:m], steps=20)
        print(f"uniform n={m:>11d}: avg {avg * 1e3:9.1f} us  best {best * 1e3:9.1f} us  {m / best / 1e6:9.3f} Gelem/s", flush=True)
    m = 100_000_000
    fb.fill_loguniform_device(x[:m], 0xFABE1304)
    for core in (0, 1):
        fb.set_option("core", core)
        for shift in (0, 3, 20):
            fb.set_option("worklist_shift", shift)
            for sp in (4, 8, 16):
                fb.set_option("second_pass_blocks_per_sm", sp)
                avg, best = time_steps(x[:m], s[:m], c[:m], steps=5)
                report(f"c4 dense loguniform n={m} core={core} worklist_shift={shift} second_pass_bps={sp}", m, avg, best)
    fb.set_option("second_pass_blocks_per_sm", 4)
    fb.set_option("core", 0)
    fb.set_option("worklist_shift", 3)
    # sparse mix: 1% extreme lanes
    fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
    h = torch.empty(m // 100, dtype=torch.float64, device=dev)
    fb.fill_loguniform_device(h, 0xFABE1304)
    x[:m][37::100] = h
    avg, best = time_steps(x[:m], s[:m], c[:m], steps=5)
    report(f"c4 sparse 1% extreme n={m}", m, avg, best)
    fb.set_option("worklist_shift", 20)
    avg, best = time_steps(x[:m], s[:m], c[:m], steps=5)
    report(f"c4 sparse 1% extreme n={m} (no worklist: rescan only)", m, avg, best)


if __name__ == "__main__":
    main()

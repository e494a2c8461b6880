#!/usr/bin/env python
"""Small-N launch overhead (BASELINE config 5): 200 calls of fabe13_sincos_device back to back on one stream, issued
directly and replayed from a CUDA graph that captured them (the call is one kernel launch with no allocation below
2^26 elements, so it captures as it is)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    calls = 200
    # which kernel serves small arrays: the one-element-per-thread kernel below option small_n (default 2^20), the
    # streaming kernel (256-bit accesses, programmatic dependent launch) from there on -- python tools/graph_latency.py 0
    # measures the streaming kernel at every size
    if len(sys.argv) > 1:
        fb.set_option("small_n", int(sys.argv[1]))
    print(f"# small_n = {fb.get_option('small_n')}")
    print("| n per call | direct, us per call | graph replay, us per call |")
    print("|---|---|---|")
    for n in (10, 1000, 10_000, 32_768, 100_000, 300_000, 1_000_000, 2_000_000, 4_000_000):
        x = torch.empty(n, dtype=torch.float64, device=dev).uniform_(-1e4, 1e4)
        s = torch.empty_like(x)
        c = torch.empty_like(x)
        for _ in range(5):
            fb.sincos(x, out_sin=s, out_cos=c)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best_direct = 1e9
        for _ in range(5):
            a.record()
            for _ in range(calls):
                fb.sincos(x, out_sin=s, out_cos=c)
            b.record()
            torch.cuda.synchronize()
            best_direct = min(best_direct, a.elapsed_time(b) * 1e3 / calls)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fb.sincos(x, out_sin=s, out_cos=c)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(calls):
                fb.sincos(x, out_sin=s, out_cos=c)
        g.replay()
        torch.cuda.synchronize()
        best_graph = 1e9
        for _ in range(5):
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            best_graph = min(best_graph, a.elapsed_time(b) * 1e3 / calls)
        print(f"| {n} | {best_direct:.2f} | {best_graph:.2f} |", flush=True)


if __name__ == "__main__":
    main()

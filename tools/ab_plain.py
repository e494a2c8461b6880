#!/usr/bin/env python
"""Plain-data A/B of kernel-variant builds in ONE process (alternating, so clocks and temperature are shared):
    python tools/ab_plain.py name=path.so [name=path.so ...]
Each library is loaded separately through ctypes; C3 (1e9) burst + sustained and C2 (1e8), device-resident."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fabe_b200.build import LIBPATH  # noqa: E402

vp = ctypes.c_void_p


def load(path):
    lib = ctypes.CDLL(os.path.abspath(path))
    lib.fabe13_sincos_device.argtypes = [vp, vp, vp, ctypes.c_size_t, vp]
    lib.fabe13_sincos_device.restype = ctypes.c_int
    lib.fabe13_debug_fill_uniform_device.argtypes = [vp, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_double, ctypes.c_double, vp]
    lib.fabe13_cuda_set_option.argtypes = [ctypes.c_char_p, ctypes.c_longlong]
    return lib


def run(lib, x, s, c, n, steps):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        lib.fabe13_sincos_device(x.data_ptr(), s.data_ptr(), c.data_ptr(), n, st)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        lib.fabe13_sincos_device(x.data_ptr(), s.data_ptr(), c.data_ptr(), n, st)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return n / (sum(ms) / len(ms)) / 1e6, n / min(ms) / 1e6


def main():
    c4 = "--c4" in sys.argv
    libs = [("default", LIBPATH)] + [tuple(a.split("=", 1)) for a in sys.argv[1:] if not a.startswith("--")]
    loaded = []
    for name, spec in libs:  # spec = path[:option=value[:option=value ...]]
        path, *opts = spec.split(":")
        lib = load(path)
        for o in opts:
            k, v = o.split("=")
            assert lib.fabe13_cuda_set_option(k.encode(), int(v)) == 0
        loaded.append((name, lib))
    libs = loaded
    dev = torch.device("cuda:0")
    n = 1_000_000_000
    x = torch.empty(n, dtype=torch.float64, device=dev)
    s = torch.empty_like(x)
    c = torch.empty_like(x)
    libs[0][1].fabe13_debug_fill_uniform_device(x.data_ptr(), n, 0xFABE1303, 0, -1e4, 1e4, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    if c4:  # BASELINE config 4: N = 1e8, |x| log-uniform to 2^1024 (dense Payne-Hanek), plus two mixes
        import fabe_b200 as fb

        m = 100_000_000
        for label, step in (("c4 dense", 1), ("every 2nd huge", 2), ("5% huge", 20)):
            fb.fill_uniform_device(x[:m], 0xFABE1302, -1e4, 1e4)
            h = torch.empty((m + step - 1) // step, dtype=torch.float64, device=dev)
            fb.fill_loguniform_device(h, 0xFABE1304)
            x[:m][::step] = h
            torch.cuda.synchronize()
            for rep in range(3):
                for name, lib in libs:
                    avg, best = run(lib, x, s, c, m, 20)
                    print(f"rep {rep} {name:24s} {label:16s} 1e8 (20)  avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)
        return
    for rep in range(3):
        for name, lib in libs:
            avg, best = run(lib, x, s, c, n, 20)
            print(f"rep {rep} {name:24s} c3 1e9 burst(20)      avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)
    for rep in range(2):
        for name, lib in libs:
            avg, best = run(lib, x, s, c, n, 200)
            print(f"rep {rep} {name:24s} c3 1e9 sustained(200) avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)
    for rep in range(3):
        for name, lib in libs:
            avg, best = run(lib, x, s, c, 100_000_000, 30)
            print(f"rep {rep} {name:24s} c2 1e8 (30)           avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)
    for rep in range(3):  # the 8-GPU slice of C3: the launch sequence's fixed cost and the grid's tail show here
        for name, lib in libs:
            avg, best = run(lib, x, s, c, 125_000_000, 40)
            print(f"rep {rep} {name:24s} c3 slice 1.25e8 (40)  avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)
    for n_small in (1_000_000, 10_000_000):
        for rep in range(2):
            for name, lib in libs:
                avg, best = run(lib, x, s, c, n_small, 50)
                print(f"rep {rep} {name:24s} n = {n_small:<9d} (50)     avg {avg:7.2f} best {best:7.2f} Gelem/s", flush=True)


if __name__ == "__main__":
    main()

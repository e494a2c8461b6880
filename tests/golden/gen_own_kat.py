#!/usr/bin/env python
"""Known-answer vectors of THIS implementation (not of the reference: those are in fabe13_golden_v1.npz).

    python tests/golden/gen_own_kat.py      # rewrites tests/golden/fabe13_own_kat_v1.npz

Inputs: a fixed mix of plain, tiny, special and Payne-Hanek arguments; outputs: the bit patterns the host instantiation
of the shared core produces for sincos (both cores), tan / cot / sinc and the FP32 entry.  The device agrees with the
host instantiation bit for bit (tests/test_gpu_parity.py), so these vectors pin the numerics of the whole library: a
change of a single result bit anywhere -- a refitted coefficient, a reordered FMA, a different reduction -- shows up
here, on the CPU, and has to be made on purpose (regenerate and say why in the commit)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import fabe_b200 as fb  # noqa: E402


def inputs():
    rng = np.random.default_rng(0xFABE13)
    parts = [
        rng.uniform(-np.pi, np.pi, 400), rng.uniform(-1e4, 1e4, 400), rng.uniform(-1e6, 1e6, 200), rng.uniform(-3e13, 3e13, 200),
        np.ldexp(rng.uniform(1, 2, 400), rng.integers(45, 1024, 400)) * rng.choice([-1.0, 1.0], 400),       # Payne-Hanek
        np.ldexp(rng.uniform(1, 2, 100), rng.integers(-1074, -27, 100)) * rng.choice([-1.0, 1.0], 100),     # tiny
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 2.0 ** -27, 2.0 ** 45, -(2.0 ** 45), 1e22, 1.7976931348623157e308,
                  6381956970095103.0 * 2.0 ** 797, 5e-324, 1e-9, 9.99e-10, np.pi / 2, np.pi, 3 * np.pi / 4]),
    ]
    return np.concatenate(parts)


def main():
    x = inputs()
    out = {"x": x.view(np.uint64)}
    for core, tag in ((0, "poly"), (1, "psi")):
        s, c, k = fb.host_core(x, core)
        out[f"{tag}/sin"], out[f"{tag}/cos"], out[f"{tag}/k"] = s.view(np.uint64), c.view(np.uint64), k
        for op, name in ((fb.OP_TAN, "tan"), (fb.OP_COT, "cot"), (fb.OP_SINC, "sinc")):
            out[f"{tag}/{name}"] = fb.host_derived(x, op, core).view(np.uint64)
        with np.errstate(all="ignore"):
            xf = x.astype(np.float32)
        sf, cf = fb.host_sincosf(xf, core)
        out[f"{tag}/sinf"], out[f"{tag}/cosf"] = sf.view(np.uint32), cf.view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "fabe13_own_kat_v1.npz"), **out)
    print("wrote", len(x), "inputs,", len(out) - 1, "output arrays")


if __name__ == "__main__":
    main()

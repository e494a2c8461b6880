#!/usr/bin/env python
"""Accuracy table: max |error| against glibc for this implementation (both cores; host instantiation, which the GPU
tests show to be bit-identical to the device) and for the reference's strict and build.sh builds, per input range.

    python tools/accuracy_report.py > profiles/r01_accuracy.txt        (needs oracle/_ref, i.e. the build container)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fabe_b200 as fb  # noqa: E402
from oracle.loader import Oracle, Reference, ref_variant  # noqa: E402


def err(s, c, ls, lc):
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.maximum(np.abs(s - ls), np.abs(c - lc))
    d = np.where(np.isfinite(d), d, np.inf)
    return d.max()


def main():
    o = Oracle()
    refs = {}
    for kind in ("strict", "buildsh"):
        p = ref_variant(kind)
        if p:
            refs[kind] = Reference(path=p)
    n = 1 << 20
    ranges = [("[-pi, pi]", -np.pi, np.pi), ("[-1e4, 1e4]", -1e4, 1e4), ("[-1e5*pi, 1e5*pi]", -1e5 * np.pi, 1e5 * np.pi),
              ("[-1e6, 1e6]", -1e6, 1e6), ("[-1e9, 1e9]", -1e9, 1e9), ("[-3e13, 3e13]", -3e13, 3e13),
              ("[1e15, 1e16]", 1e15, 1e16)]
    print(f"max(|sin err|, |cos err|) against glibc, {n} uniform samples per range")
    print(f"{'range':24s} {'this (poly)':>14s} {'this (psi)':>14s} {'reference strict':>18s} {'reference build.sh':>20s}")
    rows = []
    for name, lo, hi in ranges:
        x = o.fill_uniform(n, 0xACC0 + len(rows), lo, hi)
        rows.append((name, x))
    rows.append(("log-uniform to 2^1024", o.fill_loguniform(n, 0xACCF)))
    for name, x in rows:
        ls, lc = o.libm(x)
        cols = []
        for core in (0, 1):
            s, c, _ = fb.host_core(x, core)
            cols.append(err(s, c, ls, lc))
        for kind in ("strict", "buildsh"):
            if kind in refs:
                s, c = refs[kind].sincos_bulk(x)
                cols.append(err(s, c, ls, lc))
            else:
                cols.append(float("nan"))
        print(f"{name:24s} {cols[0]:14.3e} {cols[1]:14.3e} {cols[2]:18.3e} {cols[3]:20.3e}")
    # the reference's tails / n < 32 / scalar API (its Psi core)
    x = o.fill_uniform(4096, 0xACD0, -10.0, 10.0)
    ls, lc = o.libm(x)
    rs = np.array([o.scalar_core(v) for v in x])
    s, c, _ = fb.host_core(x, 0)
    print(f"\nscalar API / tails on [-10, 10]: this {err(s, c, ls, lc):.3e}   reference scalar Psi core {err(rs[:, 0], rs[:, 1], ls, lc):.3e}")
    print("quadrant index k: identical to the reference formula on every sample above with |x| < 2^45:",
          all(np.array_equal(fb.host_core(x, 0)[2][np.abs(x) < 2.0 ** 45], o.k(x)[np.abs(x) < 2.0 ** 45]) for _, x in rows))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generate tests/golden/fabe13_golden_v1.npz from the UNMODIFIED reference (oracle/_ref strict build).

The reference has no golden vectors of its own, so the pins are outputs of the reference
itself, produced in the build container where /root/reference is mounted, and committed so the GPU box (which has
no reference source) can check against them.  Run:  make -C oracle all && python oracle/gen_golden.py

Stored per case (all as uint64 bit patterns, so nothing is lost in text round trips):
  x        inputs
  ref_sin, ref_cos      reference strict build, bulk SIMD path (calls padded to n%8==0, n>=32)
  libm_sin, libm_cos    glibc sin/cos of the build container (truth for |x| > 1e6, where the reference is off)
  q_exact               rint(x*2/pi) mod 4 from exact big-integer arithmetic (1400-bit 2/pi)
  k_ref                 the reference's k formula (fabe13/fabe13.c:568-571), via the restated oracle
"""
import os
import sys
from fractions import Fraction

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.loader import Oracle, Reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fabe13_golden_v1.npz")

mp.mp.prec = 1500
TWO_OVER_PI_BITS = 1400
TWO_OVER_PI_INT = int(mp.floor((2 / mp.pi) * mp.mpf(2) ** TWO_OVER_PI_BITS))


def q_exact(x):
    """rint(x * 2/pi) mod 4 exactly (x finite).  Ties cannot occur (2/pi is irrational)."""
    if x == 0.0:
        return 0
    fr = Fraction(x)  # exact
    num = fr.numerator * TWO_OVER_PI_INT
    den = fr.denominator << TWO_OVER_PI_BITS
    k = (2 * num + den) // (2 * den)  # floor(v + 1/2)
    return int(k % 4)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def from_bits(values):
    return np.array(values, dtype=np.uint64).view(np.float64)


def case_specials():
    # SURVEY.md section 8c "golden facts" list
    return from_bits([
        0x0000000000000000, 0x8000000000000000,            # +-0
        0x0000000000000001, 0x8000000000000001,            # smallest subnormals
        0x000FFFFFFFFFFFFF, 0x0010000000000000,            # largest subnormal, DBL_MIN
        bits(np.array([1e-300]))[0], bits(np.array([-1e-300]))[0],
        0x3E40000000000000, 0xBE40000000000000,            # +-2^-27 (inclusive tiny threshold)
        0x3E40000000000001, 0xBE40000000000001,            # one ulp above: first value on the polynomial path
        bits(np.array([1e-8]))[0],
        0x7FF0000000000000, 0xFFF0000000000000,            # +-Inf
        0x7FF8000000000000, 0xFFF8000000000000,            # +-qNaN
        0x7FF0000000000001, 0xFFF8000000001234,            # sNaN, payload NaN
        0x3FF921FB54442D18, 0x400921FB54442D18, 0x3FE921FB54442D18,  # pi/2, pi, pi/4
    ])


def case_reference_test_inputs():
    # the input lists of the reference's example test: tests/test_fabe13.c:107-112 (scalars) and :200-202 (ramp)
    pi = np.pi
    scalars = [0.0, 1.0, -1.0, pi / 6.0, pi / 4.0, pi / 2.0, pi, 3.0 * pi / 2.0, 2.0 * pi, 1000.0, -1000.0, 1e-10,
               -1e-10, 0.5, -0.5, 2.0, -2.0, np.inf, -np.inf, np.nan]
    ramp = [float(i) * 0.1 - 5.0 for i in range(100)]
    return np.array(scalars + ramp, dtype=np.float64)


def case_pio2_multiples(m_values):
    # RN(m*pi/2) and its +-1, +-2 ulp neighbours, both signs: the hardest inputs for k and for cancellation in r
    out = []
    for m in m_values:
        c = float(mp.nstr(mp.mpf(m) * mp.pi / 2, 40))
        for _ in range(1):
            lo2 = np.nextafter(np.nextafter(c, -np.inf), -np.inf)
            lo1 = np.nextafter(c, -np.inf)
            hi1 = np.nextafter(c, np.inf)
            hi2 = np.nextafter(hi1, np.inf)
            for v in (lo2, lo1, c, hi1, hi2):
                out += [v, -v]
    return np.array(out, dtype=np.float64)


def main():
    oracle = Oracle()
    ref = Reference(kind="strict")
    print("reference:", ref.path, ref.name)
    cases = {}
    cases["specials"] = case_specials()
    cases["reference_test_inputs"] = case_reference_test_inputs()
    cases["pio2_multiples_small"] = case_pio2_multiples(list(range(0, 400)))
    cases["pio2_multiples_large"] = case_pio2_multiples(
        [10**3 + 1, 10**4 + 7, 10**5 + 3, 636619, 636620, 10**6 + 1, 10**7 + 9, 10**9 + 7, 10**12 + 39, 2**40 + 1, 2**44 + 5])
    # odd multiples of pi/4 (where k flips) +-2 ulp
    quarter = []
    for m in list(range(1, 200, 2)) + [2 * 10**5 + 1, 2 * 10**6 + 1, 1273239, 1273241]:
        c = float(mp.nstr(mp.mpf(m) * mp.pi / 4, 40))
        v = c
        for _ in range(2):
            v = np.nextafter(v, -np.inf)
        for _ in range(5):
            quarter += [v, -v]
            v = np.nextafter(v, np.inf)
    cases["pio4_odd_multiples"] = np.array(quarter, dtype=np.float64)
    seed = 0xFABE1300
    n = 1024
    cases["uniform_pi"] = oracle.fill_uniform(n, seed + 1, -np.pi, np.pi)
    cases["uniform_1e4"] = oracle.fill_uniform(n, seed + 2, -1e4, 1e4)
    cases["uniform_1e5pi"] = oracle.fill_uniform(n, seed + 3, -1e5 * np.pi, 1e5 * np.pi)  # benchmark_fabe13.c:192 range
    cases["uniform_1e6"] = oracle.fill_uniform(n, seed + 4, -1e6, 1e6)
    cases["uniform_1e9"] = oracle.fill_uniform(n, seed + 5, -1e9, 1e9)
    cases["uniform_3e13"] = oracle.fill_uniform(n, seed + 6, -3e13, 3e13)
    cases["edge_2p45"] = np.concatenate([
        oracle.fill_uniform(n // 2, seed + 7, 2.0**44, 2.0**46),
        -oracle.fill_uniform(n // 4, seed + 8, 2.0**44, 2.0**46),
        from_bits([0x42BFFFFFFFFFFFFF, 0x42C0000000000000, 0x42C0000000000001, 0xC2BFFFFFFFFFFFFF, 0xC2C0000000000000]),
    ])
    cases["loguniform_huge"] = np.concatenate([
        oracle.fill_loguniform(n, seed + 9),
        np.array([1e22, -1e22, 1.7976931348623157e308, -1.7976931348623157e308, 2.0**1023, 6381956970095103.0 * 2.0**797]),
    ])
    names = sorted(cases)
    out = {"case_names": np.array(names)}
    for name in names:
        x = np.ascontiguousarray(cases[name], dtype=np.float64)
        rs, rc = ref.sincos_bulk(x)
        os_, oc_, ok_ = oracle.sincos(x)
        # the restatement must be bit-identical to the reference on every pinned input
        assert np.array_equal(bits(rs), bits(os_)) and np.array_equal(bits(rc), bits(oc_)), name
        ls, lc = oracle.libm(x)
        finite = np.isfinite(x)
        q = np.array([q_exact(float(v)) if f else -1 for v, f in zip(x, finite)], dtype=np.int64)
        out[f"{name}/x"] = bits(x)
        out[f"{name}/ref_sin"] = bits(rs)
        out[f"{name}/ref_cos"] = bits(rc)
        out[f"{name}/libm_sin"] = bits(ls)
        out[f"{name}/libm_cos"] = bits(lc)
        out[f"{name}/q_exact"] = q
        out[f"{name}/k_ref"] = ok_
        err = np.nanmax(np.abs(np.where(finite, rs - ls, 0.0))) if finite.any() else 0.0
        print(f"{name:24s} n={x.size:5d}  max|ref-libm| sin {err:.3e}")
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()

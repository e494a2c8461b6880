"""The per-element algorithm (fabe_b200/csrc/fabe13_core.h) is one piece of code compiled for the device and for
the host.  These tests run its HOST instantiation (the implementation of the scalar API, reached in bulk through the
fabe13_debug_host_core test hook) against the oracle, the golden vectors and exact quadrants -- so the algorithm is
checked on every commit even where no GPU exists.  The `-m gpu` tests then check the device instantiation."""
import numpy as np
import pytest

from parity import HUGE, QUALITY_TOL, bits, check_against_oracle, check_k, check_values

CORES = [0, 1]


@pytest.mark.parametrize("core", CORES)
def test_golden_cases(fabe13, golden, core):
    for name, g in golden.items():
        s, c, k = fabe13.host_core(g["x"], core)
        check_values(g["x"], s, c, g["ref_sin"], g["ref_cos"], g["libm_sin"], g["libm_cos"], label=f"{name}/core{core}")
        check_k(g["x"], k, g["k_ref"], g["q_exact"], label=name)


@pytest.mark.parametrize("core", CORES)
def test_uniform_ranges_against_oracle(fabe13, oracle, core):
    for seed, lo, hi in [(0xFABE1301, -np.pi, np.pi), (0xFABE1302, -1e4, 1e4), (0xFABE1303, -1e5 * np.pi, 1e5 * np.pi),
                         (21, -1e6, 1e6), (22, -1e9, 1e9), (23, -3e13, 3e13)]:
        x = oracle.fill_uniform(200_000, seed, lo, hi)
        s, c, k = fabe13.host_core(x, core)
        check_against_oracle(x, s, c, k, oracle, label=f"[{lo},{hi}] core{core}")


@pytest.mark.parametrize("core", CORES)
def test_extreme_range_against_libm(fabe13, oracle, core):
    x = oracle.fill_loguniform(200_000, 0xFABE1304)
    s, c, k = fabe13.host_core(x, core)
    ls, lc = oracle.libm(x)
    assert max(np.abs(s - ls).max(), np.abs(c - lc).max()) <= QUALITY_TOL
    assert set(np.unique(k[np.abs(x) >= HUGE])) <= {0, 1, 2, 3}


def test_payne_hanek_quadrants_are_exact(fabe13, oracle):
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    from gen_golden import q_exact  # exact big-integer rint(x*2/pi) mod 4

    x = oracle.fill_loguniform(4000, 77)
    x = x[np.abs(x) >= HUGE]
    _, _, k = fabe13.host_core(x, 0)
    want = np.array([q_exact(float(v)) for v in x])
    assert np.array_equal(k, want)


def _reduce_huge(fabe13, x):
    lib = fabe13._lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = x.size
    r, re = np.empty(n), np.empty(n)
    q, qe, ok = (np.empty(n, dtype=np.int32) for _ in range(3))
    lib.fabe13_debug_reduce_huge(x.ctypes.data, r.ctypes.data, q.ctypes.data, re.ctypes.data, qe.ctypes.data, ok.ctypes.data, n)
    return r, q, re, qe, ok


def _huge_from(rng, n, mantissa=None):
    e = rng.integers(1068, 2047, n).astype(np.uint64)
    m = rng.integers(0, 1 << 52, n).astype(np.uint64) if mantissa is None else mantissa(rng.integers(0, 1 << 52, n).astype(np.uint64))
    sgn = rng.integers(0, 2, n).astype(np.uint64)
    return ((sgn << np.uint64(63)) | (e << np.uint64(52)) | m).view(np.float64)


def test_fast_large_argument_reduction_agrees_with_the_exact_one(fabe13):
    """Arguments >= 2^45 are reduced by the fast form (tabulated remainders of the powers of two modulo 2 pi, TwoSum,
    Cody-Waite: csrc/fabe13_core.h 2c), which hands the rare result it cannot vouch for to the exact 192-bit form.
    Over every exponent and over mantissas chosen to hurt the split M = Mh + Ml: the quadrant is ALWAYS the exact one,
    r is within one ulp of the exact form's r (both are within ~half an ulp of the truth), the fast form vouches for all
    but ~1e-5 of its results, and where it does not the product returns the exact form's bits."""
    rng = np.random.default_rng(20261020)
    low27 = np.uint64((1 << 27) - 1)
    cases = [
        _huge_from(rng, 600_000),
        _huge_from(rng, 100_000, lambda m: m & ~low27),                                          # Ml = 0
        _huge_from(rng, 100_000, lambda m: (m & ~low27) | (m & np.uint64(3))),                   # Ml = 0 .. 3
        _huge_from(rng, 100_000, lambda m: m & low27),                                            # Mh minimal
        _huge_from(rng, 100_000, lambda m: m | low27),                                            # Ml maximal
        np.ldexp(1.0, np.arange(45, 1024)), -np.ldexp(1.0 + 2.0 ** -52, np.arange(45, 1024)),     # every exponent, extreme mantissas
        np.ldexp(2.0 - 2.0 ** -52, np.arange(45, 1024)),
    ]
    x = np.concatenate(cases)
    r, q, re, qe, ok = _reduce_huge(fabe13, x)
    assert np.array_equal(q, qe)
    assert np.abs(r.view(np.int64) - re.view(np.int64)).max() <= 1
    assert np.array_equal(bits(r[ok == 0]), bits(re[ok == 0]))
    assert ok.mean() > 0.9999
    assert np.all(np.abs(r[ok == 1]) >= 2.0 ** -20) and np.all(np.abs(r[ok == 1]) < 0.7853943487001827)


def test_fast_large_argument_reduction_hands_over_where_it_must(fabe13):
    """Inputs found by search (30e6 random huge doubles) whose remainder lies in one of the two zones the fast form does
    not vouch for -- |r| < 2^-20 (relative accuracy) and |r| > pi/4 - 2^-18 (the rounding of k) -- plus the double
    closest to a multiple of pi/2: the fast form reports them, and the product's result is the exact form's."""
    near_zero = ["0x1.04583f610414bp+164", "0x1.ca3b5a182946dp+798", "0x1.55148d5cb8e00p+747", "0x1.f7f4f6fc7b558p+594",
                 "0x1.b8dce83231005p+70", "0x1.d395978f3c6ccp+329"]
    near_tie = ["0x1.d7f618bf16351p+345", "0x1.82ebc2139778ap+741", "0x1.2c902b36da02ep+610", "0x1.3ba460994ce1fp+532",
                "0x1.c9b83f1554344p+415", "0x1.540e9c14dcf46p+511"]
    x = np.array([float.fromhex(h) for h in near_zero + near_tie] + [6381956970095103.0 * 2.0 ** 797])
    x = np.concatenate([x, -x])
    r, q, re, qe, ok = _reduce_huge(fabe13, x)
    assert not ok.any()
    assert np.array_equal(bits(r), bits(re)) and np.array_equal(q, qe)
    assert np.all(np.abs(r[:6]) < 2.0 ** -20) and np.all(np.abs(r[6:12]) > 0.78539434) and abs(r[12]) < 2.0 ** -60
    # arguments outside the reduction's range are refused, not looked up
    r, q, re, qe, ok = _reduce_huge(fabe13, np.array([1.0, -3.0e13, 0.0, np.inf, np.nan, 5e-324]))
    assert np.all(np.isnan(r)) and np.all(np.isnan(re)) and np.all(q == -1) and np.all(qe == -1) and not ok.any()


def test_fast_large_argument_reduction_against_mpmath(fabe13):
    """The remainder itself against 1400-bit arithmetic: r = x - rint(x 2/pi) pi/2 to within 0.501 ulp, q exact."""
    import mpmath as mp

    mp.mp.prec = 1400
    rng = np.random.default_rng(7)
    x = np.abs(_huge_from(rng, 1500))
    r, q, _, _, ok = _reduce_huge(fabe13, x)
    pio2 = mp.pi / 2
    worst = 0.0
    for xi, ri, qi in zip(x, r, q):
        xm = mp.mpf(float(xi))
        k = mp.nint(xm / pio2)
        true = xm - k * pio2
        assert int(k) % 4 == qi, xi
        ulp = 2.0 ** (np.frexp(ri)[1] - 53)
        worst = max(worst, float(abs(mp.mpf(float(ri)) - true) / ulp))
    assert worst <= 0.501, worst
    assert ok.sum() >= 1490
    # ... and where the fast form's absolute accuracy (~2^-74) weighs most: remainders just above its hand-over
    # threshold, 2^-20 <= |r| < 2^-14 (a few hundred of 3e6 random arguments)
    x = np.abs(_huge_from(rng, 3_000_000))
    r, q, _, _, ok = _reduce_huge(fabe13, x)
    near = (ok == 1) & (np.abs(r) < 2.0 ** -14)
    assert 100 <= near.sum() <= 2000, near.sum()
    worst = 0.0
    for xi, ri, qi in zip(x[near], r[near], q[near]):
        xm = mp.mpf(float(xi))
        k = mp.nint(xm / pio2)
        assert int(k) % 4 == qi, xi
        ulp = 2.0 ** (np.frexp(ri)[1] - 53)
        worst = max(worst, float(abs(mp.mpf(float(ri)) - (xm - k * pio2)) / ulp))
    assert worst <= 0.75, worst


def test_ph_remainder_table_is_what_the_generator_produces(tmp_path):
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "rem.h"
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_ph_remainders.py"), str(out)], check=True, capture_output=True)
    assert out.read_text() == open(os.path.join(root, "fabe_b200", "csrc", "fabe13_ph_remainders.h")).read()


def test_scalar_api_is_the_array_core(fabe13, oracle):
    x = np.concatenate([oracle.fill_uniform(300, 5, -50.0, 50.0), oracle.fill_loguniform(100, 6)])
    s, c, _ = fabe13.host_core(x, 0)
    for xi, si, ci in zip(x, s, c):
        assert bits(np.array([fabe13.sin(xi)]))[0] == bits(np.array([si]))[0]
        assert bits(np.array([fabe13.cos(xi)]))[0] == bits(np.array([ci]))[0]


def test_scalar_api_deviates_from_reference_scalar_core_on_purpose(fabe13, oracle):
    # reference fabe13_sin(1.0) = 0.8734 (SURVEY finding 2); ours is the correctly rounded neighbourhood of sin(1)
    ref_s, _ = oracle.scalar_core(1.0)
    assert abs(ref_s - np.sin(1.0)) > 3e-2
    assert abs(fabe13.sin(1.0) - np.sin(1.0)) <= 1.2e-16


def test_symmetry_bit_exact(fabe13, oracle):
    x = np.concatenate([oracle.fill_uniform(50_000, 31, 0.0, 1e6), np.abs(oracle.fill_loguniform(50_000, 32))])
    s, c, _ = fabe13.host_core(x, 0)
    sn, cn, _ = fabe13.host_core(-x, 0)
    assert np.array_equal(bits(sn), bits(-s)) and np.array_equal(bits(cn), bits(c))


def test_pythagorean_identity(fabe13, oracle):
    x = np.concatenate([oracle.fill_uniform(100_000, 41, -1e6, 1e6), oracle.fill_loguniform(100_000, 42)])
    for core in CORES:
        s, c, _ = fabe13.host_core(x, core)
        assert np.abs(s * s + c * c - 1.0).max() <= 8e-16


def test_tiny_threshold_high_word_test_is_equivalent(fabe13, oracle):
    """The device classifies |x| <= 2^-27 by the high word only, which also sweeps in 2^-27 < |x| < 2^-27(1+2^-20).
    There the polynomial path returns exactly (x, 1.0) too, so outputs must still equal the reference's bit for bit
    on both sides of the high-word boundary."""
    base = np.uint64(0x3E40000000000000)
    xs = np.concatenate([
        (base + np.arange(0, 1 << 20, 7, dtype=np.uint64)).view(np.float64),                           # in the sliver
        (base + np.uint64((1 << 32) - 1) - np.arange(0, 50_000, dtype=np.uint64)).view(np.float64),    # its top end
        (base + np.uint64(1 << 32) + np.arange(0, 50_000, dtype=np.uint64)).view(np.float64),          # just above it
    ])
    xs = np.concatenate([xs, -xs])
    s, c, _ = fabe13.host_core(xs, 0)
    rs, rc, _ = oracle.sincos(xs)
    assert np.array_equal(bits(s), bits(rs)) and np.array_equal(bits(c), bits(rc))


def test_against_mpmath_truth(fabe13, oracle):
    """Independent truth (mpmath, 1200-bit pi) rather than glibc, on inputs chosen to hurt: huge magnitudes, the
    double closest to a multiple of pi/2 (6381956970095103 * 2^797), both ends of the Payne-Hanek threshold."""
    import mpmath as mp

    mp.mp.prec = 1400
    xs = np.concatenate([
        np.array([1e22, -1e22, 1.7976931348623157e308, 2.0 ** 1023, 6381956970095103.0 * 2.0 ** 797, 2.0 ** 45,
                  np.nextafter(2.0 ** 45, 0), 3.0e13, 1e15, 1e16, 1e100, 5e-324, 1e-300, 0.7853981633974483]),
        oracle.fill_loguniform(150, 1234),
        oracle.fill_uniform(150, 1235, -1e6, 1e6),
    ])
    for core, tol in ((0, 1.2e-16), (1, 4.5e-16)):
        s, c, _ = fabe13.host_core(xs, core)
        for x, si, ci in zip(xs, s, c):
            xm = mp.mpf(float(x))
            assert abs(mp.mpf(float(si)) - mp.sin(xm)) <= tol, (x, si)
            assert abs(mp.mpf(float(ci)) - mp.cos(xm)) <= tol, (x, ci)


def test_random_bit_patterns_host(fabe13, oracle):
    """Same fuzz as the GPU test, on the host instantiation (runs in CPU-only CI)."""
    rng = np.random.default_rng(20261019)
    x = rng.integers(0, 1 << 64, size=500_000, dtype=np.uint64).view(np.float64)
    for core in (0, 1):
        s, c, k = fabe13.host_core(x, core)
        fin = np.isfinite(x)
        ls, lc = oracle.libm(x[fin])
        assert max(np.abs(s[fin] - ls).max(), np.abs(c[fin] - lc).max()) <= QUALITY_TOL
        assert np.all(bits(s)[~fin] == np.uint64(0x7FF8000000000000)) and np.all(bits(c)[~fin] == np.uint64(0x7FF8000000000000))
        tiny = fin & (np.abs(x) <= 2.0 ** -27)
        assert np.array_equal(bits(s)[tiny], bits(x)[tiny]) and np.all(c[tiny] == 1.0)


# ---------------------------------------------------------------------------------------------------------------
# derived functions: tan, cot, sinc from one (sin, cos) pair and one division
# ---------------------------------------------------------------------------------------------------------------
def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_derived_functions_match_libm(fabe13, oracle):
    x = np.concatenate([oracle.fill_uniform(200_000, 41, -1e4, 1e4), oracle.fill_uniform(50_000, 42, -3.2, 3.2),
                        oracle.fill_loguniform(50_000, 43)])
    ls, lc = oracle.libm(x)
    t = fabe13.host_derived(x, fabe13.OP_TAN)
    ct = fabe13.host_derived(x, fabe13.OP_COT)
    sc = fabe13.host_derived(x, fabe13.OP_SINC)
    # sin and cos are within 1 ulp each and the division rounds once: a few ulp of relative error, whatever |x|
    assert _rel(t, ls / lc).max() < 1e-15
    assert _rel(ct, lc / ls).max() < 1e-15
    assert np.abs(sc - ls / x).max() < 2.3e-16 and _rel(sc[np.abs(x) < 1e8], (ls / x)[np.abs(x) < 1e8]).max() < 1e-15


def test_derived_functions_guards_and_specials(fabe13):
    qnan = np.uint64(0x7FF8000000000000)
    x = np.array([0.0, -0.0, 5e-10, -9.9e-10, 1e-9, 2.0**-27, 1e-300, 5e-324, np.inf, -np.inf, np.nan, 1e308])
    t = fabe13.host_derived(x, fabe13.OP_TAN)
    ct = fabe13.host_derived(x, fabe13.OP_COT)
    sc = fabe13.host_derived(x, fabe13.OP_SINC)
    # reference guards (fabe13/fabe13.c:264-266): |x| < 1e-9 -> sinc = 1; sin == 0 -> cot = NaN; NaN/Inf in -> NaN out
    assert np.all(sc[:4] == 1.0) and sc[4] == 1.0 and sc[5] == 1.0 and sc[6] == 1.0 and sc[7] == 1.0
    assert bits(ct)[0] == qnan and bits(ct)[1] == qnan
    assert ct[5] == 2.0**27 and ct[6] == 1.0 / 1e-300 and np.isinf(ct[7])        # cot(tiny) = 1/x, overflowing for subnormals
    assert np.array_equal(bits(t[:8]), bits(x[:8]))                         # tan(tiny) = x bit for bit (+-0 included)
    for arr in (t, ct, sc):
        assert np.all(bits(arr)[8:11] == qnan)                              # canonical quiet NaN, sign/payload dropped
    assert abs(t[11] - np.tan(1e308)) <= 4e-16 * abs(np.tan(1e308))


def test_scalar_derived_api_is_the_array_code(fabe13, oracle):
    x = np.concatenate([oracle.fill_uniform(300, 5, -50.0, 50.0), oracle.fill_loguniform(100, 6), [0.0, 1e-10, np.inf]])
    for op, fn in ((fabe13.OP_TAN, fabe13.tan), (fabe13.OP_COT, fabe13.cot), (fabe13.OP_SINC, fabe13.sinc)):
        want = fabe13.host_derived(x, op)
        got = np.array([fn(v) for v in x])
        assert np.array_equal(bits(got), bits(want))


# ---------------------------------------------------------------------------------------------------------------
# FP32 arrays (last item): widen, the FP64 core, round once
# ---------------------------------------------------------------------------------------------------------------
def test_sincosf_is_within_half_an_ulp_and_almost_always_correctly_rounded(fabe13):
    """Our own contract for the FP32 entry (the reference has none): every result within 0.5003 ulp (float) of the true
    value; equal to the correctly rounded float except in a ~3e-5 fraction of the cases (then 1 ulp away)."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1e4, 1e4, 400_000), rng.uniform(-4, 4, 200_000), rng.uniform(-3e13, 3e13, 100_000),
                        np.ldexp(rng.uniform(-2, 2, 200_000), rng.integers(-140, 128, 200_000))]).astype(np.float32)
    xd = x.astype(np.float64)
    true = {"sin": np.sin(xd), "cos": np.cos(xd)}                              # glibc double: <= 1 ulp of double
    for core, budget in ((0, 2e-4), (1, 1e-6)):                                # psi core: the full FP64 routine
        s, c = fabe13.host_sincosf(x, core)
        for got, t in ((s, true["sin"]), (c, true["cos"])):
            want = t.astype(np.float32)
            diff = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64))
            assert diff.max() <= 1 and np.count_nonzero(diff) <= max(3, budget * x.size)
            ulp = np.spacing(np.abs(want)).astype(np.float64)
            assert (np.abs(got.astype(np.float64) - t) / ulp).max() <= 0.5003


def test_sincosf_specials(fabe13):
    x = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 2.0**-27, 3.0e38, -3.0e38], dtype=np.float32)
    x[4:5].view(np.uint32)[:] = 0xFFC01234                                     # negative NaN with a payload
    s, c = fabe13.host_sincosf(x)
    assert np.array_equal(s[:2].view(np.uint32), x[:2].view(np.uint32)) and np.all(c[:2] == 1.0)
    assert np.all(s[2:5].view(np.uint32) == 0x7FC00000) and np.all(c[2:5].view(np.uint32) == 0x7FC00000)
    assert s[5] == x[5] and c[5] == 1.0 and s[6] == x[6] and c[6] == 1.0         # tiny: (x, 1) bit for bit
    big = x[7:].astype(np.float64)
    assert np.array_equal(s[7:], np.sin(big).astype(np.float32)) and np.array_equal(c[7:], np.cos(big).astype(np.float32))


def test_constants_header_is_what_the_generator_produces(tmp_path):
    """fabe13_constants.h is generated (tools/gen_constants.py: Remez fits in 80-digit arithmetic, the 2/pi bit table, the
    pi/2 splits); the committed header must be byte-identical to a fresh run, i.e. no hand-edited constant."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "constants.h"
    subprocess.run([sys.executable, os.path.join(root, "tools", "gen_constants.py"), str(out)], check=True, capture_output=True)
    committed = open(os.path.join(root, "fabe_b200", "csrc", "fabe13_constants.h")).read()
    assert out.read_text() == committed


def test_own_known_answer_vectors(fabe13):
    """tests/golden/fabe13_own_kat_v1.npz pins every result bit of this implementation (sincos with both cores, tan, cot,
    sinc, FP32) on 1717 inputs of every class: a numerics change has to be made on purpose (gen_own_kat.py)."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fabe13_own_kat_v1.npz"))
    x = z["x"].view(np.float64)
    with np.errstate(all="ignore"):
        xf = x.astype(np.float32)
    for core, tag in ((0, "poly"), (1, "psi")):
        s, c, k = fabe13.host_core(x, core)
        assert np.array_equal(bits(s), z[f"{tag}/sin"]) and np.array_equal(bits(c), z[f"{tag}/cos"]) and np.array_equal(k, z[f"{tag}/k"])
        for op, name in ((fabe13.OP_TAN, "tan"), (fabe13.OP_COT, "cot"), (fabe13.OP_SINC, "sinc")):
            assert np.array_equal(bits(fabe13.host_derived(x, op, core)), z[f"{tag}/{name}"]), name
        sf, cf = fabe13.host_sincosf(xf, core)
        assert np.array_equal(sf.view(np.uint32), z[f"{tag}/sinf"]) and np.array_equal(cf.view(np.uint32), z[f"{tag}/cosf"])


def test_hypothesis_edge_biased_floats(fabe13, oracle):
    """hypothesis draws doubles with a bias towards the nasty ones (subnormals, boundaries of binades, huge values, NaN,
    infinities, signed zeros): contract on each -- NaN/Inf -> canonical qNaN twice, |x| <= 2^-27 -> (x, 1) bit for bit,
    everything else within 1.2e-16 of glibc, odd/even symmetry bit-exact, and the derived functions consistent with the pair."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    qnan = np.uint64(0x7FF8000000000000)

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.lists(st.floats(width=64, allow_nan=True, allow_infinity=True), min_size=1, max_size=200))
    def run(values):
        x = np.array(values, dtype=np.float64)
        s, c, _ = fabe13.host_core(x, 0)
        fin = np.isfinite(x)
        assert np.all(bits(s)[~fin] == qnan) and np.all(bits(c)[~fin] == qnan)
        tiny = fin & (np.abs(x) <= 2.0 ** -27)
        assert np.array_equal(bits(s)[tiny], bits(x)[tiny]) and np.all(c[tiny] == 1.0)
        body = fin & ~tiny
        if body.any():
            ls, lc = oracle.libm(x[body])
            assert max(np.abs(s[body] - ls).max(), np.abs(c[body] - lc).max()) <= 1.2e-16
        sn, cn, _ = fabe13.host_core(-x, 0)
        assert np.array_equal(bits(sn)[fin], bits(-s)[fin]) and np.array_equal(bits(cn)[fin], bits(c)[fin])
        t = fabe13.host_derived(x, fabe13.OP_TAN)
        ok = fin & (c != 0)
        with np.errstate(all="ignore"):
            assert np.array_equal(bits(t)[ok], bits(s / c)[ok])   # one IEEE division of the pair, nothing else

    run()


@pytest.mark.parametrize("flavour", ["address+undefined", "thread"])
def test_host_side_is_clean_under_sanitizers(tmp_path, flavour):
    """The host side of the library -- scalar API, the host instantiation of the per-element core over every exponent,
    2e6 random bit patterns, the FP32 and derived variants, option table, shard arithmetic, and eight threads calling
    the scalar API and the option table at once -- built with -fsanitize=address,undefined (-fno-sanitize-recover) and
    with -fsanitize=thread (tools/host_sanitize.cu).  No GPU needed."""
    import os
    import re
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_sanitize")
    if flavour == "thread":
        san = ["-fsanitize=thread"]
    else:
        san = ["-fsanitize=address", "-fsanitize=undefined", "-fno-sanitize-recover=all"]
    san += ["-ffp-contract=off", "-mfma"]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-g", "-std=c++17", "-cudart", "static"]
    for f in san:
        cmd += ["-Xcompiler", f]
    cmd += ["-o", exe] + [os.path.join(root, p) for p in ("tools/host_sanitize.cu", "fabe_b200/csrc/fabe13_kernels.cu",
                                                           "fabe_b200/csrc/fabe13_shim.cu", "fabe_b200/csrc/fabe13_hostcore.cpp")] + ["-lpthread"]
    build = subprocess.run(cmd, capture_output=True, text=True)
    if build.returncode != 0 and re.search(r"lib(a|t|ub)san", build.stderr + build.stdout):
        pytest.skip("sanitizer runtime not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66")
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0, (run.stdout + run.stderr)[-4000:]
    assert "host_sanitize: ok" in run.stdout and "WARNING: ThreadSanitizer" not in run.stderr

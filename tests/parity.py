"""The accuracy contract of BASELINE.json's north_star, as assertions shared by the CPU and GPU parity tests.

  * |x| <= 1e6 : |sin - ref_sin|, |cos - ref_cos| <= 4e-11 against the reference's strict build / its restatement
  * elsewhere  : <= 2e-11 against libm (the reference's own stated bound, README "<= 2e-11 compared to libm")
  * quadrant index k bit-exact with the reference formula for |x| < 2^45; k mod 4 exact beyond (Payne-Hanek range)
  * NaN/Inf -> 0x7FF8000000000000 in both outputs; |x| <= 2^-27 -> (x bit-for-bit, 1.0)
On top of the contract, QUALITY_TOL checks this implementation's own claim (about 2 ulp of 1.0 against libm, for
every finite input up to DBL_MAX).
"""
import numpy as np

TOL_VS_REFERENCE = 4e-11   # north_star, |x| <= 1e6
TOL_VS_LIBM = 2e-11        # north_star / reference README, elsewhere
QUALITY_TOL = 4.5e-16      # this implementation's own bound vs libm (either core)
QNAN = np.uint64(0x7FF8000000000000)
ONE = np.uint64(0x3FF0000000000000)
TINY = 2.0 ** -27
HUGE = 2.0 ** 45


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def check_values(x, s, c, ref_s, ref_c, libm_s, libm_c, quality=True, label=""):
    x = np.asarray(x)
    ax = np.abs(x)
    finite = np.isfinite(x)
    special = ~finite
    tiny = finite & (ax <= TINY)
    body = finite & ~tiny
    # specials and tiny: bit-exact
    assert np.all(bits(s)[special] == QNAN), f"{label}: sin of NaN/Inf is not the canonical qNaN"
    assert np.all(bits(c)[special] == QNAN), f"{label}: cos of NaN/Inf is not the canonical qNaN"
    assert np.array_equal(bits(s)[tiny], bits(x)[tiny]), f"{label}: sin(x) != x bit-for-bit for |x| <= 2^-27"
    assert np.all(bits(c)[tiny] == ONE), f"{label}: cos(x) != 1.0 for |x| <= 2^-27"
    small = body & (ax <= 1e6)
    large = body & (ax > 1e6)
    if small.any():
        e = max(np.abs(s[small] - ref_s[small]).max(), np.abs(c[small] - ref_c[small]).max())
        assert e <= TOL_VS_REFERENCE, f"{label}: |x|<=1e6 differs from the reference by {e:.3e} > {TOL_VS_REFERENCE}"
    if large.any():
        e = max(np.abs(s[large] - libm_s[large]).max(), np.abs(c[large] - libm_c[large]).max())
        assert e <= TOL_VS_LIBM, f"{label}: |x|>1e6 differs from libm by {e:.3e} > {TOL_VS_LIBM}"
    if quality and body.any():
        e = max(np.abs(s[body] - libm_s[body]).max(), np.abs(c[body] - libm_c[body]).max())
        assert e <= QUALITY_TOL, f"{label}: differs from libm by {e:.3e} > {QUALITY_TOL} (own accuracy claim)"


def check_k(x, k, k_ref, q_exact=None, label=""):
    x = np.asarray(x)
    ax = np.abs(x)
    finite = np.isfinite(x)
    cw = finite & (ax < HUGE)
    assert np.array_equal(k[cw], k_ref[cw]), f"{label}: k differs from the reference formula on the Cody-Waite range"
    ph = finite & (ax >= HUGE)
    if q_exact is not None and ph.any():
        assert np.array_equal(k[ph], q_exact[ph]), f"{label}: k mod 4 differs from the exact value (Payne-Hanek range)"
    assert np.all(k[~finite] == 0), f"{label}: k of NaN/Inf should be reported as 0"


def check_against_oracle(x, s, c, k, oracle, quality=True, label=""):
    ref_s, ref_c, ref_k = oracle.sincos(x)
    libm_s, libm_c = oracle.libm(x)
    check_values(x, s, c, ref_s, ref_c, libm_s, libm_c, quality=quality, label=label)
    if k is not None:
        check_k(x, k, ref_k, None, label=label)

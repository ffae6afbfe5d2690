"""The Trimaran kernels replace IEEE x/d by a hoisted reciprocal + FMA residual correction
(trimaran.cu: div_inv).  That is only admissible if it is bit-identical to the division.  The kernel's
argument (comment above div_inv): Brisebarre / Muller / Raina 2004, Theorem 4 -- the 1-mul + 2-FMA sequence is
correctly rounded for every numerator when the divisor's last significand bit is 0 (every divisor on the path is an
integer < 2^52, possibly times 2^-20), absent overflow / underflow; divisors or quotients outside those conditions
take the IEEE division.  Checked here on the device against the hardware division: the plugins' ranges, divisors that
FAIL the condition (they must still be exact, through the fallback), numerators built on rounding boundaries of the
quotient, and non-finite / zero / tiny / huge operands."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_div_by_invariant_is_bit_exact(eng):
    rng = np.random.default_rng(12345)
    n = 4_000_000
    xs, ds = [], []
    # TLP: 100*(utilMillis + podCPU + missing) / capMillis, capMillis integer milli-cores
    d = rng.integers(1, 200_000, n).astype(np.float64)
    xs.append(100 * (rng.random(n) * d + rng.integers(0, 16_000, n))); ds.append(d)
    # TLP second division: t*(100-pred)/(100-t) and (100-t)*pred/t for t in 1..99
    t = rng.integers(1, 100, n).astype(np.float64)
    xs.append(t * (100 - rng.random(n) * 100)); ds.append(100 - t)
    xs.append((100 - t) * rng.random(n) * 100); ds.append(t)
    # LVRB: (usedAvg + req)/capacity, capacity in milli-cores or MiB (x 2^-20 scaled bytes)
    d = rng.integers(1, 1 << 40, n).astype(np.float64) * 2.0**-20
    xs.append(rng.random(n) * d * 1.5 + rng.integers(0, 1 << 35, n) * 2.0**-20); ds.append(d)
    # adversarial divisors / numerators
    d = np.ldexp(np.float64((1 << 53) - 1), rng.integers(-60, 10, n))           # significand all ones
    xs.append(np.ldexp(rng.random(n) + 1, rng.integers(-30, 30, n))); ds.append(d)
    d = np.nextafter(np.ldexp(1.0, rng.integers(-20, 40, n)), np.where(rng.random(n) < 0.5, 0, np.inf))
    xs.append(np.ldexp(rng.random(n) + 1, rng.integers(-30, 30, n))); ds.append(d)
    q = rng.integers(1, 1 << 26, n).astype(np.float64) + 0.5                       # quotients near ties
    d = rng.integers(1, 1 << 26, n).astype(np.float64)
    xs.append(np.nextafter(q * d, np.where(rng.random(n) < 0.5, 0, np.inf))); ds.append(d)
    for x, d in zip(xs, ds):
        assert eng.debug_div_check(x, d) == 0


def test_div_by_invariant_rounding_boundaries_and_fallbacks(eng):
    """numerators placed within a few ulps of d * (m + 1/2 ulp): the quotient sits next to a rounding boundary, which is
    where a non-faithful first estimate would show; plus everything that must take the IEEE fallback."""
    rng = np.random.default_rng(777)
    n = 2_000_000
    # eligible divisors (integers < 2^52, optionally x 2^-20), quotient significands with a long run of ones / zeros
    d = rng.integers(1, 1 << 51, n).astype(np.float64) * np.where(rng.random(n) < 0.5, 1.0, 2.0**-20)
    m = np.ldexp(rng.integers(1 << 52, 1 << 53, n).astype(np.float64), -52)          # quotient in [1, 2)
    half_ulp = 2.0**-53
    for shift in (0.0, half_ulp, -half_ulp):
        x = d * (m + shift)                                                          # RN(d * (m +- ulp/2))
        for k in (-2, -1, 0, 1, 2):
            xx = x.copy()
            for _ in range(abs(k)):
                xx = np.nextafter(xx, np.inf if k > 0 else -np.inf)
            assert eng.debug_div_check(xx, d) == 0
    # small odd divisors and their multiples (exact quotients, and one ulp off)
    d = (2 * rng.integers(0, 50_000, n) + 1).astype(np.float64)
    x = d * rng.integers(0, 1 << 30, n).astype(np.float64)
    assert eng.debug_div_check(x, d) == 0
    assert eng.debug_div_check(np.nextafter(x, np.inf), d) == 0
    # fallbacks: non-finite / zero / tiny / huge numerators, zero / tiny / huge / odd-last-bit divisors
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e-310, 1e-300, 1e300, 1.7e308, 1.0, 3.0, 100.0])
    dd = np.array([1.0, 3.0, 7.0, 40.0, 60.0, 1e-310, 1e-200, 1e200, np.inf, np.nextafter(1.0, 2.0), 0.1, 1 / 3, 2.0**52 + 1])
    X, D = np.meshgrid(special, dd)
    assert eng.debug_div_check(X.ravel().copy(), D.ravel().copy()) == 0

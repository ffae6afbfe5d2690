"""The Trimaran kernels replace IEEE x/d by a hoisted reciprocal + FMA residual correction
(trimaran.cu: div_inv).  That is only admissible if it is bit-identical to the division — checked on
the device against the hardware division over the ranges the plugins produce and over adversarial
divisors (significands of all ones, powers of two +- 1 ulp, quotients that sit on rounding ties)."""
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

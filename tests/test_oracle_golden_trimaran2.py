"""Pins the oracle's Peaks and LowRiskOverCommitment restatement (oracle/trimaran2.c) against the reference's own
unit-test vectors (tests/golden/peaks.json, lowrisk.json) and -- because those vectors never reach the beta
distribution with a non-zero variance -- cross-checks the regularised incomplete beta function against
scipy.special.betainc.  See trimaran2.c's header: bit-identity with gonum / Go's assembly math.Exp is NOT claimed."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


PEAKS = load("peaks.json")
LOWRISK = load("lowrisk.json")


@pytest.mark.parametrize("case", PEAKS["score_cases"], ids=lambda c: c["name"])
def test_peaks_score(oracle, case):
    m = PEAKS["power_model"]["node-1"]
    got = oracle.peaks_score(case["util"], case["cap_milli"], case["flags"], m["k1"], m["k2"], case["pod_cpu_milli"])
    if "expected" in case:
        assert got == case["expected"]
    else:  # peaks_test.go:236-238: the test computes its expectation with the plugin's own helper
        jump = m["k1"] * (oracle.go_exp(m["k2"] * 100) - oracle.go_exp(m["k2"] * 0))
        assert got == int(jump * 1e15) and got > 0
        # libm's exp is within an ulp of Go's portable algorithm: a few units at the 1e15 magnification
        assert abs(got - int(m["k1"] * (math.exp(m["k2"] * 100) - 1.0) * 1e15)) <= 16


@pytest.mark.parametrize("case", PEAKS["normalize_cases"], ids=lambda c: c["name"])
def test_peaks_normalize(oracle, case):
    assert list(oracle.peaks_normalize(case["scores"])) == case["expected"]


def test_peaks_normalize_general(oracle):
    # peaks.go:158-160: 100 - int64(100 * (s - min) / (max - min)), float64 arithmetic, truncation
    s = [10, 250, 1000, 999, 11]
    want = [100 - int(100.0 * (x - 10) / 990.0) for x in s]
    assert list(oracle.peaks_normalize(s)) == want
    assert list(oracle.peaks_normalize([-5, -5])) == [100, 100]  # max == min != 0 -> every node gets 100
    assert list(oracle.peaks_normalize([])) == []


def test_go_exp_matches_libm_within_an_ulp(oracle):
    g = np.random.default_rng(7)
    for x in np.concatenate([g.uniform(-30, 30, 4000), g.uniform(-745, 709, 500), [0.0, 1e-10, -1e-10, 709.78, -745.1]]):
        a, b = oracle.go_exp(float(x)), math.exp(float(x)) if x < 709.78 else math.inf
        assert a == b or abs(a - b) <= 2 * math.ulp(b), x
    assert oracle.go_exp(math.inf) == math.inf and oracle.go_exp(-math.inf) == 0 and math.isnan(oracle.go_exp(math.nan))


@pytest.mark.parametrize("case", LOWRISK["compute_risk_cases"], ids=lambda c: c["name"])
def test_lowrisk_compute_risk(oracle, case):
    cap = case["capacity"]
    cap_f = float(cap) if case["resource"] == "cpu" else float(cap) * (1.0 / 1024.0 / 1024.0)
    got = oracle.lowrisk_compute_risk(True, case["util"], case["std"], cap_f, cap, case["node_req"], case["node_lim"],
                                      case["pod_req"], case["pod_lim"], 5, 0.5)
    assert got == case["expected"]  # the Go test compares with ==


@pytest.mark.parametrize("case", LOWRISK["score_cases"], ids=lambda c: c["name"])
def test_lowrisk_score(oracle, case):
    got = oracle.lowrisk_score(case["cpu_avg"], case["cpu_std"], case["mem_avg"], case["mem_std"],
                               case["alloc_cpu_milli"], case["alloc_mem_bytes"], case["flags"], *case["node"],
                               *case["pod"])
    assert got == case["expected"]


@pytest.mark.parametrize("case", LOWRISK["cdf_cases"], ids=lambda c: c["name"])
def test_lowrisk_beta_cdf_vectors(oracle, case):
    x = case["x"]
    got = 0.0 if x == 0 else 1.0 if x == 1 else oracle.incbet(case["alpha"], case["beta"], x)  # beta.go:162-169
    assert abs(got - case["expected"]) < LOWRISK["tolerance"]


def test_incbet_against_scipy(oracle):
    from scipy.special import betainc

    g = np.random.default_rng(11)
    worst = 0.0
    for _ in range(6000):
        a, b, x = 10 ** g.uniform(-2, 3.5), 10 ** g.uniform(-2, 3.5), g.uniform(0, 1)
        worst = max(worst, abs(oracle.incbet(a, b, x) - betainc(a, b, x)))
    assert worst < 1e-9  # lgamma cancellation at a + b in the thousands (Cephes has the same limit)
    # the shapes the plugin produces: alpha + beta = mu (1 - mu) / var - 1 with var <= 0.99 mu (1 - mu)
    for _ in range(6000):
        mu, frac = g.uniform(0.001, 0.999), 10 ** g.uniform(-6, math.log10(0.99))
        t = 1.0 / frac - 1.0
        a, b, x = mu * t, (1 - mu) * t, g.uniform(0, 1)
        assert abs(oracle.incbet(a, b, x) - betainc(a, b, x)) < 1e-10


def test_lowrisk_risk_load_properties(oracle):
    # more allocated (threshold up) -> P(load <= threshold) up -> risk down; risk stays in [0, 1]
    prev = 1.0
    for req in range(0, 4001, 250):
        r = oracle.lowrisk_risk_load(True, 40.0, 10.0, 4000.0, 4000, req, 10 ** 9, 5)
        assert 0.0 <= r <= 1.0 and r <= prev + 1e-12
        prev = r
    assert oracle.lowrisk_risk_load(False, 40.0, 10.0, 4000.0, 4000, 1000, 2000, 5) == 0.0   # no stats: zero value
    assert oracle.lowrisk_risk_load(True, 0.0, 10.0, 4000.0, 4000, 1000, 2000, 5) == 0.0     # mu == 0 -> prob 1
    assert oracle.lowrisk_risk_load(True, 40.0, 10.0, 4000.0, 4000, 0, 0, 5) == 0.0          # zero over zero :235

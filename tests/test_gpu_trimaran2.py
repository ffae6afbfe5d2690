"""GPU parity: Trimaran Peaks and LowRiskOverCommitment through the C-ABI vs the oracle (oracle/trimaran2.c).

Peaks is bit-exact against the oracle (both restate Go's portable math.Exp; everything else is exact arithmetic).
LowRiskOverCommitment goes through the regularised incomplete beta function, where the device uses CUDA's
log/pow/tgamma/lgamma and the oracle libm's: the per-node risk agrees to ~1e-12, so integer scores are equal except
where 100 * rank sits within TOL of a rounding boundary -- the test allows +-1 there and nowhere else."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from scheduler_plugins_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-9


def snapshot_both(eng, nodes, tri, t2):
    eng.snapshot_begin(nodes["N"])
    eng.snapshot_peaks(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["tlp_flags"], t2["k1"], t2["k2"])
    eng.snapshot_low_risk(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                          nodes["alloc_mem_bytes"], tri["lvrb_flags"], t2["node_req_cpu"], t2["node_req_mem"],
                          t2["node_lim_cpu"], t2["node_lim_mem"])
    eng.snapshot_commit()


@pytest.mark.parametrize("P,N,masked", [(96, 3001, True), (33, 513, False), (1, 1, True), (7, 130, True)])
def test_peaks_matches_oracle(eng, engine_mod, oracle, P, N, masked):
    E = engine_mod
    seed = synth.BASE_SEED + 71
    nodes = synth.gen_nodes(seed, N)
    tri, t2 = synth.gen_trimaran(seed, nodes), synth.gen_trimaran2(seed, nodes, P)
    nodes["cap_cpu_milli"][::89] = 0  # capacity 0: predicted stays 0 (peaks.go:134-137)
    snapshot_both(eng, nodes, tri, t2)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad) if masked else None
    eng.pods_upload(P, feasible=feas, peaks_pod_cpu_milli=t2["peaks_pod_cpu_milli"])
    eng.eval(E.PLUGIN_PEAKS)
    got = eng.fetch_scores(E.PLUGIN_PEAKS)
    want = oracle.peaks_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["tlp_flags"], t2["k1"], t2["k2"],
                              t2["peaks_pod_cpu_milli"], feas, pitch=eng.Npad)
    assert np.array_equal(got, want)
    assert got.min() >= 0 and got.max() <= 100
    if N > 100:
        assert (got == 100).any() and (got == 0).any() and ((got > 0) & (got < 100)).any()
    eng.eval(E.PLUGIN_PEAKS, E.OUT_U8)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_PEAKS, E.OUT_U8).astype(np.int64), want)


def test_peaks_golden_through_cuda(eng, engine_mod, oracle):
    """peaks_test.go:174-426 rows: one node each, raw score visible through a second all-zero node (normalisation
    maps the larger raw score to 0 and the smaller to 100)."""
    E = engine_mod
    g = json.load(open(os.path.join(GOLDEN, "peaks.json")))
    m = g["power_model"]["node-1"]
    for case in g["score_cases"]:
        raw = oracle.peaks_score(case["util"], case["cap_milli"], case["flags"], m["k1"], m["k2"], case["pod_cpu_milli"])
        if "expected" in case:
            assert raw == case["expected"]
        # node 0 = the case; node 1 = a node without metrics (raw 0)
        eng.snapshot_begin(2)
        eng.snapshot_peaks([case["util"], 0.0], [case["cap_milli"], 1000], [case["flags"], 0], [m["k1"], 0.0],
                           [m["k2"], 0.0])
        eng.snapshot_commit()
        eng.pods_upload(1, peaks_pod_cpu_milli=[case["pod_cpu_milli"]])
        eng.eval(E.PLUGIN_PEAKS)
        got = eng.fetch_scores(E.PLUGIN_PEAKS)[0, :2]
        assert list(got) == list(oracle.peaks_normalize([raw, 0])), case["name"]
    for case in g["normalize_cases"]:  # the min == max and all-zero rules, through raw scores the kernel produces
        assert list(oracle.peaks_normalize(case["scores"])) == case["expected"]


def lowrisk_want(oracle, nodes, tri, t2, window, wc, wm, pitch):
    pod = t2["low_risk_pod"]
    return oracle.lowrisk_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                                nodes["alloc_mem_bytes"], tri["lvrb_flags"], t2["node_req_cpu"], t2["node_req_mem"],
                                t2["node_lim_cpu"], t2["node_lim_mem"], pod[0], pod[1], pod[2], pod[3], window, wc, wm,
                                pitch=pitch)


@pytest.mark.parametrize("P,N,window,wc,wm", [(64, 2049, 5, 0.5, 0.5), (33, 777, 1, 0.2, 0.9), (5, 129, 12, 1.0, 0.0)])
def test_low_risk_matches_oracle(eng, engine_mod, oracle, P, N, window, wc, wm):
    E = engine_mod
    seed = synth.BASE_SEED + 72
    nodes = synth.gen_nodes(seed, N)
    tri, t2 = synth.gen_trimaran(seed, nodes), synth.gen_trimaran2(seed, nodes, P)
    nodes["alloc_cpu_milli"][::97] = 0  # capacity 0: NaN thresholds take the reference's NaN paths
    snapshot_both(eng, nodes, tri, t2)
    eng.config_low_risk(window, wc, wm)
    eng.pods_upload(P, low_risk_pod=t2["low_risk_pod"])
    eng.eval(E.PLUGIN_LOW_RISK)
    got = eng.fetch_scores(E.PLUGIN_LOW_RISK)
    want = lowrisk_want(oracle, nodes, tri, t2, window, wc, wm, eng.Npad)
    diff = got != want
    if diff.any():  # only a rounding boundary may differ, and only by one
        assert np.abs(got - want)[diff].max() == 1
        pod = t2["low_risk_pod"]
        for p, n in zip(*np.nonzero(diff)):
            rc = oracle.lowrisk_compute_risk(bool(tri["lvrb_flags"][n] & 2), tri["cpu_avg"][n], tri["cpu_std"][n],
                                             float(nodes["alloc_cpu_milli"][n]), int(nodes["alloc_cpu_milli"][n]),
                                             int(t2["node_req_cpu"][n]), int(t2["node_lim_cpu"][n]), int(pod[0][p]),
                                             int(pod[2][p]), window, wc)
            rm = oracle.lowrisk_compute_risk(bool(tri["lvrb_flags"][n] & 4), tri["mem_avg"][n], tri["mem_std"][n],
                                             float(nodes["alloc_mem_bytes"][n]) * (1.0 / 1024.0 / 1024.0),
                                             int(nodes["alloc_mem_bytes"][n]), int(t2["node_req_mem"][n]),
                                             int(t2["node_lim_mem"][n]), int(pod[1][p]), int(pod[3][p]), window, wm)
            x = (1 - max(rc, rm)) * 100.0
            assert abs(x - np.floor(x) - 0.5) < TOL, (p, n, x)
    assert diff.mean() < 1e-4
    assert got.min() >= 0 and got.max() <= 100 and len(np.unique(got)) > 10
    eng.eval(E.PLUGIN_LOW_RISK, E.OUT_U8)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_LOW_RISK, E.OUT_U8).astype(np.int64), got)


def test_low_risk_golden_through_cuda(eng, engine_mod):
    """lowriskovercommitment_test.go:342-395 (computeRisk, node_A) as scores: rank = 1 - max(riskCPU, riskMemory)."""
    E = engine_mod
    g = json.load(open(os.path.join(GOLDEN, "lowrisk.json")))
    by = {c["name"]: c for c in g["compute_risk_cases"]}
    for cpu, mem in (("test-cpu-1", "test-mem-1"), ("test-cpu-2", "test-mem-2")):
        c, m = by[cpu], by[mem]
        eng.snapshot_begin(1)
        eng.snapshot_low_risk([c["util"]], [c["std"]], [m["util"]], [m["std"]], [c["capacity"]], [m["capacity"]], [7],
                              [c["node_req"]], [m["node_req"]], [c["node_lim"]], [m["node_lim"]])
        eng.snapshot_commit()
        eng.config_low_risk(5, 0.5, 0.5)
        eng.pods_upload(1, low_risk_pod=[[c["pod_req"]], [m["pod_req"]], [c["pod_lim"]], [m["pod_lim"]]])
        eng.eval(E.PLUGIN_LOW_RISK)
        want = int(np.floor((1 - max(c["expected"], m["expected"])) * 100 + 0.5))
        assert eng.fetch_scores(E.PLUGIN_LOW_RISK)[0, 0] == want
    case = g["score_cases"][0]  # best-effort pod -> MinNodeScore
    eng.snapshot_begin(1)
    eng.snapshot_low_risk([case["cpu_avg"]], [0.0], [0.0], [0.0], [1000], [1 << 30], [case["flags"]], [0], [0], [0], [0])
    eng.snapshot_commit()
    eng.pods_upload(1, low_risk_pod=[[0], [0], [0], [0]])
    eng.eval(E.PLUGIN_LOW_RISK)
    assert eng.fetch_scores(E.PLUGIN_LOW_RISK)[0, 0] == case["expected"]


def test_trimaran2_state_errors(eng, engine_mod):
    E = engine_mod
    nodes = synth.gen_nodes(synth.BASE_SEED, 200)
    eng.snapshot_begin(200)
    eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
    eng.snapshot_commit()
    eng.pods_upload(2, peaks_pod_cpu_milli=[0, 1], low_risk_pod=np.zeros((4, 2), np.int64))
    for pl in (E.PLUGIN_PEAKS, E.PLUGIN_LOW_RISK):
        with pytest.raises(E.B200SError):  # no such columns in the snapshot
            eng.eval(pl)
    with pytest.raises(E.B200SError):
        eng.config_low_risk(0, 0.5, 0.5)
    with pytest.raises(E.B200SError):
        eng.config_low_risk(5, 1.5, 0.5)

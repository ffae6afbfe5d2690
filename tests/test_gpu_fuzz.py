"""Differential fuzz: random small shapes (ragged N, tiny P, sparse/dense/empty feasibility, every
plugin and the combined profile) — CUDA through the C-ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth
from test_gpu_combined import build_inputs, load_engine, oracle_combined

pytestmark = pytest.mark.gpu


def random_mask(rng, E, P, N, npad):
    kind = rng.integers(0, 5)
    if kind == 0:
        return None
    dens = [0.0, 0.02, 0.5, 0.98][kind - 1]
    m = rng.random((P, N)) < dens
    if kind == 1 and P > 1:
        m[rng.integers(0, P)] = True  # one fully feasible row among empty ones
    return E.pack_bits(m, npad)


@pytest.mark.parametrize("it", range(24))
def test_fuzz_all_plugins(eng, engine_mod, oracle, it):
    from oracle import pyoracle_nrt

    E = engine_mod
    rng = np.random.default_rng(1000 + it)
    P, N = int(rng.integers(1, 40)), int(rng.integers(1, 700))
    Z = int(rng.choice([1, 2, 4, 8]))
    d = build_inputs(5000 + it, P, N, Z=Z)
    npad = E.npad_of(N)
    feas = random_mask(rng, E, P, N, npad)
    strategy = int(rng.integers(0, 4))
    load_engine(eng, E, d, N, P, feas, nrt_strategy=strategy)
    nodes, pods, tri, net = d["nodes"], d["pods"], d["tri"], d["net"]
    # --- each plugin alone
    eng.eval(E.PLUGIN_ALLOCATABLE)
    want = oracle.alloc_batch([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]], [1 << 20, 1], 1, P, feas, pitch=npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_ALLOCATABLE), want)
    eng.eval(E.PLUGIN_TLP)
    want = oracle.tlp_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"],
                            pods["tlp_pod_cpu_milli"], 40, pitch=npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_TLP), want)
    eng.eval(E.PLUGIN_LVRB)
    want = oracle.lvrb_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                             nodes["alloc_mem_bytes"], tri["lvrb_flags"], pods["req_cpu_milli"], pods["req_mem_bytes"],
                             1.0, 1.0, pitch=npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_LVRB), want)
    eng.eval(E.PLUGIN_NRT)
    ws, wf, wr = pyoracle_nrt.nrt_batch(d["nrt_nodes"], d["nrt_pods"], strategy, [1, 1, 1, 1], feas, pitch=npad)
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), wr)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NRT), wf)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NRT), ws)
    for counts in (False, True):  # pair-table path and the materialising path
        eng.config_network_overhead(want_counts=counts)
        eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
        ws, wf, wr = oracle.netoh_batch(net["zone_cost"], net["region_cost"], net["region_all"], net["zone_all"],
                                        net["score_equally"], net["dep_offset"], net["deps"], feas, pitch=npad)
        assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NETWORK_OVERHEAD), wf)
        assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NETWORK_OVERHEAD), wr)
        assert np.array_equal(eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD), ws)
    eng.config_network_overhead(want_counts=False)
    # --- the combined cycle
    k = int(rng.integers(1, 6))
    mask = int(rng.integers(1, 32))
    weights = [int(x) for x in rng.integers(1, 7, 5)]
    eng.eval_combined(mask, weights, k=k, write_total=True)
    wt, wfe, wk = oracle_combined(d, P, N, npad, feas, weights, k, mask, nrt_strategy=strategy)
    assert np.array_equal(eng.fetch_total_feasible(), wfe)
    assert np.array_equal(eng.fetch_total(), wt)
    got = eng.fetch_topk()
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got[p]] == wk[p], (p, mask, k)

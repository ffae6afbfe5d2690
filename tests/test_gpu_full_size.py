"""GPU parity at BASELINE.json's full sizes for the configs that round 1 only covered at <= 96 x 3 001: config 3
(TargetLoadPacking + LoadVariationRiskBalancing, 10k pods x 50k nodes) and the per-GPU shard of config 5 (all five
plugins, 50k x 200k over 8 GPUs = 25k nodes per GPU; here one 5k-pod chunk of it).  The oracle checks sampled pods
bit-exactly; size-independent properties cover the whole result.  (Config 2: test_gpu_parity.py; config 4:
test_gpu_nrt_batched.py.)"""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth

pytestmark = pytest.mark.gpu


def test_config3_full_size(eng, engine_mod, oracle):
    E = engine_mod
    P, N = 10_000, 50_000
    seed = synth.BASE_SEED + 3
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    eng.snapshot_begin(N)
    eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
    eng.snapshot_lvrb(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                      nodes["alloc_mem_bytes"], tri["lvrb_flags"])
    eng.snapshot_commit()
    eng.config_tlp(40)
    eng.config_lvrb(1.0, 1.0)
    eng.pods_upload(P, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=pods["req_cpu_milli"],
                    lvrb_req_mem_bytes=pods["req_mem_bytes"])
    rows = np.sort(np.random.default_rng(3).choice(P, 24, replace=False))
    want_tlp = oracle.tlp_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"],
                                pods["tlp_pod_cpu_milli"][rows], 40, pitch=eng.Npad)
    want_lvrb = oracle.lvrb_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                                  nodes["alloc_mem_bytes"], tri["lvrb_flags"], pods["req_cpu_milli"][rows],
                                  pods["req_mem_bytes"][rows], 1.0, 1.0, pitch=eng.Npad)
    for plugin, want in ((E.PLUGIN_TLP, want_tlp), (E.PLUGIN_LVRB, want_lvrb)):
        eng.eval(plugin, E.OUT_U8)
        got8 = eng.fetch_scores(plugin, E.OUT_U8)
        assert np.array_equal(got8[rows].astype(np.int64), want)
        assert got8.max() <= 100 and not got8[:, N:].any()
        # pods with the same request columns get the same row (the scores depend on the pod only through them)
        key = pods["tlp_pod_cpu_milli"] if plugin == E.PLUGIN_TLP else pods["req_cpu_milli"] * (1 << 40) + pods["req_mem_bytes"]
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        probe = np.random.default_rng(5).choice(P, 200, replace=False)
        assert all(np.array_equal(got8[p], got8[first[inv[p]]]) for p in probe)
        eng.eval(plugin, E.OUT_I64)
        got64 = eng.fetch_scores(plugin, E.OUT_I64)
        assert np.array_equal(got64[rows], want)
        assert int(got64.sum()) == int(got8.astype(np.int64).sum())  # both transports agree on the whole matrix
    # nodes without metrics score 0 for every pod (targetloadpacking.go:114-120, loadvariationriskbalancing.go:91-94)
    assert not got64[:, :N][:, tri["lvrb_flags"] == 0].any()


def test_config5_shard_chunk_full_size(eng, engine_mod, oracle):
    """One 5 000-pod chunk of config 5 on one 25 000-node shard: every plugin's matrix on sampled pods, the weighted
    total and the per-pod top-1 against the oracle's restatement of the upstream cycle."""
    from test_gpu_combined import build_inputs, load_engine, oracle_combined

    E = engine_mod
    P, N = 5_000, 25_000
    seed = synth.BASE_SEED + 5
    d = build_inputs(seed, P, N)
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    load_engine(eng, E, d, N, P, feas)
    weights = [1, 1, 1, 1, 5]
    eng.eval_combined(0b11111, weights, k=1, write_total=True)
    assert eng.nrt_last_path() == E.NRT_PATH_BATCHED, eng.nrt_path_note()
    topk, total, tfeas = eng.fetch_topk(), eng.fetch_total(), eng.fetch_total_feasible()
    rows = np.sort(np.random.default_rng(9).choice(P, 12, replace=False))
    sub = dict(d, pods={k: (v[rows] if isinstance(v, np.ndarray) else v) for k, v in d["pods"].items()},
               nrt_pods={k: (v[rows] if isinstance(v, np.ndarray) else v) for k, v in d["nrt_pods"].items()})
    net = d["net"]
    offs, deps = [0], []
    for p in rows:
        a, b = int(net["dep_offset"][p]), int(net["dep_offset"][p + 1])
        deps.append(net["deps"][a:b])
        offs.append(offs[-1] + b - a)
    sub["net"] = dict(net, score_equally=net["score_equally"][rows], dep_offset=np.array(offs, dtype=np.int32),
                      deps=np.concatenate(deps) if offs[-1] else net["deps"][:0])
    want_total, want_feas, want_topk = oracle_combined(sub, len(rows), N, eng.Npad, feas[rows], weights, 1, 0b11111)
    assert np.array_equal(total[rows], want_total)
    assert np.array_equal(tfeas[rows], want_feas)
    for i, p in enumerate(rows):
        assert (int(topk[p][0]["score"]), int(topk[p][0]["node"])) == tuple(want_topk[i][0]), p
    # whole result: the winner of every pod is a feasible node carrying the row's maximum total, lowest index first
    bits = E.unpack_bits(tfeas, N)
    tot = np.where(bits, total[:, :N], -1)
    best = tot.max(axis=1)
    has = bits.any(axis=1)
    nodes_won = topk["node"][:, 0]
    assert np.array_equal(nodes_won[~has], np.full((~has).sum(), -1))
    assert np.array_equal(topk["score"][has, 0], best[has])
    assert np.array_equal(nodes_won[has], np.argmax(tot[has] == best[has, None], axis=1))
    assert not total[:, :N][~bits].any()


@pytest.mark.parametrize("shared", [True, False])
def test_trimaran_score_table_path(eng, engine_mod, oracle, shared):
    """TargetLoadPacking / LoadVariationRiskBalancing on a batch whose pods share a few request keys go through one
    table row per distinct key + a streaming expansion (2 launches); a batch of all-distinct keys, or a snapshot with a
    non-finite metric, keeps the direct kernel (1 launch).  Same scores either way."""
    E = engine_mod
    P, N = 700, 2100
    seed = 4242
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    if not shared:
        pods = dict(pods, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"] + np.arange(P), req_cpu_milli=pods["req_cpu_milli"] + np.arange(P))
    for poison in (False, True):
        t = dict(tri, cpu_avg=tri["cpu_avg"].copy())
        if poison:
            t["cpu_avg"][17] = np.nan  # Go: NaN propagates to MinInt64 in the int64 score -- not a byte-table value
        eng.snapshot_begin(N)
        eng.snapshot_tlp(t["cpu_avg"], nodes["cap_cpu_milli"], t["missing_milli"], t["tlp_flags"])
        eng.snapshot_lvrb(t["cpu_avg"], t["cpu_std"], t["mem_avg"], t["mem_std"], nodes["alloc_cpu_milli"],
                          nodes["alloc_mem_bytes"], t["lvrb_flags"])
        eng.snapshot_commit()
        eng.config_tlp(40)
        eng.config_lvrb(1.0, 1.0)
        eng.pods_upload(P, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=pods["req_cpu_milli"],
                        lvrb_req_mem_bytes=pods["req_mem_bytes"])
        want = {E.PLUGIN_TLP: oracle.tlp_batch(t["cpu_avg"], nodes["cap_cpu_milli"], t["missing_milli"], t["tlp_flags"],
                                               pods["tlp_pod_cpu_milli"], 40, pitch=eng.Npad),
                E.PLUGIN_LVRB: oracle.lvrb_batch(t["cpu_avg"], t["cpu_std"], t["mem_avg"], t["mem_std"], nodes["alloc_cpu_milli"],
                                                 nodes["alloc_mem_bytes"], t["lvrb_flags"], pods["req_cpu_milli"],
                                                 pods["req_mem_bytes"], 1.0, 1.0, pitch=eng.Npad)}
        for plugin in (E.PLUGIN_TLP, E.PLUGIN_LVRB):
            l0 = eng.launches
            eng.eval(plugin, E.OUT_I64)
            assert eng.launches - l0 == (2 if shared and not poison else 1)
            assert np.array_equal(eng.fetch_scores(plugin), want[plugin])
            if not poison:
                eng.eval(plugin, E.OUT_U8)
                assert np.array_equal(eng.fetch_scores(plugin, E.OUT_U8).astype(np.int64), want[plugin])

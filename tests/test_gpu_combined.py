"""GPU parity: the combined profile (all five plugins, weighted sum, per-pod top-k) vs the oracle's
restatement of the upstream cycle; single GPU here, the sharded path in test_multi_gpu.py."""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth

pytestmark = pytest.mark.gpu


def build_inputs(seed, P, N, Z=4):
    nodes = synth.gen_nodes(seed, N)
    pods = synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    nrt_nodes, nrt_pods = synth.gen_nrt(seed, N, P, Z=Z)
    net = synth.gen_netoh(seed, N, P)
    return dict(nodes=nodes, pods=pods, tri=tri, nrt_nodes=nrt_nodes, nrt_pods=nrt_pods, net=net)


def load_engine(eng, E, d, N, P, feas, node_offset=0, n_global=None, nrt_strategy=2):
    nodes, pods, tri, net = d["nodes"], d["pods"], d["tri"], d["net"]
    sl = slice(node_offset, node_offset + N)
    eng.snapshot_begin(N, node_offset=node_offset, n_nodes_global=n_global or (node_offset + N))
    eng.snapshot_allocatable([nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl]])
    eng.snapshot_tlp(tri["cpu_avg"][sl], nodes["cap_cpu_milli"][sl], tri["missing_milli"][sl], tri["tlp_flags"][sl])
    eng.snapshot_lvrb(tri["cpu_avg"][sl], tri["cpu_std"][sl], tri["mem_avg"][sl], tri["mem_std"][sl],
                      nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl], tri["lvrb_flags"][sl])
    nn = d["nrt_nodes"]
    eng.snapshot_nrt(dict(nn, node_flags=nn["node_flags"][sl], max_numa=nn["max_numa"][sl],
                          n_zones_node=nn["n_zones_node"][sl], node_res_mask=nn["node_res_mask"][sl],
                          zone_res_mask=nn["zone_res_mask"][:, sl], avail=nn["avail"][:, :, sl],
                          cost=nn["cost"][:, :, sl]))
    eng.snapshot_network_overhead(net["region_all"][sl], net["zone_all"][sl], net["zone_cost"], net["region_cost"])
    eng.snapshot_commit()
    eng.config_allocatable(E.ALLOC_MOST, [1 << 20, 1])
    eng.config_tlp(40)
    eng.config_lvrb(1.0, 1.0)
    eng.config_nrt(nrt_strategy, [1, 1, 1, 1])
    eng.pods_upload(P, feasible=feas, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"],
                    lvrb_req_cpu_milli=pods["req_cpu_milli"], lvrb_req_mem_bytes=pods["req_mem_bytes"],
                    nrt=d["nrt_pods"], netoh=net)


def oracle_combined(d, P, N, pitch, feas, weights, k, mask, node_offset=0, nrt_strategy=2):
    from oracle import combined as OC

    nodes, pods, tri, net = d["nodes"], d["pods"], d["tri"], d["net"]
    sl = slice(node_offset, node_offset + N)
    nn = d["nrt_nodes"]
    nrt_nodes = dict(nn, node_flags=nn["node_flags"][sl], max_numa=nn["max_numa"][sl], n_zones_node=nn["n_zones_node"][sl],
                     node_res_mask=nn["node_res_mask"][sl], zone_res_mask=nn["zone_res_mask"][:, sl],
                     avail=nn["avail"][:, :, sl], cost=nn["cost"][:, :, sl])
    kw = {}
    if mask & 1:
        kw["alloc"] = dict(cols=[nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl]], weights=[1 << 20, 1], mode=1)
    if mask & 2:
        kw["tlp"] = dict(util=tri["cpu_avg"][sl], cap=nodes["cap_cpu_milli"][sl], missing=tri["missing_milli"][sl],
                         flags=tri["tlp_flags"][sl], pod_cpu=pods["tlp_pod_cpu_milli"], target=40)
    if mask & 4:
        kw["lvrb"] = dict(node_cols=[tri["cpu_avg"][sl], tri["cpu_std"][sl], tri["mem_avg"][sl], tri["mem_std"][sl],
                                     nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl], tri["lvrb_flags"][sl]],
                          req_cpu=pods["req_cpu_milli"], req_mem=pods["req_mem_bytes"], margin=1.0, sens=1.0)
    if mask & 8:
        kw["nrt"] = dict(nodes=nrt_nodes, pods=d["nrt_pods"], strategy=nrt_strategy, weights=[1, 1, 1, 1])
    if mask & 16:
        kw["netoh"] = dict(zone_cost=net["zone_cost"], region_cost=net["region_cost"], region_id=net["region_all"][sl],
                           zone_id=net["zone_all"][sl], score_equally=net["score_equally"], dep_offset=net["dep_offset"],
                           deps=net["deps"])
    return OC.combined(P, N, pitch, feas, weights, k, node_offset=node_offset, **kw)


@pytest.mark.parametrize("mask,k", [(0b11111, 1), (0b11111, 4), (0b00111, 1), (0b01001, 3), (0b10001, 16), (0b00010, 2)])
def test_combined_matches_oracle(eng, engine_mod, mask, k):
    E = engine_mod
    P, N = 48, 1300
    seed = synth.BASE_SEED + 5
    d = build_inputs(seed, P, N)
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    load_engine(eng, E, d, N, P, feas)
    weights = [2, 1, 1, 3, 5]  # profile weights are launch parameters (NetworkOverhead weight 5 in the shipped profile)
    eng.eval_combined(mask, weights, k=k, write_total=True)
    got_topk = eng.fetch_topk()
    got_total = eng.fetch_total()
    got_feas = eng.fetch_total_feasible()
    want_total, want_feas, want_topk = oracle_combined(d, P, N, eng.Npad, feas, weights, k, mask)
    assert np.array_equal(got_feas, want_feas)
    assert np.array_equal(got_total, want_total)
    for p in range(P):
        got = [(int(e["score"]), int(e["node"])) for e in got_topk[p]]
        assert got == want_topk[p], (p, got, want_topk[p])
    assert any(r[0][1] >= 0 for r in want_topk)
    # winners only (no total matrix): k = 1 with small weights is the 16-nodes-per-thread read-stream kernel, weights
    # beyond 32-bit sums keep the general one -- same winners either way
    eng.eval_combined(mask, weights, k=k, write_total=False)
    got_topk = eng.fetch_topk()
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got_topk[p]] == want_topk[p], p
    big = [w << 21 for w in weights]
    _, _, want_big = oracle_combined(d, P, N, eng.Npad, feas, big, k, mask)
    eng.eval_combined(mask, big, k=k, write_total=False)
    got_topk = eng.fetch_topk()
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got_topk[p]] == want_big[p], p


def test_combined_without_total_matrix_and_no_upstream_mask(eng, engine_mod):
    E = engine_mod
    P, N = 20, 777
    d = build_inputs(77, P, N, Z=2)
    load_engine(eng, E, d, N, P, None, nrt_strategy=3)
    weights = [1, 1, 1, 1, 1]
    eng.eval_combined(0b11111, weights, k=2, write_total=False)
    got = eng.fetch_topk()
    with pytest.raises(E.B200SError):
        eng.fetch_total()
    _, _, want = oracle_combined(d, P, N, eng.Npad, None, weights, 2, 0b11111, nrt_strategy=3)
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got[p]] == want[p]


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("mask,k,strategy", [(0b11111, 1, 2), (0b11111, 4, 0), (0b11111, 16, 1), (0b01001, 3, 2), (0b10110, 2, 2),
                                             (0b00001, 1, 2), (0b10000, 5, 2), (0b01000, 2, 1)])
def test_fused_cycle_matches_oracle_and_the_plugin_by_plugin_path(eng, engine_mod, P, mask, k, strategy):
    """cycle.cu: the whole cycle of a handful of pods in two launches -- the same winners and the same final
    feasible set as the oracle's restatement of the upstream cycle and as the plugin-by-plugin path."""
    E = engine_mod
    N = 3000 + 37 * P
    seed = synth.BASE_SEED + 5 + P
    d = build_inputs(seed, P, N)
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N)) if mask != 0b00001 else None
    load_engine(eng, E, d, N, P, feas, nrt_strategy=strategy)
    weights = [2, 1, 1, 3, 5]
    launches0 = eng.launches
    eng.eval_combined(mask, weights, k=k, write_total=False)
    assert eng.launches - launches0 <= 5  # three small launches (+ the snapshot-time Allocatable raw / sort on first use)
    got, got_feas = eng.fetch_topk(), eng.fetch_total_feasible()
    _, want_feas, want = oracle_combined(d, P, N, eng.Npad, feas, weights, k, mask, nrt_strategy=strategy)
    assert np.array_equal(got_feas, want_feas)
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got[p]] == want[p], p
    eng.config_fused_cycle(False)
    eng.eval_combined(mask, weights, k=k, write_total=False)
    slow = eng.fetch_topk()
    eng.config_fused_cycle(True)
    assert np.array_equal(slow["score"], got["score"]) and np.array_equal(slow["node"], got["node"])
    # the one-call form: upload + cycle + winners to host with a single synchronisation
    batch, keep = eng.make_batch(P, feasible=feas, tlp_pod_cpu_milli=d["pods"]["tlp_pod_cpu_milli"],
                                 lvrb_req_cpu_milli=d["pods"]["req_cpu_milli"], lvrb_req_mem_bytes=d["pods"]["req_mem_bytes"],
                                 nrt=d["nrt_pods"], netoh=d["net"])
    one = eng.schedule_batch(batch, mask, weights, k=k)
    assert np.array_equal(one["score"], got["score"]) and np.array_equal(one["node"], got["node"])
    assert np.array_equal(eng.fetch_total_feasible(), want_feas)
    # ... which is one graph launch for <= 4 pods: re-launched as is, patched when arguments change, and identical to the
    # same two kernels issued as plain launches
    again = eng.schedule_batch(batch, mask, weights, k=k)
    assert np.array_equal(again["score"], got["score"]) and np.array_equal(again["node"], got["node"])
    w2 = [1, 4, 2, 1, 1]
    _, _, want2 = oracle_combined(d, P, N, eng.Npad, feas, w2, k, mask, nrt_strategy=strategy)
    other = eng.schedule_batch(batch, mask, w2, k=k)
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in other[p]] == want2[p], p
    eng.config_fused_cycle(2)
    plain = eng.schedule_batch(batch, mask, w2, k=k)
    eng.config_fused_cycle(True)
    assert np.array_equal(plain["score"], other["score"]) and np.array_equal(plain["node"], other["node"])


@pytest.mark.parametrize("mask,strategy", [(0b11111, 2), (0b01010, 0), (0b01000, 1)])
def test_schedule_sequence_with_on_device_assume(eng, engine_mod, mask, strategy):
    """SURVEY.md §8f-4, second half: a batch is placed pod by pod WITHOUT leaving the device -- after each cycle the
    winner is assumed (OverReserve deduction on every listing zone of the winner node, store.go:129-160;
    TargetLoadPacking's missing utilisation, handler.go:131-167) and the next pod sees it.  Checked against the same
    loop on the host: oracle cycle of one pod, then the oracle's own deduction, then the next pod."""
    import ctypes as C

    from oracle import pyoracle as orc

    E = engine_mod
    P, N = 14, 1400
    seed = synth.BASE_SEED + 9
    d = build_inputs(seed, P, N)
    # make the assume matter: every pod is Guaranteed with a sizeable request, few roomy nodes
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    load_engine(eng, E, d, N, P, feas, nrt_strategy=strategy)
    weights = [2, 1, 1, 3, 5]
    batch, keep = eng.make_batch(P, feasible=feas, tlp_pod_cpu_milli=d["pods"]["tlp_pod_cpu_milli"],
                                 lvrb_req_cpu_milli=d["pods"]["req_cpu_milli"], lvrb_req_mem_bytes=d["pods"]["req_mem_bytes"],
                                 nrt=d["nrt_pods"], netoh=d["net"])
    got = eng.schedule_sequence(batch, mask, weights)
    # ---- the same loop on the host
    state = dict(d, nrt_nodes=dict(d["nrt_nodes"], avail=d["nrt_nodes"]["avail"].copy()),
                 tri=dict(d["tri"], missing_milli=d["tri"]["missing_milli"].copy()))
    Z, R = state["nrt_nodes"]["n_zones"], state["nrt_nodes"]["n_res"]
    net = d["net"]
    placed_on_a_deducted_node = 0
    touched = set()
    for p in range(P):
        one = dict(state, pods={k: (v[p:p + 1] if isinstance(v, np.ndarray) else v) for k, v in d["pods"].items()},
                   nrt_pods={k: (v[p:p + 1] if isinstance(v, np.ndarray) else v) for k, v in d["nrt_pods"].items()})
        a, b = int(net["dep_offset"][p]), int(net["dep_offset"][p + 1])
        one["net"] = dict(net, score_equally=net["score_equally"][p:p + 1], dep_offset=np.array([0, b - a], dtype=np.int32),
                          deps=net["deps"][a:b])
        _, _, topk = oracle_combined(one, 1, N, eng.Npad, feas[p:p + 1], weights, 1, mask, nrt_strategy=strategy)
        want = topk[0][0]
        assert (int(got[p]["score"]), int(got[p]["node"])) == tuple(want), (p, got[p], want)
        n = int(want[1])
        if n < 0:
            continue
        placed_on_a_deducted_node += n in touched
        touched.add(n)
        if mask & 8:  # OverReserve: the pod's effective request comes off every listing zone of node n
            req, rm = d["nrt_pods"]["req"][p, 8], int(d["nrt_pods"]["req_mask"][p, 8])
            av = np.ascontiguousarray(state["nrt_nodes"]["avail"][:, :, n])
            zm = np.ascontiguousarray(state["nrt_nodes"]["zone_res_mask"][:, n])
            ded = np.ascontiguousarray(req.astype(np.int64))
            orc.lib().orc_nrt_overreserve_deduct(C.c_void_p(av.ctypes.data), C.c_void_p(zm.ctypes.data), C.c_int(Z), C.c_int(R),
                                                 C.c_uint8(rm), C.c_void_p(ded.ctypes.data))
            state["nrt_nodes"]["avail"][:, :, n] = av
        if mask & 2:
            state["tri"]["missing_milli"][n] += d["pods"]["tlp_pod_cpu_milli"][p]
    # the engine's resident columns are the host loop's final state: a plain evaluation afterwards agrees
    if mask & 8:
        from oracle import pyoracle_nrt

        eng.pods_upload(P, feasible=feas, nrt=d["nrt_pods"])
        eng.eval(E.PLUGIN_NRT)
        ws, wf, wr = pyoracle_nrt.nrt_batch(state["nrt_nodes"], d["nrt_pods"], strategy, [1, 1, 1, 1], feas, pitch=eng.Npad)
        assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), wr) and np.array_equal(eng.fetch_scores(E.PLUGIN_NRT), ws)

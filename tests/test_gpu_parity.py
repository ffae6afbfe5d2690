"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the CPU
oracle on the same seeded inputs.  Bit-exact for every int64 score and every mask word."""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth

pytestmark = pytest.mark.gpu

W_DEFAULT = [1 << 20, 1]  # cpu, memory (resource_allocation.go:36)


def setup_alloc(eng, E, nodes, mode, weights=W_DEFAULT, extra_cols=()):
    cols = [nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"], *extra_cols]
    eng.snapshot_begin(nodes["N"])
    eng.snapshot_allocatable(cols)
    eng.snapshot_commit()
    eng.config_allocatable(mode, weights)
    return cols


@pytest.mark.parametrize("P,N,mode,masked", [
    (128, 1000, 0, False),   # config c1: 128 x 1000, Least
    (128, 1000, 0, True),
    (37, 4097, 1, True),     # ragged: N not a multiple of 128, odd P
    (257, 513, 1, False),
    (1, 1, 0, True),         # single node
    (3, 130, 1, True),
])
def test_allocatable_matches_oracle(eng, engine_mod, oracle, P, N, mode, masked):
    E = engine_mod
    seed = synth.BASE_SEED + 1 + mode
    nodes = synth.gen_nodes(seed, N)
    cols = setup_alloc(eng, E, nodes, mode)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad) if masked else None
    eng.pods_upload(P, feasible=feas)
    eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_I64)
    got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)
    want = oracle.alloc_batch(cols, W_DEFAULT, mode, P, feas, pitch=eng.Npad)
    assert np.array_equal(got, want)
    assert got.min() >= 0 and got.max() <= 100
    # compact u8 transport carries the same values
    eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
    got8 = eng.fetch_scores(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
    assert np.array_equal(got8.astype(np.int64), want)


def test_async_upload_of_consecutive_pod_chunks(eng, engine_mod, oracle):
    """b200s_config_async_upload: pod chunks are queued back to back without synchronising -- small columns through the
    double-buffered staging block, the upstream mask (> 4 MiB here) on its own stream into alternating buffers.  Every
    chunk must be evaluated with ITS mask (NormalizeScore depends on it)."""
    E = engine_mod
    P, N = 640, 60_000  # 640 x 60 032 / 8 = 4.8 MB of mask per chunk
    seed = synth.BASE_SEED + 3
    nodes = synth.gen_nodes(seed, N)
    cols = setup_alloc(eng, E, nodes, 1)
    masks = [synth.gen_feasible_words(seed + i, P, N, eng.Npad) for i in range(5)]
    pinned = []
    for m in masks:
        buf = eng.pinned(m.nbytes)
        v = buf.view(np.uint64, m.shape)
        v[:] = m
        pinned.append((buf, v))
    eng.config_async_upload(True)
    try:
        outs = []
        for i, (_, v) in enumerate(pinned):
            eng.pods_upload(P, feasible=v)
            eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
            if i % 2 == 1 or i == len(pinned) - 1:  # chunks 0 and 2 are overtaken by the next upload before any fetch
                outs.append((i, eng.fetch_scores(E.PLUGIN_ALLOCATABLE, E.OUT_U8)))
        for i, got in outs:
            want = oracle.alloc_batch(cols, W_DEFAULT, 1, P, masks[i], pitch=eng.Npad)
            assert np.array_equal(got.astype(np.int64), want), i
    finally:
        eng.config_async_upload(False)
    eng.pods_upload(P, feasible=masks[0])
    eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
    want = oracle.alloc_batch(cols, W_DEFAULT, 1, P, masks[0], pitch=eng.Npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_ALLOCATABLE, E.OUT_U8).astype(np.int64), want)


def test_allocatable_golden_through_cuda(eng, engine_mod):
    """The reference's own table (allocatable_test.go:114-221) through the CUDA path."""
    import json
    import os

    from conftest import GOLDEN

    E = engine_mod
    g = json.load(open(os.path.join(GOLDEN, "allocatable.json")))
    for case in g["cases"]:
        nodes = case["nodes"]
        eng.snapshot_begin(len(nodes))
        eng.snapshot_allocatable([[n[0] for n in nodes], [n[1] for n in nodes]])
        eng.snapshot_commit()
        eng.config_allocatable({"Least": 0, "Most": 1}[case["mode"]], case["weights"])
        eng.pods_upload(1)
        eng.eval(E.PLUGIN_ALLOCATABLE)
        got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)[0, :len(nodes)]
        assert list(got) == case["expected"], case["name"]
    for bad in g["invalid_args"]:
        with pytest.raises(E.B200SError):
            eng.config_allocatable(0, bad["weights"])


def test_allocatable_edge_feasibility(eng, engine_mod, oracle):
    E = engine_mod
    N, P = 700, 6
    nodes = synth.gen_nodes(7, N)
    cols = setup_alloc(eng, E, nodes, 1)
    feas_b = np.zeros((P, N), dtype=bool)
    feas_b[1, :] = True                 # all feasible
    feas_b[2, 5] = True                 # one feasible node -> range 0 -> score 0
    feas_b[3, [10, 11]] = True          # two nodes
    feas_b[4, ::7] = True
    feas_b[5, N - 1] = feas_b[5, 0] = True
    # pod 0: nothing feasible
    feas = E.pack_bits(feas_b, eng.Npad)
    eng.pods_upload(P, feasible=feas)
    eng.eval(E.PLUGIN_ALLOCATABLE)
    got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)
    want = oracle.alloc_batch(cols, W_DEFAULT, 1, P, feas, pitch=eng.Npad)
    assert np.array_equal(got, want)
    assert not got[0].any() and not got[2].any()


def test_allocatable_sparse_and_empty_feasible_sets_far_from_the_ends(eng, engine_mod, oracle):
    """Behind a chain of filters a pod's feasible set can be empty or a handful of nodes in the MIDDLE of the sorted
    order: the bounded scan from the two ends (512 entries each) gives up and the warp takes the exact min / max over
    the feasibility words instead -- same NormalizeScore."""
    E = engine_mod
    N, P = 9000, 12
    nodes = synth.gen_nodes(17, N)
    rng = np.random.default_rng(17)
    for mode in (0, 1):
        cols = setup_alloc(eng, E, nodes, mode)
        raw_order = np.argsort(nodes["alloc_cpu_milli"].astype(np.float64) * (1 << 20) + nodes["alloc_mem_bytes"], kind="stable")
        feas_b = np.zeros((P, N), dtype=bool)
        feas_b[1, raw_order[N // 2]] = True                      # one node in the middle
        feas_b[2, raw_order[N // 2 - 40:N // 2 + 40:7]] = True   # a few around the middle
        feas_b[3, raw_order[600]] = feas_b[3, raw_order[N - 700]] = True   # just beyond both scan windows
        feas_b[4, raw_order[3]] = feas_b[4, raw_order[N // 3]] = True      # low end found by the scan, high end not
        feas_b[5, raw_order[N // 3]] = feas_b[5, raw_order[N - 2]] = True  # the other way round
        feas_b[6] = rng.random(N) < 0.002
        feas_b[7] = True
        feas_b[8, raw_order[511]] = feas_b[8, raw_order[N - 512]] = True   # last entries of the scan windows
        feas_b[9, raw_order[512]] = feas_b[9, raw_order[N - 513]] = True   # first entries beyond them
        feas_b[10] = rng.random(N) < 0.5
        # pods 0 and 11: nothing feasible
        feas = E.pack_bits(feas_b, eng.Npad)
        eng.pods_upload(P, feasible=feas)
        for dt in (E.OUT_I64, E.OUT_U8):
            eng.eval(E.PLUGIN_ALLOCATABLE, dt)
            got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE, dt).astype(np.int64)
            want = oracle.alloc_batch(cols, W_DEFAULT, mode, P, feas, pitch=eng.Npad)
            assert np.array_equal(got, want), (mode, dt, np.argwhere(got != want)[:5])
        assert not got[0].any() and not got[11].any()


def test_allocatable_generic_int64_path(eng, engine_mod, oracle):
    """Huge weights / wrapping ranges take the exact generic path (Go wraps, allocatable.go:126,163)."""
    E = engine_mod
    N, P = 300, 4
    rng = np.random.default_rng(5)
    cpu = rng.integers(1, 1 << 40, N).astype(np.int64)
    mem = rng.integers(1, 1 << 62, N).astype(np.int64)
    w = [(1 << 40) + 12345, 3]
    eng.snapshot_begin(N)
    eng.snapshot_allocatable([cpu, mem])
    eng.snapshot_commit()
    for mode in (0, 1):
        eng.config_allocatable(mode, w)
        feas = synth.gen_feasible_words(11, P, N, eng.Npad)
        eng.pods_upload(P, feasible=feas)
        eng.eval(E.PLUGIN_ALLOCATABLE)
        got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)
        want = oracle.alloc_batch([cpu, mem], w, mode, P, feas, pitch=eng.Npad)
        assert np.array_equal(got, want)


def test_allocatable_three_resources_and_reconfig(eng, engine_mod, oracle):
    E = engine_mod
    N, P = 1000, 16
    nodes = synth.gen_nodes(3, N)
    w = [1 << 20, 1, 7]
    cols = setup_alloc(eng, E, nodes, 0, w, extra_cols=[nodes["alloc_ephemeral_bytes"]])
    eng.pods_upload(P)
    eng.eval(E.PLUGIN_ALLOCATABLE)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_ALLOCATABLE), oracle.alloc_batch(cols, w, 0, P, pitch=eng.Npad))
    eng.config_allocatable(1, [5, 1, 1])  # args change without a new snapshot
    eng.eval(E.PLUGIN_ALLOCATABLE)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_ALLOCATABLE),
                          oracle.alloc_batch(cols, [5, 1, 1], 1, P, pitch=eng.Npad))
    with pytest.raises(E.B200SError):  # weights must match the snapshot's resource columns
        eng.config_allocatable(1, [1, 1])
        eng.eval(E.PLUGIN_ALLOCATABLE)


def test_empty_batch_and_state_errors(eng, engine_mod):
    E = engine_mod
    with pytest.raises(E.B200SError):
        eng.pods_upload(1)  # no snapshot
    nodes = synth.gen_nodes(1, 10)
    setup_alloc(eng, E, nodes, 0)
    with pytest.raises(E.B200SError):
        eng.eval(E.PLUGIN_ALLOCATABLE)  # no pods
    eng.pods_upload(0)
    eng.eval(E.PLUGIN_ALLOCATABLE)
    assert eng.fetch_scores(E.PLUGIN_ALLOCATABLE).shape == (0, eng.Npad)
    with pytest.raises(E.B200SError):
        eng.eval(E.PLUGIN_TLP)  # no TLP columns in this snapshot


def trimaran_snapshot(eng, nodes, tri):
    eng.snapshot_begin(nodes["N"])
    eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
    eng.snapshot_lvrb(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                      nodes["alloc_mem_bytes"], tri["lvrb_flags"])
    eng.snapshot_commit()


@pytest.mark.parametrize("P,N,target", [(128, 1000, 40), (33, 2049, 40), (64, 777, 70), (5, 129, 100), (5, 129, 1)])
def test_tlp_matches_oracle(eng, engine_mod, oracle, P, N, target):
    E = engine_mod
    seed = synth.BASE_SEED + 3
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    nodes["cap_cpu_milli"][::97] = 0  # cap == 0 branch (targetloadpacking.go:170)
    trimaran_snapshot(eng, nodes, tri)
    eng.config_tlp(target)
    eng.pods_upload(P, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"])
    eng.eval(E.PLUGIN_TLP)
    got = eng.fetch_scores(E.PLUGIN_TLP)
    want = oracle.tlp_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"],
                            pods["tlp_pod_cpu_milli"], target, pitch=eng.Npad)
    assert np.array_equal(got, want)
    eng.eval(E.PLUGIN_TLP, E.OUT_U8)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_TLP, E.OUT_U8).astype(np.int64), want)


@pytest.mark.parametrize("margin,sens", [(1.0, 1.0), (1.0, 2.0), (2.0, 0.0), (1.0, -1.0), (-1.0, 1.0), (0.5, 2.0)])
def test_lvrb_matches_oracle(eng, engine_mod, oracle, margin, sens):
    E = engine_mod
    P, N = 96, 1537
    seed = synth.BASE_SEED + 3
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    nodes["alloc_cpu_milli"][::101] = 0  # capacity <= 0 (analysis.go:35)
    trimaran_snapshot(eng, nodes, tri)
    eng.config_lvrb(margin, sens)
    eng.pods_upload(P, lvrb_req_cpu_milli=pods["req_cpu_milli"], lvrb_req_mem_bytes=pods["req_mem_bytes"])
    eng.eval(E.PLUGIN_LVRB)
    got = eng.fetch_scores(E.PLUGIN_LVRB)
    want = oracle.lvrb_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"],
                             nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"], tri["lvrb_flags"],
                             pods["req_cpu_milli"], pods["req_mem_bytes"], margin, sens, pitch=eng.Npad)
    assert np.array_equal(got, want)


def test_lvrb_general_pow_tolerance(eng, engine_mod, oracle):
    """Non-special sensitivities go through pow(): equal, or off by one only at a rounding boundary
    (SURVEY §8c rule ii) — and reported, never silently accepted."""
    E = engine_mod
    P, N = 64, 2000
    seed = synth.BASE_SEED + 3
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    trimaran_snapshot(eng, nodes, tri)
    eng.config_lvrb(1.0, 3.0)
    eng.pods_upload(P, lvrb_req_cpu_milli=pods["req_cpu_milli"], lvrb_req_mem_bytes=pods["req_mem_bytes"])
    eng.eval(E.PLUGIN_LVRB)
    got = eng.fetch_scores(E.PLUGIN_LVRB)
    want = oracle.lvrb_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"],
                             nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"], tri["lvrb_flags"],
                             pods["req_cpu_milli"], pods["req_mem_bytes"], 1.0, 3.0, pitch=eng.Npad)
    diff = np.abs(got - want)
    assert diff.max() <= 1
    print(f"general pow: {int((diff != 0).sum())} of {diff.size} scores differ by 1 (boundary cases)")
    assert (diff != 0).mean() < 1e-6


def test_trimaran_golden_through_cuda(eng, engine_mod):
    import json
    import os

    from conftest import GOLDEN

    E = engine_mod
    for case in json.load(open(os.path.join(GOLDEN, "tlp.json")))["cases"]:
        eng.snapshot_begin(1)
        eng.snapshot_tlp([case["util"]], [case["cap_milli"]], [case["missing_milli"]], [case["flags"]])
        eng.snapshot_commit()
        eng.config_tlp(case["target"])
        eng.pods_upload(1, tlp_pod_cpu_milli=[case["pod_cpu_milli"]])
        eng.eval(E.PLUGIN_TLP)
        assert eng.fetch_scores(E.PLUGIN_TLP)[0, 0] == case["expected"], case["name"]
    g = json.load(open(os.path.join(GOLDEN, "lvrb.json")))["score"]
    for case in g["cases"]:
        eng.snapshot_begin(1)
        eng.snapshot_lvrb([case["cpu_avg"]], [case["cpu_std"]], [case["mem_avg"]], [case["mem_std"]],
                          [g["alloc_cpu_milli"]], [g["alloc_mem_bytes"]], [case["flags"]])
        eng.snapshot_commit()
        eng.config_lvrb(g["margin"], g["sensitivity"])
        eng.pods_upload(1, lvrb_req_cpu_milli=[case["req_cpu_milli"]], lvrb_req_mem_bytes=[case["req_mem_bytes"]])
        eng.eval(E.PLUGIN_LVRB)
        assert eng.fetch_scores(E.PLUGIN_LVRB)[0, 0] == case["expected"], case["name"]


def test_full_size_properties_config2(eng, engine_mod, oracle):
    """BASELINE config 2 at full size (10k pods x 50k nodes, Most + NormalizeScore): the oracle
    checks a sample of rows bit-exactly, size-independent properties cover the whole matrix."""
    E = engine_mod
    import torch

    P, N = 10_000, 50_000
    seed = synth.BASE_SEED + 2
    nodes = synth.gen_nodes(seed, N)
    cols = setup_alloc(eng, E, nodes, 1)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad)
    eng.pods_upload(P, feasible=feas)
    eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
    got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE, E.OUT_U8)
    rows = np.r_[0:8, P // 2:P // 2 + 8, P - 8:P]
    want = oracle.alloc_batch(cols, W_DEFAULT, 1, len(rows), feas[rows], pitch=eng.Npad)
    assert np.array_equal(got[rows].astype(np.int64), want)
    # properties over the full matrix: range, infeasible -> 0, every row with >= 2 distinct
    # feasible raw values reaches both 0 and 100, monotone in raw score
    assert got.max() <= 100
    fb = E.unpack_bits(feas[:64], N)
    g64 = got[:64, :N]
    assert not g64[~fb].any()
    raw = np.array([oracle.alloc_score([cols[0][n], cols[1][n]], W_DEFAULT, 1) for n in range(0, N, 97)])
    sub = g64[:, ::97]
    order = np.argsort(raw, kind="stable")
    for p in range(64):
        f = fb[p, ::97][order]
        s = sub[p][order][f]
        assert np.all(np.diff(s.astype(np.int64)) >= 0)
    assert (got[:, :N].max(axis=1) == 100).all()
    # i64 layout == u8 layout
    eng.eval(E.PLUGIN_ALLOCATABLE, E.OUT_I64)
    t = torch.empty(0)  # noqa: F841 (torch is only the device plumbing here)
    got64 = eng.fetch_scores(E.PLUGIN_ALLOCATABLE, E.OUT_I64)
    assert np.array_equal(got64[rows], want)
    # checksum of checksums: both transports agree on the whole matrix
    assert int(got64.sum()) == int(got.astype(np.int64).sum())


# ---------------------------------------------------------------- NetworkOverhead
def netoh_setup(eng, E, net, N, node_offset=0, n_global=None):
    eng.snapshot_begin(N, node_offset=node_offset, n_nodes_global=n_global or (node_offset + N))
    eng.snapshot_network_overhead(net["region_all"][node_offset:node_offset + N],
                                  net["zone_all"][node_offset:node_offset + N], net["zone_cost"], net["region_cost"])
    eng.snapshot_commit()


@pytest.mark.parametrize("P,N,masked", [(64, 1000, False), (97, 3001, True), (8, 64, True)])
def test_network_overhead_matches_oracle(eng, engine_mod, oracle, P, N, masked):
    E = engine_mod
    seed = synth.BASE_SEED + 5
    net = synth.gen_netoh(seed, N, P)
    netoh_setup(eng, E, net, N)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad) if masked else None
    eng.pods_upload(P, feasible=feas, netoh=net)
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
    got = eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD)
    gf = eng.fetch_feasible(E.PLUGIN_NETWORK_OVERHEAD)
    gr = eng.fetch_reasons(E.PLUGIN_NETWORK_OVERHEAD)
    ws, wf, wr = oracle.netoh_batch(net["zone_cost"], net["region_cost"], net["region_all"], net["zone_all"],
                                    net["score_equally"], net["dep_offset"], net["deps"], feas, pitch=eng.Npad)
    assert np.array_equal(gf, wf)
    assert np.array_equal(gr, wr)
    assert np.array_equal(got, ws)
    assert (gr == 7).any() and (got == 100).any()  # the fixture exercises rejects and best nodes
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD, E.OUT_U8)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD, E.OUT_U8).astype(np.int64), ws)


@pytest.mark.parametrize("regions,zones_per_region", [(2, 3), (20, 16)])
def test_network_overhead_small_and_large_pair_dictionaries(eng, engine_mod, oracle, regions, zones_per_region):
    """The pair tables of a pod tile live in shared memory when the (region, zone) dictionary is small (4 nodes per
    thread, packed stores); 320 zones exceed it and keep the global-table kernel.  Ragged N, hosted nodes, masks."""
    E = engine_mod
    P, N = 70, 5000 + 37
    net = synth.gen_netoh(synth.BASE_SEED + 11, N, P, n_regions=regions, zones_per_region=zones_per_region)
    netoh_setup(eng, E, net, N)
    feas = synth.gen_feasible_words(synth.BASE_SEED + 11, P, N, eng.Npad)
    eng.pods_upload(P, feasible=feas, netoh=net)
    ws, wf, wr = oracle.netoh_batch(net["zone_cost"], net["region_cost"], net["region_all"], net["zone_all"],
                                    net["score_equally"], net["dep_offset"], net["deps"], feas, pitch=eng.Npad)
    for dt in (E.OUT_I64, E.OUT_U8):
        eng.eval(E.PLUGIN_NETWORK_OVERHEAD, dt)
        assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NETWORK_OVERHEAD), wf)
        assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NETWORK_OVERHEAD), wr)
        assert np.array_equal(eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD, dt).astype(np.int64), ws)


def test_network_overhead_generic_normalize(eng, engine_mod, oracle):
    """Huge costs: the float64 NormalizeScore formula verbatim (networkoverhead.go:406-410)."""
    E = engine_mod
    P, N = 16, 500
    net = synth.gen_netoh(99, N, P)
    zc = net["zone_cost"]
    zc[zc != -(2**63)] *= 10**9 + 7
    rc = net["region_cost"]
    rc[rc != -(2**63)] *= 3 * 10**15 + 1
    net["deps"]["max_network_cost"] = 2**62
    netoh_setup(eng, E, net, N)
    eng.pods_upload(P, netoh=net)
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
    ws, wf, _ = oracle.netoh_batch(net["zone_cost"], net["region_cost"], net["region_all"], net["zone_all"],
                                   net["score_equally"], net["dep_offset"], net["deps"], None, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NETWORK_OVERHEAD), wf)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD), ws)


def test_score_batch_chunked_pipeline(eng, engine_mod, oracle):
    """b200s_score_batch splits a large batch of a score-only plugin into pod chunks (D2H of chunk i overlaps H2D of
    chunk i+1): the host matrix must equal the unchunked upload + eval + fetch, and the engine-resident matrices are
    invalidated afterwards."""
    E = engine_mod
    P, N, seed = 3000, 20_000, synth.BASE_SEED + 9
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
    tri = synth.gen_trimaran(seed, nodes)
    eng.snapshot_begin(N)
    eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
    eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
    eng.snapshot_commit()
    eng.config_allocatable(1, W_DEFAULT)
    eng.config_tlp(40)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad)
    cols = dict(feasible=feas, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"])
    for plugin, dtype, npdt in ((E.PLUGIN_ALLOCATABLE, E.OUT_I64, np.int64), (E.PLUGIN_TLP, E.OUT_I64, np.int64),
                                (E.PLUGIN_ALLOCATABLE, E.OUT_U8, np.uint8)):
        eng.pods_upload(P, **cols)
        eng.eval(plugin, dtype)
        want = eng.fetch_scores(plugin, dtype)
        batch, keep = eng.make_batch(P, **cols)
        got = np.full((P, eng.Npad), 77, dtype=npdt)
        eng.score_batch(plugin, batch, dtype, got)
        assert got.nbytes >= (96 << 20) or dtype == E.OUT_U8  # the int64 cases take the chunked path
        assert np.array_equal(got, want)
        if got.nbytes >= (96 << 20):
            with pytest.raises(E.B200SError):  # only the last chunk is resident: fetching would be wrong, so it fails
                eng.fetch_scores(plugin, dtype)
    # rows spot-checked against the oracle as well
    rows = [0, 1499, 2999]
    want_rows = oracle.tlp_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"],
                                 pods["tlp_pod_cpu_milli"][rows], 40, pitch=eng.Npad)
    got = np.zeros((P, eng.Npad), dtype=np.int64)
    batch, keep = eng.make_batch(P, **cols)
    eng.score_batch(E.PLUGIN_TLP, batch, E.OUT_I64, got)
    assert np.array_equal(got[rows], want_rows)

"""The reference's own table-driven plugin tests, replayed through the C++ host mirror of the
plugin interface (scheduler-plugins_b200/host) -> C-ABI -> CUDA.  Same plugin names, methods,
argument meaning, status codes and messages as the Go plugins; expectations come from
tests/golden/*.json (transcribed / extracted from the reference's *_test.go files)."""
import json
import os

import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H(built):
    from scheduler_plugins_b200 import _b200host

    return _b200host


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def make_node(H, name, alloc=None, capacity=None, labels=None):
    n = H.Node()
    n.name = name
    n.allocatable = H.resource_list({k: str(v) for k, v in (alloc or {}).items()})
    n.capacity = H.resource_list({k: str(v) for k, v in (capacity or alloc or {}).items()})
    n.labels = labels or {}
    return n


def make_pod(H, spec, name="p", labels=None, node_name=""):
    from test_host_cpu import mkpod

    p = mkpod(H, spec)
    p.name = name
    p.labels = labels or {}
    p.node_name = node_name
    return p


def handle_with(H, nodes):
    h = H.Handle()
    h.node_infos = [H.NodeInfo(n) for n in nodes]
    return h


# ------------------------------------------------------------------ NodeResourcesAllocatable
@pytest.mark.parametrize("case", load("allocatable.json")["cases"], ids=lambda c: c["name"][:60])
def test_node_resources_allocatable(H, case):
    """allocatable_test.go:239-312: Score on every node, then NormalizeScore on the list."""
    nodes = [make_node(H, f"machine{i + 1}", {"cpu": f"{c}m", "memory": str(m)}) for i, (c, m) in enumerate(case["nodes"])]
    fh = handle_with(H, nodes)
    args = H.NodeResourcesAllocatableArgs()
    args.mode = case["mode"]
    args.resources = [H.ResourceSpec("cpu", case["weights"][0]), H.ResourceSpec("memory", case["weights"][1])]
    p = H.Allocatable.new(args, fh)
    assert p.name() == "NodeResourcesAllocatable"
    state, pod = H.CycleState(), make_pod(H, {})
    assert p.pre_score(state, pod, fh.node_infos).is_success()  # upstream hands PreScore the feasible nodes
    got = []
    for ni in fh.node_infos:
        score, status = p.score(state, pod, ni)
        assert status.is_success()
        got.append(H.NodeScore(ni.node.name, score))
    status, got = p.normalize_score(state, pod, got)
    assert status.is_success()
    assert [g.score for g in got] == case["expected"]


def test_node_resources_allocatable_invalid_args(H):
    fh = handle_with(H, [make_node(H, "machine", {"cpu": "4000m", "memory": "10000"})])
    for res, want in (([("memory", -1), ("cpu", 1)],
                       "resources[0].weight: Invalid value: -1: resource weight of memory should be a positive value, got :-1"),
                      ([("memory", 1), ("cpu", 0)],
                       "resources[1].weight: Invalid value: 0: resource weight of cpu should be a positive value, got :0")):
        args = H.NodeResourcesAllocatableArgs()
        args.resources = [H.ResourceSpec(n, w) for n, w in res]
        with pytest.raises(ValueError) as ei:
            H.Allocatable.new(args, fh)
        assert want in str(ei.value)  # allocatable_test.go:222-237


def test_node_resources_allocatable_defaults_and_nil_node(H):
    nodes = [make_node(H, "a", {"cpu": "4", "memory": "10Gi"}), make_node(H, "b", {"cpu": "8", "memory": "10Gi"})]
    fh = handle_with(H, nodes)
    p = H.Allocatable.new(None, fh)  # nil args: Least, {cpu: 1<<20, memory: 1}
    state, pod = H.CycleState(), make_pod(H, {})
    assert [p.score(state, pod, ni)[0] for ni in fh.node_infos] == [100, 0]
    score, status = p.score(state, pod, H.NodeInfo())
    assert status.code == H.Code.Error and status.message == "node not found"  # resource_allocation.go:53-56


# ------------------------------------------------------------------ Trimaran
def watcher(H, node_metrics):
    wm = H.WatcherMetrics()
    wm.has_map = node_metrics is not None
    if node_metrics:
        d = {}
        for node, ms in node_metrics.items():
            nm = H.NodeMetrics()
            nm.metrics = [H.Metric(*m) for m in ms]
            d[node] = nm
        wm.node_metrics = d
    return wm


def test_target_load_packing_scoring(H):
    """targetloadpacking_test.go:118-281 (node capacity cpu 1000m / memory 1Gi; target 40; args without
    DefaultRequests -> requestsMilliCores 0)."""
    cases = [
        ("new node", {}, {"node-1": [("CPU", "Latest", 0.0)]}, 40),
        ("hot node", {}, {"node-1": [("CPU", "Latest", 50.0)]}, 33),
        ("excess utilization returns min score",
         {"containers": [{"requests": {"cpu": "1000m"}, "limits": {"cpu": "1000m"}}], "overhead": {"cpu": "0"}},
         {"node-1": [("CPU", "Latest", 30.0)]}, 0),
        ("404 resp from watcher", {}, None, 0),
    ]
    for name, podspec, metrics, want in cases:
        fh = handle_with(H, [make_node(H, "node-1", {"cpu": "1000m", "memory": "1Gi"})])
        fh.metrics = watcher(H, metrics)
        args = H.TargetLoadPackingArgs()
        args.default_requests_cpu_milli = 0
        p = H.TargetLoadPacking.new(args, fh)
        score, status = p.score(H.CycleState(), make_pod(H, podspec), fh.node_infos[0])
        assert status.is_success() and score == want, name


def test_target_load_packing_missing_utilisation_and_last_metric(H):
    """:131-140 the LAST matching CPU metric wins; :151-167 recently bound pods add predicted CPU."""
    fh = handle_with(H, [make_node(H, "node-1", {"cpu": "1000m"})])
    wm = watcher(H, {"node-1": [("CPU", "AVG", 90.0), ("CPU", "Latest", 10.0), ("CPU", "STD", 5.0)]})
    wm.window_end = 1000
    fh.metrics = wm
    bound = make_pod(H, {"containers": [{"requests": {"cpu": "100m"}, "limits": {"cpu": "100m"}}]}, name="bound")
    old = make_pod(H, {"containers": [{"limits": {"cpu": "500m"}, "requests": {}}]}, name="old")
    fh.scheduled_pods_cache = {"node-1": [H.ScheduledPodInfo(990, bound), H.ScheduledPodInfo(900, old)]}
    p = H.TargetLoadPacking.new(H.TargetLoadPackingArgs(), fh)
    pod = make_pod(H, {"containers": [{"requests": {"cpu": "100m"}, "limits": {}}]})  # 100m * 1.5 = 150
    score, _ = p.score(H.CycleState(), pod, fh.node_infos[0])
    predicted = 100 * (100 + 150 + 100) / 1000  # util 10% of 1000m + pod + bound pod (inside 60 s of window end)
    import math

    assert score == math.floor((100 - 40) * predicted / 40 + 40 + 0.5) == 93  # math.Round(92.5) = 93 (half away from zero)


def test_load_variation_risk_balancing_score(H):
    """loadvariationriskbalancing_test.go:136-384."""
    g = load("lvrb.json")["score"]
    specs = {
        "new node": ({}, [("CPU", "AVG", 50.0)]),
        "hot node": ({}, [("CPU", "AVG", 100.0)]),
        "average and stDev metrics": ({"containers": [{"requests": {"cpu": "200m", "memory": str(256 << 20)}}], "overhead": {"cpu": "0"}},
                                      [("CPU", "AVG", 30.0), ("CPU", "STD", 16.0)]),
        "CPU and Memory metrics": ({"containers": [{"requests": {"cpu": "100m", "memory": str(512 << 20)}}], "overhead": {"cpu": "0"}},
                                   [("CPU", "AVG", 40.0), ("CPU", "STD", 16.0), ("Memory", "AVG", 50.0), ("Memory", "STD", 10.0)]),
        "pick worst case: CPU or Memory": ({"containers": [{"requests": {"cpu": "100m", "memory": str(512 << 20)}}], "overhead": {"cpu": "0"}},
                                           [("CPU", "AVG", 80.0), ("CPU", "STD", 20.0), ("Memory", "AVG", 25.0), ("Memory", "STD", 15.0)]),
        "404 resp from watcher": ({}, None),
    }
    for case in g["cases"]:
        podspec, ms = specs[case["name"]]
        fh = handle_with(H, [make_node(H, "node-1", {"cpu": "1000m", "memory": "1Gi"})])
        fh.metrics = watcher(H, None if ms is None else {"node-1": ms})
        p = H.LoadVariationRiskBalancing.new(H.LoadVariationRiskBalancingArgs(), fh)
        score, status = p.score(H.CycleState(), make_pod(H, podspec), fh.node_infos[0])
        assert status.is_success() and score == case["expected"], case["name"]


# ------------------------------------------------------------------ NodeResourceTopologyMatch
def nrt_handle(H, suite_nodes):
    from oracle import flatten as F  # only for quantity arithmetic of the fixture (sum of zone Available)

    nodes = []
    fh = H.Handle()
    nrts = {}
    for n in suite_nodes:
        alloc = {}
        for z in n["zones"]:
            for r, q in z["resources"].items():
                alloc[r] = alloc.get(r, 0) + F.milli(q["available"])
        alloc = {r: f"{v}m" for r, v in alloc.items()}
        alloc.update(n.get("node_extra", {}))
        nodes.append(make_node(H, n["name"], alloc))
        t = H.NodeResourceTopology()
        t.name = n["name"]
        t.topology_policies = n["policies"]
        t.attributes = n.get("attributes", {})
        zs = []
        for z in n["zones"]:
            zz = H.Zone()
            zz.name, zz.type = z["name"], z.get("type", "Node")
            zz.resources = {r: H.ZoneResource(H.parse_quantity(q["capacity"]), H.parse_quantity(q["available"]))
                            for r, q in z["resources"].items()}
            zz.costs = z.get("costs", {})
            zs.append(zz)
        t.zones = zs
        nrts[n["name"]] = t
    fh.node_infos = [H.NodeInfo(n) for n in nodes]
    fh.nrts = nrts
    return fh


def nrt_filter_cases():
    out = []
    for s in load("nrt_filter.json")["suites"]:
        for c in s["cases"]:
            out.append(pytest.param(s, c, id=f"{s['suite'][-14:]}::{c['name'][:60]}"))
    return out


@pytest.mark.parametrize("suite,case", nrt_filter_cases())
def test_topology_match_filter(H, suite, case):
    """filter_test.go: tm.Filter(ctx, cycleState, pod, nodeInfo) vs wantStatus (code + message prefix)."""
    fh = nrt_handle(H, suite["nodes"])
    tm = H.TopologyMatch.new(H.NodeResourceTopologyMatchArgs(), fh)
    got = tm.filter(H.CycleState(), make_pod(H, case["pod"]), fh.node_infos[case["node"]])
    if case["want"] is None:
        assert got.is_success(), got.message
    else:
        assert got.code == H.Code.Unschedulable and got.message.startswith(case["want"]["message"])  # quasiEqualStatus


def test_topology_match_scores(H):
    g = load("nrt_score.json")
    s0 = g["suites"][0]
    fixture = [dict(n, policies=[s0["policy_override"]]) for n in s0["nodes"]]
    for case in s0["cases"]:
        fh = nrt_handle(H, fixture)
        args = H.NodeResourceTopologyMatchArgs()
        args.scoring_strategy = case["strategy"]
        tm = H.TopologyMatch.new(args, fh)
        state, pod = H.CycleState(), make_pod(H, case["pod"])
        got = {ni.node.name: tm.score(state, pod, ni)[0] for ni in fh.node_infos}
        (wn, ws), = case["want_max"].items()
        assert got[wn] == ws and max(got.values()) == ws, case["name"]
    s1 = g["suites"][1]
    for case in s1["cases"]:
        fx = s1["fixtures"][case["fixture"]]
        if case["policy_override"]:
            fx = [dict(n, policies=[case["policy_override"]]) for n in fx]
        fh = nrt_handle(H, fx)
        args = H.NodeResourceTopologyMatchArgs()
        args.scoring_strategy = "LeastNUMANodes"
        tm = H.TopologyMatch.new(args, fh)
        state, pod = H.CycleState(), make_pod(H, case["pod"])
        got = {ni.node.name: tm.score(state, pod, ni)[0] for ni in fh.node_infos}
        assert got == case["want"], case["name"]
    with pytest.raises(ValueError):
        bad = H.NodeResourceTopologyMatchArgs()
        bad.scoring_strategy = "Nope"
        H.TopologyMatch.new(bad, nrt_handle(H, fixture))


def test_topology_match_outside_the_dense_encoding(H):
    """Reason code 9 has somewhere to go: nodes whose zone list is not in NUMA-id order (or is sparse) are answered by
    the host's scalar path (host/nrt_scalar.cpp) -- Filter statuses and Least/Most/Balanced scores equal those of the
    same node listed in id order, which the engine evaluates itself."""
    g = load("nrt_filter.json")
    for suite in g["suites"]:
        straight = nrt_handle(H, suite["nodes"])
        shuffled = nrt_handle(H, [dict(n, zones=list(reversed(n["zones"]))) for n in suite["nodes"]])
        for strategy in ("LeastAllocated", "MostAllocated", "BalancedAllocation"):
            args = H.NodeResourceTopologyMatchArgs()
            args.scoring_strategy = strategy
            a, b = H.TopologyMatch.new(args, straight), H.TopologyMatch.new(args, shuffled)
            for case in suite["cases"][:12]:
                pod = make_pod(H, case["pod"])
                sa, sb = H.CycleState(), H.CycleState()
                for i in range(len(suite["nodes"])):
                    fa, fb = a.filter(sa, pod, straight.node_infos[i]), b.filter(sb, pod, shuffled.node_infos[i])
                    assert (fa.code, fa.message) == (fb.code, fb.message), (suite["suite"], case["name"], i)
                    if fa.is_success():
                        assert a.score(sa, pod, straight.node_infos[i])[0] == b.score(sb, pod, shuffled.node_infos[i])[0], \
                            (suite["suite"], case["name"], i, strategy)


def test_topology_match_stale_and_missing_nrt(H):
    """filter.go:194-200 / score.go:79-86."""
    g = load("nrt_score.json")["suites"][0]
    fixture = [dict(n, policies=["SingleNUMANodePodLevel"]) for n in g["nodes"]]
    fh = nrt_handle(H, fixture)
    fh.nrt_not_fresh = {"Node1": True}
    nrts = dict(fh.nrts)
    del nrts["Node2"]
    fh.nrts = nrts
    tm = H.TopologyMatch.new(H.NodeResourceTopologyMatchArgs(), fh)
    pod = make_pod(H, g["cases"][0]["pod"])
    state = H.CycleState()
    st = tm.filter(state, pod, fh.node_infos[0])
    assert st.code == H.Code.Unschedulable and st.message == "invalid node topology data"
    assert tm.filter(state, pod, fh.node_infos[1]).is_success()  # no NRT object: pass
    assert tm.score(state, pod, fh.node_infos[1])[0] == 0         # ... but score 0
    assert tm.filter(state, pod, H.NodeInfo()).message == "node not found"


# ------------------------------------------------------------------ NetworkOverhead
AG, SEL = "appgroup.diktyo.x-k8s.io", "appgroup.diktyo.x-k8s.io.workload"


def netoh_fixture(H, placed):
    """GetAppGroupCRBasic / GetNetworkTopologyCRBasic / nodes n-1..n-8 (networkoverhead_test.go:188-347, :579-596)."""
    nodes = []
    for i, (region, zone) in enumerate([("us-west-1", "Z1"), ("us-west-1", "Z1"), ("us-west-1", "Z2"), ("us-west-1", "Z2"),
                                        ("us-east-1", "Z3"), ("us-east-1", "Z3"), ("us-east-1", "Z4"), ("us-east-1", "Z4")]):
        nodes.append(make_node(H, f"n-{i + 1}", {"cpu": "8000m", "memory": "16Gi"},
                               labels={"topology.kubernetes.io/region": region, "topology.kubernetes.io/zone": zone}))
    fh = handle_with(H, nodes)
    fh.app_groups = {"basic": H.AppGroup("basic", [
        H.AppGroupWorkload("p1", [H.DependencyInfo("p2")]), H.AppGroupWorkload("p2", [H.DependencyInfo("p3")]),
        H.AppGroupWorkload("p3", [])])}
    fh.network_topologies = {"nt-test": H.NetworkTopology("nt-test", [H.WeightInfo("UserDefined", [
        H.TopologyInfo("topology.kubernetes.io/region", [
            H.OriginInfo("us-west-1", [H.CostInfo("us-east-1", 20)]), H.OriginInfo("us-east-1", [H.CostInfo("us-west-1", 20)])]),
        H.TopologyInfo("topology.kubernetes.io/zone", [
            H.OriginInfo("Z1", [H.CostInfo("Z2", 5)]), H.OriginInfo("Z2", [H.CostInfo("Z1", 5)]),
            H.OriginInfo("Z3", [H.CostInfo("Z4", 10)]), H.OriginInfo("Z4", [H.CostInfo("Z3", 10)])])])])}
    fh.pods = [make_pod(H, {"containers": [{}]}, name=f"{sel}-deployment", labels={AG: "basic", SEL: sel}, node_name=host)
               for sel, host in placed]
    args = H.NetworkOverheadArgs()
    args.network_topology_name = "nt-test"
    return fh, H.NetworkOverhead.new(args, fh)


def test_network_overhead_score(H):
    """TestNetworkOverheadScore (networkoverhead_test.go:572-818): raw Score per node, then NormalizeScore."""
    for case, sel in zip(load("network_overhead.json")["score_cases"], ("p1", "p2", "p3")):
        fh, no = netoh_fixture(H, [("p1", "n-2"), ("p2", "n-5"), ("p3", "n-1")])
        pod = make_pod(H, {"containers": [{}]}, name=f"{sel}-deployment", labels={AG: "basic", SEL: sel})
        state = H.CycleState()
        assert no.pre_filter(state, pod, fh.node_infos).is_success()
        scores = []
        for ni in fh.node_infos:
            s, st = no.score(state, pod, ni)
            assert st.is_success()
            scores.append(H.NodeScore(ni.node.name, s))
        assert [x.score for x in scores] == case["raw"], case["name"]           # wantedScoresBefore
        st, scores = no.normalize_score(state, pod, scores)
        assert st.is_success() and [x.score for x in scores] == case["normalized"], case["name"]  # wantedScoresAfter


def test_network_overhead_filter(H):
    """TestNetworkOverheadFilter (networkoverhead_test.go:1055-1276), exact status messages."""
    pods_placed = [("p1", "n-2"), ("p2", "n-5"), ("p3", "n-8")]
    cases = [("p1", "basic", 0, "Node n-1 does not meet several network requirements from Workload dependencies: Satisfied: 0 Violated: 1"),
             ("p1", "basic", 5, None),
             ("p2", "basic", 4, "Node n-5 does not meet several network requirements from Workload dependencies: Satisfied: 0 Violated: 1"),
             ("p2", "basic", 6, None), ("p3", "basic", 0, None), ("p10", "", 0, None)]
    for sel, ag, node, want in cases:
        fh, no = netoh_fixture(H, pods_placed)
        pod = make_pod(H, {"containers": [{}]}, name=f"{sel}-deployment", labels={AG: ag, SEL: sel} if ag else {SEL: sel, AG: ""})
        state = H.CycleState()
        assert no.pre_filter(state, pod, fh.node_infos).is_success()
        got = no.filter(state, pod, fh.node_infos[node])
        if want is None:
            assert got.is_success(), (sel, node, got.message)
        else:
            assert got.code == H.Code.Unschedulable and got.message == want
    # Filter without PreFilter: error status (networkoverhead.go:336-340)
    fh, no = netoh_fixture(H, pods_placed)
    st = no.filter(H.CycleState(), make_pod(H, {}), fh.node_infos[0])
    assert st.code == H.Code.Error and "failed to read from cycleState" in st.message


# ------------------------------------------------------------------ incremental snapshot
def _scores(p, H, pod, fh, pre=True):
    state = H.CycleState()
    if pre:
        assert p.pre_score(state, pod, fh.node_infos).is_success()
    out = []
    for ni in fh.node_infos:
        s, st = p.score(state, pod, ni)
        assert st.is_success()
        out.append(s)
    return out


def test_incremental_snapshot_allocatable(H):
    """A node event touches one NodeInfo: the plugin rewrites that row (b200s_snapshot_patch_allocatable) and
    scores exactly like a plugin instance that flattens the changed cluster from scratch."""
    nodes = [make_node(H, f"machine{i}", {"cpu": f"{1000 * (i % 7 + 1)}m", "memory": str((i % 5 + 1) << 30)}) for i in range(40)]
    fh = handle_with(H, nodes)
    p = H.Allocatable.new(None, fh)
    pod = make_pod(H, {})
    before = _scores(p, H, pod, fh)
    assert p.patched_rows() == 0
    for i, cpu in ((3, "64000m"), (17, "100m"), (3, "48000m")):
        nodes[i].allocatable = H.resource_list({"cpu": cpu, "memory": str(1 << 30)})
        fh.touch_node(i)
    after = _scores(p, H, pod, fh)
    assert p.patched_rows() == 2                      # nodes 3 and 17, once each
    assert after != before
    assert after == _scores(H.Allocatable.new(None, fh), H, pod, fh)
    fh.touch()                                        # a list-level change re-flattens everything
    assert _scores(p, H, pod, fh) == after and p.patched_rows() == 2
    for i in range(20):                               # too many rows for a patch to pay off: bulk path
        fh.touch_node(i)
    assert _scores(p, H, pod, fh) == after and p.patched_rows() == 2


def test_incremental_snapshot_survives_a_failed_patch(H):
    """A patch that fails half-way leaves the engine's snapshot open; the plugin must fall back to the full upload
    (which resets it) instead of returning an Error status for every later cycle."""
    nodes = [make_node(H, f"machine{i}", {"cpu": f"{1000 * (i % 7 + 1)}m", "memory": str((i % 5 + 1) << 30)}) for i in range(40)]
    fh = handle_with(H, nodes)
    p = H.Allocatable.new(None, fh)
    pod = make_pod(H, {})
    _scores(p, H, pod, fh)
    nodes[5].allocatable = H.resource_list({"cpu": "64000m", "memory": str(1 << 30)})
    fh.touch_node(5)
    assert p.debug_leave_patch_open() == 0            # the engine now refuses b200s_snapshot_patch_begin
    after = _scores(p, H, pod, fh)                    # ... and the plugin recovers through the bulk path
    assert p.patched_rows() == 0
    assert after == _scores(H.Allocatable.new(None, fh), H, pod, fh)
    nodes[6].allocatable = H.resource_list({"cpu": "100m", "memory": str(1 << 30)})
    fh.touch_node(6)
    assert _scores(p, H, pod, fh) == _scores(H.Allocatable.new(None, fh), H, pod, fh) and p.patched_rows() == 1


def test_incremental_snapshot_trimaran_bind(H):
    """handler.go:131-167: a bind adds the pod to ScheduledPodsCache[node]; only that node's missing-utilisation
    changes, so TargetLoadPacking patches one row.  LoadVariationRiskBalancing follows node allocatable changes."""
    nodes = [make_node(H, f"node-{i}", {"cpu": "4000m", "memory": "8Gi"}) for i in range(12)]
    fh = handle_with(H, nodes)
    wm = watcher(H, {f"node-{i}": [("CPU", "AVG", 10.0 + 5 * i), ("CPU", "STD", 2.0), ("Memory", "AVG", 30.0), ("Memory", "STD", 3.0)]
                     for i in range(12)})
    wm.window_end = 1000
    fh.metrics = wm
    tlp = H.TargetLoadPacking.new(H.TargetLoadPackingArgs(), fh)
    lvrb = H.LoadVariationRiskBalancing.new(H.LoadVariationRiskBalancingArgs(), fh)
    pod = make_pod(H, {"containers": [{"requests": {"cpu": "500m", "memory": "1Gi"}, "limits": {}}]})
    t0, l0 = _scores(tlp, H, pod, fh), _scores(lvrb, H, pod, fh)
    bound = make_pod(H, {"containers": [{"requests": {"cpu": "1000m"}, "limits": {"cpu": "1000m"}}]}, name="bound")
    fh.scheduled_pods_cache = {"node-4": [H.ScheduledPodInfo(995, bound)]}
    fh.touch_node(4)
    nodes[9].allocatable = H.resource_list({"cpu": "2000m", "memory": "4Gi"})
    fh.touch_node(9)
    t1, l1 = _scores(tlp, H, pod, fh), _scores(lvrb, H, pod, fh)
    assert tlp.patched_rows() == 2 and lvrb.patched_rows() == 2
    assert t1[4] != t0[4] and [s for i, s in enumerate(t1) if i != 4] == [s for i, s in enumerate(t0) if i != 4]
    assert l1[9] != l0[9] and [s for i, s in enumerate(l1) if i != 9] == [s for i, s in enumerate(l0) if i != 9]
    assert t1 == _scores(H.TargetLoadPacking.new(H.TargetLoadPackingArgs(), fh), H, pod, fh)
    assert l1 == _scores(H.LoadVariationRiskBalancing.new(H.LoadVariationRiskBalancingArgs(), fh), H, pod, fh)


# ------------------------------------------------------------------ Trimaran Peaks / LowRiskOverCommitment
def test_peaks_score_table(H, oracle):
    """peaks_test.go:174-426: Score then NormalizeScore (a no-op in the mirror: Score is already normalised over
    the PreScore list).  With one node in the list a non-zero raw score normalises to 100 (max == min), zero to 0."""
    g = load("peaks.json")
    m = g["power_model"]["node-1"]
    args = H.PeaksArgs()
    args.node_power_model = {"node-1": H.PowerModel(m["k0"], m["k1"], m["k2"])}
    pods = {
        "Pod with Requests": {"containers": [{"requests": {"cpu": "1", "memory": "2"}}]},
        "No CPU metrics found": {"containers": [{"requests": {"cpu": "1", "memory": "2"}}]},
        "Pod with Overhead": {"overhead": {"cpu": "0"}},
        "Pod with above node resource capacity": {"containers": [{"limits": {"cpu": "2000m"}, "requests": {}}]},
        "No watcher response for node": {"containers": [{"limits": {"cpu": "2000m"}, "requests": {}}]},
        "404 resp from watcher": {},
    }
    metrics = {
        "Pod with Requests": {"node-1": [("CPU", "Latest", 0.0)]},
        "No CPU metrics found": {"node-1": [("Memory", "Latest", 0.0)]},
        "Pod with Overhead": {"node-1": [("CPU", "Latest", 0.0)]},
        "Pod with above node resource capacity": {"node-1": [("CPU", "Latest", 100.0)]},
        "No watcher response for node": {},
        "404 resp from watcher": None,
    }
    for case in g["score_cases"]:
        fh = handle_with(H, [make_node(H, "node-1", {"cpu": "1000m", "memory": "1Gi"})])
        fh.metrics = watcher(H, metrics[case["name"]])
        p = H.Peaks.new(args, fh)
        pod = make_pod(H, pods[case["name"]])
        state = H.CycleState()
        assert p.pre_score(state, pod, fh.node_infos).is_success()
        score, status = p.score(state, pod, fh.node_infos[0])
        raw = oracle.peaks_score(case["util"], case["cap_milli"], case["flags"], m["k1"], m["k2"], case["pod_cpu_milli"])
        assert status.is_success() and score == oracle.peaks_normalize([raw])[0], case["name"]
        assert score == (100 if case["name"] == "Pod with Requests" else 0)


def test_peaks_first_metric_and_normalize_over_prescore_list(H, oracle):
    """:117-126 the FIRST CPU Average|Latest metric wins (TargetLoadPacking keeps the last); NormalizeScore runs
    over the nodes upstream handed to PreScore, nodes without a power model score with {0, 0, 0}."""
    nodes = [make_node(H, f"node-{i}", {"cpu": "4000m"}) for i in range(5)]
    fh = handle_with(H, nodes)
    fh.metrics = watcher(H, {f"node-{i}": [("CPU", "AVG", 10.0 * (i + 1)), ("CPU", "Latest", 90.0)] for i in range(5)})
    args = H.PeaksArgs()
    args.node_power_model = {f"node-{i}": H.PowerModel(400.0, -90.0 - i, -0.07) for i in range(4)}  # node-4: no model
    p = H.Peaks.new(args, fh)
    pod = make_pod(H, {"containers": [{"requests": {"cpu": "500m"}}], "init": [{"requests": {"cpu": "800m"}}],
                       "overhead": {"cpu": "100m"}})  # max(500, 800) + 100 = 900m
    feasible = [fh.node_infos[i] for i in (0, 1, 3, 4)]
    state = H.CycleState()
    assert p.pre_score(state, pod, feasible).is_success()
    got = [p.score(state, pod, ni)[0] for ni in feasible]
    raw = [oracle.peaks_score(10.0 * (i + 1), 4000, 3, (-90.0 - i) if i < 4 else 0.0, -0.07 if i < 4 else 0.0, 900)
           for i in (0, 1, 3, 4)]
    assert got == list(oracle.peaks_normalize(raw)) and len(set(got)) >= 3


def test_low_risk_over_commitment(H, oracle):
    """lowriskovercommitment_test.go:140-245 'new node' (best-effort pod -> 0) and a node that already runs pods
    (GetNodeRequestsAndLimits: limits raised to requests per pod, requests capped by capacity)."""
    fh = handle_with(H, [make_node(H, "node-1", {"cpu": "1000m", "memory": "1Gi"})])
    fh.metrics = watcher(H, {"node-1": [("CPU", "AVG", 20.0)]})
    p = H.LowRiskOverCommitment.new(H.LowRiskOverCommitmentArgs(), fh)
    assert p.name() == "LowRiskOverCommitment"
    state, pod = H.CycleState(), make_pod(H, {})
    assert p.pre_score(state, pod, fh.node_infos).is_success()
    score, status = p.score(state, pod, fh.node_infos[0])
    assert status.is_success() and score == 0

    nodes = [make_node(H, f"node-{i}", {"cpu": "4000m", "memory": "8Gi"}) for i in range(3)]
    running = [make_pod(H, {"containers": [{"requests": {"cpu": "1500m", "memory": "2Gi"}, "limits": {"cpu": "1000m", "memory": "6Gi"}}]}, name="a"),
               make_pod(H, {"containers": [{"requests": {"cpu": "500m"}, "limits": {"cpu": "3000m", "memory": "1Gi"}}],
                            "init": [{"requests": {"memory": "3Gi"}, "limits": {}}]}, name="b")]
    infos = [H.NodeInfo(n) for n in nodes]
    infos[0].pods = running
    infos[1].pods = running[:1]
    fh = H.Handle()
    fh.node_infos = infos
    fh.metrics = watcher(H, {f"node-{i}": [("CPU", "AVG", 35.0), ("CPU", "STD", 6.0), ("Memory", "AVG", 50.0), ("Memory", "STD", 4.0)]
                             for i in range(2)})  # node-2 has no metrics -> 0
    args = H.LowRiskOverCommitmentArgs()
    args.risk_limit_weight_cpu = 0.3
    p = H.LowRiskOverCommitment.new(args, fh)
    pod = make_pod(H, {"containers": [{"requests": {"cpu": "1000m", "memory": "1Gi"}, "limits": {"cpu": "2000m"}}]})
    got = [p.score(H.CycleState(), pod, ni)[0] for ni in fh.node_infos]
    G = 1 << 30
    sums = [  # per node: req cpu, req mem, lim cpu, lim mem (limits raised to requests per pod)
        (1500 + 500, 2 * G + 3 * G, 1500 + 3000, 6 * G + 3 * G),
        (1500, 2 * G, 1500, 6 * G),
        (0, 0, 0, 0),
    ]
    want = [oracle.lowrisk_score(35.0, 6.0, 50.0, 4.0, 4000, 8 * G, 7 if i < 2 else 0, *sums[i], 1000, G, 2000, G, 5, 0.3, 0.5)
            for i in range(3)]
    assert got == want and got[2] == 0 and len(set(got)) == 3

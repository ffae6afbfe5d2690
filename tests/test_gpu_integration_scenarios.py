"""The reference's integration scenarios (test/integration/*_test.go: a real scheduler with ONE score plugin enabled,
pods created one after the other, the test asserts which node each pod lands on), replayed with a minimal scheduling
loop around the C++ host mirror -> C-ABI -> CUDA: per pod, the nodes with room for its requests are handed to PreScore,
every one of them is scored, and the pod is bound to the best node.  This is the closest stand-in for SURVEY §8f
rank 3 (envtest acceptance) that runs without a Go toolchain."""
import pytest

from test_gpu_host_plugins import handle_with, make_node, make_pod, watcher

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H(built):
    from scheduler_plugins_b200 import _b200host

    return _b200host


def requests_of(H, pod):
    cpu = mem = 0
    for c in pod.containers:
        cpu += c.requests.get("cpu", 0)
        mem += c.requests.get("memory", 0)
    return cpu, mem


def schedule(H, fh, plugin, pods, on_bind=None):
    """upstream's cycle reduced to what these scenarios exercise: NodeResourcesFit on cpu/memory, the plugin's
    PreScore/Score/NormalizeScore, selectHost (highest score; ties reported as a set)."""
    placed = []
    infos = list(fh.node_infos)
    for pod in pods:
        feasible = []
        for ni in infos:
            used_c = sum(requests_of(H, p)[0] for p in ni.pods)
            used_m = sum(requests_of(H, p)[1] for p in ni.pods)
            c, m = requests_of(H, pod)
            if used_c + c <= ni.node.allocatable.get("cpu", 0) and used_m + m <= ni.node.allocatable.get("memory", 0):
                feasible.append(ni)
        assert feasible, pod.name
        state = H.CycleState()
        assert plugin.pre_score(state, pod, feasible).is_success()
        scored = []
        for ni in feasible:
            s, st = plugin.score(state, pod, ni)
            assert st.is_success()
            scored.append(H.NodeScore(ni.node.name, s))
        if hasattr(plugin, "normalize_score"):
            st, scored = plugin.normalize_score(state, pod, scored)
            assert st.is_success()
        best = max(x.score for x in scored)
        winners = {x.name for x in scored if x.score == best}
        chosen = sorted(winners)[0]
        placed.append((pod.name, chosen, winners))
        for i, ni in enumerate(infos):  # bind: the pod joins the node's pod list, the snapshot moves on
            if ni.node.name == chosen:
                ni.pods = list(ni.pods) + [pod]
                infos[i] = ni
        fh.node_infos = infos
        fh.touch()
        if on_bind:
            on_bind(pod, chosen)
    return placed


@pytest.mark.parametrize("mode,expected", [
    ("Least", {"small": {"fake-node-small-1", "fake-node-small-2"}, "big": {"fake-node-big"}}),
    ("Most", {"small": {"fake-node-big"}, "big": {"fake-node-big"}}),
])
def test_allocatable_integration(H, mode, expected):
    """test/integration/allocatable_test.go:60-113: memory-only weights; small pods on small nodes under Least,
    everything on the big node under Most."""
    nodes = [make_node(H, "fake-node-small-1", {"cpu": "500m", "memory": "500"}),
             make_node(H, "fake-node-small-2", {"cpu": "500m", "memory": "500"}),
             make_node(H, "fake-node-big", {"cpu": "500m", "memory": "5000"})]
    fh = handle_with(H, nodes)
    args = H.NodeResourcesAllocatableArgs()
    args.mode = mode
    args.resources = [H.ResourceSpec("memory", 10)]
    p = H.Allocatable.new(args, fh)
    pods = [make_pod(H, {"containers": [{"requests": {"memory": "100"}}]}, name=f"small-{i}") for i in range(1, 5)]
    pods.append(make_pod(H, {"containers": [{"requests": {"memory": "2000"}}]}, name="big-1"))
    for name, chosen, winners in schedule(H, fh, p, pods):
        assert winners <= expected[name.split("-")[0]], (name, winners)


def trimaran_nodes(H, with_capacity=True):
    return [make_node(H, f"node-{i}", {"cpu": "2", "memory": "256"}) for i in (1, 2, 3)]


def trimaran_pods(H, cpus, limits=None):
    out = []
    for i, c in enumerate(cpus):
        cont = {"requests": {"cpu": f"{c}m", "memory": "50"}}
        out.append(make_pod(H, {"containers": [cont]}, name=f"pod-{i + 1}"))
    return out


def test_target_load_packing_integration(H):
    """test/integration/targetloadpacking_test.go:57-189: utilisation 10 / 60 / 0 % -> both pods on node-1."""
    fh = handle_with(H, trimaran_nodes(H))
    fh.metrics = watcher(H, {"node-1": [("CPU", "Latest", 10.0)], "node-2": [("CPU", "Latest", 60.0)],
                             "node-3": [("CPU", "Latest", 0.0)]})
    p = H.TargetLoadPacking.new(H.TargetLoadPackingArgs(), fh)
    cache = {}

    def on_bind(pod, node):  # PodAssignEventHandler (handler.go:131-167): bound after the metrics window closed
        cache.setdefault(node, []).append(H.ScheduledPodInfo(1, pod))
        fh.scheduled_pods_cache = cache
        fh.touch()

    placed = schedule(H, fh, p, trimaran_pods(H, [300, 100]), on_bind)
    assert [c for _, c, _ in placed] == ["node-1", "node-1"] and all(len(w) == 1 for _, _, w in placed)


def test_load_variation_risk_balancing_integration(H):
    """test/integration/loadVariationRiskBalancing_test.go:55-196: (avg, std) = (30, -) / (70, 20) / (40, 30)."""
    fh = handle_with(H, trimaran_nodes(H))
    fh.metrics = watcher(H, {"node-1": [("CPU", "AVG", 30.0)], "node-2": [("CPU", "AVG", 70.0), ("CPU", "STD", 20.0)],
                             "node-3": [("CPU", "AVG", 40.0), ("CPU", "STD", 30.0)]})
    p = H.LoadVariationRiskBalancing.new(H.LoadVariationRiskBalancingArgs(), fh)
    placed = schedule(H, fh, p, trimaran_pods(H, [300, 100]))
    assert [c for _, c, _ in placed] == ["node-1", "node-1"] and all(len(w) == 1 for _, _, w in placed)


def test_peaks_integration(H):
    """test/integration/peaks_test.go:47-206: three power models of growing |k1|, all nodes idle; the 300m pod goes to
    the flattest model (node-1), the 1900m pod no longer fits there and goes to node-2."""
    fh = handle_with(H, trimaran_nodes(H))
    fh.metrics = watcher(H, {f"node-{i}": [("CPU", "Latest", 0.0)] for i in (1, 2, 3)})
    args = H.PeaksArgs()
    args.node_power_model = {"node-1": H.PowerModel(471.7412504314313, -91.50493019588365, -0.07186049052516228),
                             "node-2": H.PowerModel(471.7412504314313, -1091.50493019588365, -0.07186049052516228),
                             "node-3": H.PowerModel(471.7412504314313, -2091.50493019588365, -0.07186049052516228)}
    p = H.Peaks.new(args, fh)
    placed = schedule(H, fh, p, trimaran_pods(H, [300, 1900]))
    assert [c for _, c, _ in placed] == ["node-1", "node-2"] and all(len(w) == 1 for _, _, w in placed)


def test_low_risk_over_commitment_integration(H):
    """test/integration/lowriskovercommitment_test.go:57-210: two nodes that already run one pod each (limits 500m
    and 1200m), riskLimitWeights 1/1; the pending pod (request 500m, limit 1000m) goes to node-1."""
    nodes = [make_node(H, f"node-{i}", {"cpu": "2", "memory": "256"}) for i in (1, 2)]

    def pod(name, req, lim):
        return make_pod(H, {"containers": [{"requests": {"cpu": f"{req}m", "memory": "64"},
                                            "limits": {"cpu": f"{lim}m", "memory": "64"}}]}, name=name)

    infos = [H.NodeInfo(n) for n in nodes]
    infos[0].pods = [pod("pod-1", 500, 500)]
    infos[1].pods = [pod("pod-2", 100, 1200)]
    fh = H.Handle()
    fh.node_infos = infos
    fh.metrics = watcher(H, {"node-1": [("CPU", "AVG", 60.0), ("CPU", "STD", 30.0)],
                             "node-2": [("CPU", "AVG", 30.0), ("CPU", "STD", 20.0)]})
    args = H.LowRiskOverCommitmentArgs()
    args.risk_limit_weight_cpu = 1.0
    args.risk_limit_weight_memory = 1.0
    p = H.LowRiskOverCommitment.new(args, fh)
    placed = schedule(H, fh, p, [pod("pod-3", 500, 1000)])
    assert placed[0][1] == "node-1" and placed[0][2] == {"node-1"}


# ------------------------------------------------------------------ NodeResourceTopologyMatch, scope = container
def _nrt_integration():
    import json
    import os

    from conftest import GOLDEN

    with open(os.path.join(GOLDEN, "nrt_integration.json")) as f:
        return json.load(f)


NRT_MSG = {"cannot align container", "cannot align init container", "cannot align sidecar container", "cannot align pod"}


@pytest.mark.parametrize("case", _nrt_integration()["cases"], ids=lambda c: c["name"][:70])
def test_topology_match_integration(H, case):
    """test/integration/noderesourcetopology_test.go:1742-2251: Filter + Score (MostAllocated) on two nodes; the pod
    must land on one of the expected nodes, or stay pending with the expected Filter message on every node.  Pods are
    built from LIMITS for Guaranteed cases; the API server defaults requests to limits, restated here."""
    from test_gpu_host_plugins import nrt_handle

    g = _nrt_integration()
    cap = {k: v for k, v in g["node_capacity"].items()}
    nodes = []
    for name, zones in g["nrts"].items():
        nodes.append(dict(name=name, node_extra=cap, policies=[], attributes=g["attributes"],
                          zones=[dict(name=f"node-{i}", resources={r: dict(capacity=q, available=q) for r, q in z.items()})
                                 for i, z in enumerate(zones)]))
    fh = nrt_handle(H, nodes)
    args = H.NodeResourceTopologyMatchArgs()
    args.scoring_strategy = "MostAllocated"
    tm = H.TopologyMatch.new(args, fh)

    def cont(m):
        return {"requests": dict(m), "limits": {} if case["burstable"] else dict(m)}

    pod = make_pod(H, {"init": [cont(m) for m in case["init"]], "containers": [cont(m) for m in case["containers"]]})
    state = H.CycleState()
    feasible, messages = [], []
    for ni in fh.node_infos:
        st = tm.filter(state, pod, ni)
        if st.is_success():
            feasible.append(ni)
        else:
            assert st.code == H.Code.Unschedulable
            messages.append(st.message)
    if not case["expected_nodes"]:
        assert not feasible and any(m.startswith(case["err_msg"]) for m in messages), messages
        return
    assert feasible
    scores = {ni.node.name: tm.score(state, pod, ni)[0] for ni in feasible}
    best = max(scores.values())
    winners = {n for n, s in scores.items() if s == best}
    assert winners <= set(case["expected_nodes"]), (scores, case["expected_nodes"])


def _nrt_integration_full():
    import json
    import os

    from conftest import GOLDEN

    with open(os.path.join(GOLDEN, "nrt_integration_full.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", _nrt_integration_full()["cases"], ids=lambda c: c["name"][:80])
def test_topology_match_integration_full(H, case):
    """test/integration/noderesourcetopology_test.go:239-1741: every scoring strategy (the pod's schedulerName picks
    the profile), pod and container scope, TopologyPolicies and Attributes forms, zone costs for LeastNUMANodes, nodes
    without an NRT object.  Expected: the node the pod lands on (any of the set), or nowhere."""
    g = _nrt_integration_full()
    cap = {k: str(v) for k, v in g["node_capacity"].items()}
    nodes = [make_node(H, n, cap) for n in ("fake-node-1", "fake-node-2")]
    fh = handle_with(H, nodes)
    nrts = {}
    for n in case["nrts"]:
        t = H.NodeResourceTopology()
        t.name = n["name"]
        t.topology_policies = n["policies"]
        t.attributes = n["attributes"]
        zs = []
        for z in n["zones"]:
            zz = H.Zone()
            zz.name, zz.type = z["name"], "Node"
            zz.resources = {r: H.ZoneResource(H.parse_quantity(q["capacity"]), H.parse_quantity(q["available"]))
                            for r, q in z["resources"].items()}
            zz.costs = z["costs"]
            zs.append(zz)
        t.zones = zs
        nrts[n["name"]] = t
    fh.nrts = nrts
    (spec,) = case["pods"]
    args = H.NodeResourceTopologyMatchArgs()
    args.scoring_strategy = spec["strategy"]
    tm = H.TopologyMatch.new(args, fh)

    def cont(c):  # the API server defaults a container's requests to its limits
        req = dict(c["requests"]) or dict(c["limits"])
        return {"requests": req, "limits": dict(c["limits"])}

    pod = make_pod(H, {"init": [cont(c) for c in spec["init"]], "containers": [cont(c) for c in spec["containers"]]})
    state = H.CycleState()
    feasible = [ni for ni in fh.node_infos if tm.filter(state, pod, ni).is_success()]
    if not case["expected_nodes"]:
        assert not feasible
        return
    assert feasible
    scores = {ni.node.name: tm.score(state, pod, ni)[0] for ni in feasible}
    best = max(scores.values())
    winners = {n for n, s in scores.items() if s == best}
    assert winners <= set(case["expected_nodes"]), (scores, case["expected_nodes"])


def test_network_overhead_integration(H):
    """test/integration/networkoverhead_test.go:140-306: AppGroup basic (p1 -> p2 -> p3), NetworkTopology nt-test, eight
    labelled nodes; p1, p2, p3 are created in that order and each must be scheduled (the Go test accepts any node).
    Replayed stronger: PreFilter / Filter / Score / NormalizeScore per pod with the pods placed so far in the lister;
    every node passes Filter, and a pod whose dependency is already placed lands where the cost to it is lowest."""
    from test_gpu_host_plugins import AG, SEL, netoh_fixture

    def cycle(fh, no, sel, placed):
        pod = make_pod(H, {"containers": [{"requests": {"memory": "50"}}]}, name=f"{sel}-test-1", labels={AG: "basic", SEL: sel})
        state = H.CycleState()
        assert no.pre_filter(state, pod, fh.node_infos).is_success()
        feasible = [ni for ni in fh.node_infos if no.filter(state, pod, ni).is_success()]
        assert feasible
        scores = []
        for ni in feasible:
            s, st = no.score(state, pod, ni)
            assert st.is_success()
            scores.append(H.NodeScore(ni.node.name, s))
        st, scores = no.normalize_score(state, pod, scores)
        assert st.is_success()
        best = max(x.score for x in scores)
        winners = sorted(x.name for x in scores if x.score == best)
        pod.node_name = winners[0]
        placed.append(pod)
        fh.pods = placed
        fh.touch()
        return len(feasible), winners

    # the Go test's order: no pod ever finds one of its dependencies placed -> every node passes and ties
    fh, no = netoh_fixture(H, [])
    placed = []
    for sel in ("p1", "p2", "p3"):
        assert cycle(fh, no, sel, placed) == (8, [f"n-{i}" for i in range(1, 9)])
    assert len(placed) == 3
    # dependency-first order (what the AppGroup's topology order exists for): the dependent pod follows its dependency
    fh, no = netoh_fixture(H, [])
    placed, hosts = [], {}
    for sel in ("p3", "p2", "p1"):
        n_feasible, winners = cycle(fh, no, sel, placed)
        hosts[sel] = winners[0]
        if sel == "p3":
            assert n_feasible == 8 and len(winners) == 8
        else:
            assert winners == [hosts["p3" if sel == "p2" else "p2"]], (sel, winners)  # same node: cost 0

"""CPU-side tests of the C++ host mirror: the flattening rules agree with the oracle-side
restatement (oracle/flatten.py — written independently), and plugin construction fails loudly
without a GPU (no CPU fallback)."""
import pytest

from oracle import flatten as F


@pytest.fixture(scope="module")
def H(built):
    from scheduler_plugins_b200 import _b200host

    return _b200host


@pytest.mark.parametrize("q", ["500m", "2", "1Gi", "4G", "128Mi", "1e3", "0", "1.5", "100M", "2Ti", "1500m", "0.25"])
def test_quantity_matches_oracle(H, q):
    assert H.parse_quantity(q) == F.milli(q)


def test_quantity_rejects_sub_milli(H):
    with pytest.raises(ValueError):
        H.parse_quantity("100u")
    with pytest.raises(ValueError):
        H.parse_quantity("abc")


def mkpod(H, spec):
    p = H.Pod()
    def conts(lst):
        out = []
        for c in lst:
            k = H.Container()
            k.requests = H.resource_list({a: str(b) for a, b in c.get("requests", {}).items()})
            k.limits = H.resource_list({a: str(b) for a, b in c.get("limits", {}).items()})
            k.restart_always = bool(c.get("restart_always", False))
            out.append(k)
        return out
    p.init_containers = conts(spec.get("init", []))
    p.containers = conts(spec.get("containers", []))
    if spec.get("overhead") is not None:
        p.has_overhead = True
        p.overhead = H.resource_list({a: str(b) for a, b in spec["overhead"].items()})
    return p


PODS = [
    {"containers": [{"requests": {"cpu": "1", "memory": "1Gi"}, "limits": {"cpu": "1", "memory": "1Gi"}}]},
    {"containers": [{"requests": {"cpu": "1", "memory": "1Gi"}, "limits": {"cpu": "2", "memory": "1Gi"}}]},
    {"containers": [{"requests": {"vendor/nic1": "1"}, "limits": {"vendor/nic1": "1"}}]},
    {"containers": [{"requests": {}, "limits": {"cpu": "1", "memory": "1Gi"}}]},
    {},
    {"init": [{"requests": {"cpu": "4", "memory": "1Gi"}}],
     "containers": [{"requests": {"cpu": "1", "memory": "2Gi"}, "limits": {"cpu": "2"}}, {"requests": {"cpu": "500m"}}],
     "overhead": {"cpu": "100m"}},
    {"init": [{"requests": {"cpu": "2", "memory": "3Gi", "hugepages-2Mi": "64Mi"}, "limits": {"cpu": "2", "memory": "3Gi"}}],
     "containers": [{"requests": {"cpu": "1", "memory": "1Gi"}, "limits": {"cpu": "1", "memory": "1Gi"}}] * 2},
]


@pytest.mark.parametrize("spec", PODS, ids=range(len(PODS)))
def test_pod_rules_match_oracle(H, spec):
    pod = mkpod(H, spec)
    assert H.pod_qos(pod) == F.pod_qos(spec)
    assert dict(H.pod_effective_request(pod)) == F.pod_effective_request(spec)
    assert H.pod_predicted_cpu(pod, 1000, 1.5) == F.tlp_pod_cpu(spec, 1000, 1.5)


def test_no_cpu_fallback(H):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError) as ei:
        H.Allocatable.new(None, H.Handle())
    assert "no CPU fallback" in str(ei.value)


def test_handle_node_change_log(H):
    """Per-node generations (upstream NodeInfo.Generation): a plugin may patch rows only when every change since
    its snapshot is a logged single-node change."""
    h = H.Handle()
    n = H.Node()
    n.name = "n0"
    h.node_infos = [H.NodeInfo(n) for _ in range(4)]
    g0 = h.generation
    assert h.nodes_changed_since(g0) == []
    h.touch_node(2); h.touch_node(0); h.touch_node(2)
    assert h.nodes_changed_since(g0) == [0, 2]
    assert h.nodes_changed_since(g0 + 1) == [0, 2]
    assert h.nodes_changed_since(g0 + 2) == [2]
    h.touch_node(7)                                  # index outside the list: not patchable
    assert h.nodes_changed_since(g0) is None
    h.touch()                                        # list-level change: everything before it is stale
    g1 = h.generation
    assert h.nodes_changed_since(g0) is None and h.nodes_changed_since(g1) == []
    h.touch_node(1)
    h.generation = h.generation + 1                  # moved without a log entry
    assert h.nodes_changed_since(g1) is None
    assert h.nodes_changed_since(h.generation + 5) is None

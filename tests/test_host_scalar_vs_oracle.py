"""The host's scalar NodeResourceTopologyMatch path (host/nrt_scalar.cpp -- where reason code 9 goes) against the C
oracle on seeded synthetic snapshots, pair by pair: Filter verdict + message and the score of all four strategies.

The two restatements share nothing: the oracle works on the dense encoding (zone / resource slots, masks, milli-unit
columns), the scalar path on named zones and ResourceLists as the Go plugin does.  The dense columns of
synth.gen_nrt are turned back into objects here (zones `node-<i>` in id order, so both apply) -- every pair the
reference would hand to a handler must agree.  CPU only."""
import pytest

from scheduler_plugins_b200 import synth

NAMES = ["cpu", "memory", "hugepages-2Mi", "vendor/nic1"]  # the slots of synth.gen_nrt: affine x3, host-level device
REASON_MSG = {0: None, 2: "cannot align pod", 3: "cannot align container", 4: "cannot align init container",
              5: "cannot align sidecar container"}
F_HAS_NRT, F_FRESH, F_SINGLE, F_SCOPE_POD, F_UNSUPPORTED = 1, 2, 4, 8, 16


@pytest.fixture(scope="module")
def H(built):
    from scheduler_plugins_b200 import _b200host

    return _b200host


def qty(milli):
    return f"{int(milli)}m"


def node_objects(H, nodes, n):
    """NodeInfo + NodeResourceTopology of node n from the dense columns"""
    Z, R = nodes["n_zones"], nodes["n_res"]
    t = H.NodeResourceTopology()
    t.name = f"n{n}"
    scope = "pod" if nodes["node_flags"][n] & F_SCOPE_POD else "container"
    t.attributes = {"topologyManagerPolicy": "single-numa-node", "topologyManagerScope": scope,
                    "topologyManagerMaxNUMANodes": str(int(nodes["max_numa"][n]))}
    zones = []
    for z in range(int(nodes["n_zones_node"][n])):
        zz = H.Zone()
        zz.name, zz.type = f"node-{z}", "Node"
        res = {}
        for r in range(R):
            if (int(nodes["zone_res_mask"][z, n]) >> r) & 1:
                q = H.parse_quantity(qty(nodes["avail"][z, r, n]))
                res[NAMES[r]] = H.ZoneResource(q, q)
        zz.resources = res
        if nodes.get("cost") is not None:
            zz.costs = {f"node-{b}": int(nodes["cost"][z, b, n]) for b in range(Z) if int(nodes["cost"][z, b, n]) >= 0}
        zones.append(zz)
    t.zones = zones
    nd = H.Node()
    nd.name = t.name
    alloc = {NAMES[r]: "1000" for r in range(R) if (int(nodes["node_res_mask"][n]) >> r) & 1}  # presence is what matters
    nd.allocatable = H.resource_list(alloc)
    nd.capacity = H.resource_list(alloc)
    return H.NodeInfo(nd), t


def pod_object(H, pods, p):
    from test_host_cpu import mkpod

    C = pods["cont_kind"].shape[1]
    qos = int(pods["qos"][p])
    ni, na = int(pods["n_init"][p]), int(pods["n_app"][p])

    def cont(c):
        req = {NAMES[r]: qty(pods["req"][p, c, r]) for r in range(4) if (int(pods["req_mask"][p, c]) >> r) & 1}
        out = {"requests": req, "restart_always": int(pods["cont_kind"][p, c]) == 2}
        if qos == 0:
            out["limits"] = dict(req)  # Guaranteed: limits == requests for every resource
        return out

    assert ni + na <= C
    return mkpod(H, {"init": [cont(c) for c in range(ni)], "containers": [cont(c) for c in range(ni, ni + na)]})


@pytest.mark.parametrize("seed,Z", [(101, 4), (102, 2), (103, 3), (104, 1)])
def test_scalar_path_agrees_with_the_oracle_pair_by_pair(H, oracle, seed, Z):
    from oracle import pyoracle_nrt

    N, P = 48, 40
    nodes, pods = synth.gen_nrt(seed, N, P, Z=Z)
    weights = [3, 1, 2, 1]
    wmap = {NAMES[r]: weights[r] for r in range(4)}
    per_strategy = {s: pyoracle_nrt.nrt_batch(nodes, pods, s, weights, None, pitch=128) for s in (0, 1, 2, 3)}
    node_objs = {}
    checked = filtered = scored = 0
    for p in range(P):
        if int(pods["flags"][p]) & 3:  # Filter bypass (BestEffort, native only) / unsupported pod: decided before any handler
            continue
        pod = pod_object(H, pods, p)
        assert int(H.pod_qos(pod)) == int(pods["qos"][p]), p  # the object carries the QoS class the columns say
        for n in range(N):
            fl = int(nodes["node_flags"][n])
            if (fl & (F_HAS_NRT | F_FRESH | F_SINGLE)) != (F_HAS_NRT | F_FRESH | F_SINGLE) or (fl & F_UNSUPPORTED):
                continue  # the gates before the handlers (filter.go:194-209) are not the scalar path's business
            if n not in node_objs:
                node_objs[n] = node_objects(H, nodes, n)
            ni, nrt = node_objs[n]
            st = H.scalar_filter(pod, ni, nrt)
            reason = int(per_strategy[2][2][p, n])
            assert reason in REASON_MSG, (p, n, reason)
            assert (None if st.is_success() else st.message) == REASON_MSG[reason], (seed, p, n, reason, st.message)
            checked += 1
            filtered += reason != 0
            if reason != 0:
                continue
            for s in (0, 1, 2, 3):
                want = int(per_strategy[s][0][p, n])
                got = int(H.scalar_score(pod, nrt, s, wmap)) if int(pods["qos"][p]) == 0 else 100  # score.go:72-75
                assert got == want, (seed, p, n, s, got, want)
                scored += 1
    assert checked > 500 and filtered > 20 and scored > 1000, (checked, filtered, scored)

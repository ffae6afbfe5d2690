"""least_numa_test.go's helper-level vectors (TestNormalizeScore :706-756, TestMinDistance :758-912)."""
import ctypes as C

import numpy as np
import pytest


@pytest.mark.parametrize("count,is_min,want", [(1, 0, 88), (2, 0, 76), (8, 0, 4), (1, 1, 94), (2, 1, 82), (8, 1, 10)])
def test_normalize_score(oracle, count, is_min, want):
    lib = oracle.lib()
    lib.orc_nrt_normalize_least_numa.restype = C.c_int64
    assert lib.orc_nrt_normalize_least_numa(count, is_min, 8) == want  # nodeconfig.DefaultMaxNUMANodes = 8


COSTS = [[10, 12, 20, 20], [12, 10, 20, 20], [20, 20, 10, 12], [20, 20, 12, 10]]


@pytest.mark.parametrize("combos,cost,want", [
    ([[0], [1], [2], [3]], COSTS, 10.0),
    ([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], COSTS, 11.0),
    ([[0, 1, 2], [1, 2, 3], [0, 2, 3]], COSTS, np.float32(14.888889)),
    ([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], [[-1] * 4] * 4, 255.0),  # no Costs: maxDistanceValue
])
def test_min_avg_distance(oracle, combos, cost, want):
    lib = oracle.lib()
    lib.orc_nrt_min_avg_distance.restype = C.c_float
    c = np.ascontiguousarray(cost, dtype=np.int32)
    cb = np.ascontiguousarray(combos, dtype=np.int32)
    got = lib.orc_nrt_min_avg_distance(C.c_void_p(c.ctypes.data), 4, C.c_void_p(cb.ctypes.data), len(combos), len(combos[0]))
    assert np.float32(got) == np.float32(want)  # the Go test compares float32 with !=


# ------------------------------------------------------------------ TestNUMANodesRequired (least_numa_test.go:35-704)
def _nnr_cases():
    import json
    import os

    from conftest import GOLDEN

    with open(os.path.join(GOLDEN, "numa_nodes_required.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", _nnr_cases(), ids=lambda c: c["name"][:70])
def test_numa_nodes_required_vectors(oracle, case):
    """numaNodesRequired: the smallest zone combination that hosts the request, and whether its average distance is the
    minimum for that size.  Lists whose NUMA ids are not 0..k-1 in order are outside the dense encoding: the host
    flags such a node UNSUPPORTED and keeps the Go path, which is asserted instead."""
    import ctypes as C

    import numpy as np

    from oracle import flatten as F

    ids = [z["id"] for z in case["zones"]]
    nrt = {"zones": [{"name": f"node-{z['id']}", "type": "Node", "resources": {}} for z in case["zones"]]}
    zones, got_ids, supported = F.numa_zones(nrt)
    assert got_ids == ids and supported == (ids == list(range(len(ids))))
    if not supported:
        return
    names = ["cpu", "memory", "gpu"]

    class Node(C.Structure):
        _fields_ = [("flags", C.c_uint8), ("max_numa", C.c_uint16), ("n_zones", C.c_uint8), ("node_res_mask", C.c_uint8),
                    ("zone_res_mask", C.c_uint8 * 8), ("avail", (C.c_int64 * 8) * 8), ("cost", (C.c_int32 * 8) * 8)]

    nd = Node()
    nd.n_zones = len(ids)
    for a in range(8):
        for b in range(8):
            nd.cost[a][b] = -1
    for z, zone in enumerate(case["zones"]):
        for r, n in enumerate(names):
            if n in zone["resources"]:
                nd.zone_res_mask[z] |= 1 << r
                nd.avail[z][r] = F.milli(zone["resources"][n])
        for dst, c in zone["costs"].items():
            nd.cost[z][int(dst)] = c
    req = (C.c_int64 * 8)()
    mask = 0
    for r, n in enumerate(names):
        if n in case["pod"]:
            req[r] = F.milli(case["pod"][n])
            mask |= 1 << r
    res_flags = (C.c_uint8 * 3)(1, 1, 2)  # cpu, memory NUMA-affine; gpu a device
    out_mask, is_min = C.c_uint32(0), C.c_int(0)
    fn = oracle.lib().orc_nrt_numa_nodes_required
    k = fn(C.byref(nd), res_flags, C.c_int(3), C.c_int(0), C.c_uint8(mask), req, C.byref(out_mask), C.byref(is_min))
    if case["expected_bits"] is None:
        assert k == 0
        return
    assert k == len(case["expected_bits"])
    assert out_mask.value == sum(1 << b for b in case["expected_bits"])
    assert bool(is_min.value) == case["expected_min_distance"]


# ------------------------------------------------------------------ TestOnlyNonNUMAResources (pluginhelpers_test.go:28-106)
def _only_non_numa():
    import json
    import os

    from conftest import GOLDEN

    with open(os.path.join(GOLDEN, "only_non_numa.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", _only_non_numa()["cases"], ids=lambda c: c["name"])
def test_only_non_numa_resources_vectors(oracle, case):
    """onlyNonNUMAResources on the dense encoding: a resource slot per name that occurs in the zones or the request, the
    zones' resource masks, the request mask -- true iff no zone lists any requested resource."""
    g = _only_non_numa()
    names = sorted({r for z in g["zones"] for r in z["resources"]} | set(case["resources"]))
    assert len(names) <= 8
    zmask = (C.c_uint8 * 8)()
    for z, zone in enumerate(g["zones"]):
        for r, n in enumerate(names):
            if n in zone["resources"]:
                zmask[z] |= 1 << r
    req_mask = sum(1 << r for r, n in enumerate(names) if n in case["resources"])
    got = oracle.lib().orc_nrt_only_non_numa(zmask, C.c_int(len(g["zones"])), C.c_uint8(req_mask), C.c_int(len(names)))
    assert bool(got) == case["expected"]

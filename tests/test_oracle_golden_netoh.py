"""Pins the NetworkOverhead oracle against networkoverhead_test.go's own expectations."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

MISSING = -(2**63)


def load():
    with open(os.path.join(GOLDEN, "network_overhead.json")) as f:
        return json.load(f)


def matrices(g):
    K = len(g["names"])
    zc = np.full((K, K), MISSING, dtype=np.int64)
    rc = np.full((K, K), MISSING, dtype=np.int64)
    for o, d, c in g["zone_cost"]:
        zc[o, d] = c
    for o, d, c in g["region_cost"]:
        rc[o, d] = c
    return zc, rc


def deps_of(case):
    return [(d["host_node"], d["host_region"], d["host_zone"], d["max_network_cost"]) for d in case["deps"]]


@pytest.mark.parametrize("case", load()["score_cases"], ids=lambda c: c["name"])
def test_score_and_normalize(oracle, case):
    g = load()
    zc, rc = matrices(g)
    raw = []
    for n in range(8):
        if case["score_equally"]:
            raw.append(0)  # Score returns MinNodeScore (networkoverhead.go:376-378)
        else:
            raw.append(oracle.netoh_node(zc, rc, n, g["node_region"][n], g["node_zone"][n], deps_of(case))[2])
    assert raw == case["raw"]
    assert list(oracle.netoh_normalize(raw)) == case["normalized"]


@pytest.mark.parametrize("case", load()["filter_cases"], ids=lambda c: c["name"])
def test_filter(oracle, case):
    g = load()
    zc, rc = matrices(g)
    n = case["node"]
    sat, viol, _ = oracle.netoh_node(zc, rc, n, g["node_region"][n], g["node_zone"][n], deps_of(case))
    assert (sat, viol) == (case["satisfied"], case["violated"])
    passed = bool(case["score_equally"]) or not (viol > sat)
    assert passed == case["pass"]


def test_missing_entries_and_empty_labels(oracle):
    """Filter and Score treat a missing matrix entry differently (networkoverhead.go:548-555 vs
    :617-621); 'same region' is string equality incl. the empty string (:540)."""
    g = load()
    zc, rc = matrices(g)
    # node in us-west-1/Z1, dependency hosted in us-west-1/Z9 (id 4 reused as unknown dest: Z1->Z3 missing)
    sat, viol, cost = oracle.netoh_node(zc, rc, 0, 1, 3, [(5, 1, 5, 0)])
    assert (sat, viol, cost) == (0, 0, 100)  # neither counted, MaxCost added
    # host has no region and no zone label -> violated, MaxCost
    assert oracle.netoh_node(zc, rc, 0, 1, 3, [(5, 0, 0, 50)]) == (0, 1, 100)
    # both nodes lack a region label but share a zone -> "same region" ('' == '') and same zone
    assert oracle.netoh_node(zc, rc, 0, 0, 3, [(5, 0, 3, 0)]) == (1, 0, 1)
    # cost within MaxNetworkCost is satisfied
    assert oracle.netoh_node(zc, rc, 0, 1, 3, [(5, 2, 5, 20)]) == (1, 0, 20)
    assert oracle.netoh_node(zc, rc, 0, 1, 3, [(5, 2, 5, 19)]) == (0, 1, 20)


def test_region_zone_name_collision(oracle):
    """Region and zone names live in one costMap namespace (networkoverhead.go:472-493): a node whose
    zone label equals its region label sees the region-origin entries under the zone key too."""
    K = 4  # 0 '', 1 'A' (used as region AND zone), 2 'B', 3 'C'
    zc = np.full((K, K), MISSING, dtype=np.int64)
    rc = np.full((K, K), MISSING, dtype=np.int64)
    rc[1, 2] = 7   # region list: A -> B
    zc[1, 3] = 9   # zone list:   A -> C
    # node (region A, zone A); dependency host in region A, zone B -> zone lookup (A,B): only the
    # region list has it, and it is visible because origin strings are equal
    assert oracle.netoh_node(zc, rc, 0, 1, 1, [(9, 1, 2, 100)]) == (1, 0, 7)
    # node (region A, zone C'): zone differs from region -> not visible
    zc2 = zc.copy()
    assert oracle.netoh_node(zc2, rc, 0, 1, 3, [(9, 1, 2, 100)]) == (0, 0, 100)
    # region lookup (A,C) for a host in region C: zone entry (A,C) overwrote/filled the key
    assert oracle.netoh_node(zc, rc, 0, 1, 1, [(9, 3, 2, 100)]) == (1, 0, 9)


def test_normalize_edges(oracle):
    assert list(oracle.netoh_normalize([])) == []
    assert list(oracle.netoh_normalize([0, 0])) == [0, 0]        # early return :400-402
    assert list(oracle.netoh_normalize([5, 5, 5])) == [100, 100, 100]  # max == min != 0 :411-413
    assert list(oracle.netoh_normalize([0, 3])) == [100, 0]
    assert list(oracle.netoh_normalize([1, 2, 4])) == [100, 100 - 33, 0]

"""Replays score_test.go's expectations (tests/golden/nrt_score.json) through the oracle."""
import json
import os

import pytest

from conftest import GOLDEN
from oracle import flatten as F
from test_oracle_golden_nrt import node_objects

STRATEGY = {"MostAllocated": 0, "BalancedAllocation": 1, "LeastAllocated": 2, "LeastNUMANodes": 3}


def load():
    with open(os.path.join(GOLDEN, "nrt_score.json")) as f:
        return json.load(f)


def with_policy(nodes, policy):
    if not policy:
        return nodes
    return [dict(n, policies=[policy]) for n in nodes]


def scores_for(oracle, nodes_fixture, pod, strategy):
    from oracle import pyoracle_nrt

    nodes, nrts = node_objects({"nodes": nodes_fixture})
    names = F.build_dictionary([pod])
    ns = F.flatten_nrt_nodes(nodes, nrts, names)
    ps = F.flatten_nrt_pods([pod], names)
    # Score is called directly by the Go tests (no Filter first): read the score of every node
    # through the batch driver with the filter verdict ignored -> use per-node evaluation
    sc, feas, reasons = pyoracle_nrt.nrt_batch(ns, ps, STRATEGY[strategy])
    return {n["name"]: int(sc[0, i]) for i, n in enumerate(nodes_fixture)}, reasons[0]


@pytest.mark.parametrize("case", load()["suites"][0]["cases"], ids=lambda c: c["name"])
def test_strategy_scores(oracle, case):
    s = load()["suites"][0]
    nodes = with_policy(s["nodes"], s["policy_override"])
    got, reasons = scores_for(oracle, nodes, case["pod"], case["strategy"])
    (want_node, want_score), = case["want_max"].items()
    assert got[want_node] == want_score
    assert max(got.values()) == want_score  # findMaxScoreNode picks it (score_test.go:622-633)


@pytest.mark.parametrize("case", load()["suites"][1]["cases"], ids=lambda c: c["name"])
def test_least_numa_scores(oracle, case):
    s = load()["suites"][1]
    nodes = with_policy(s["fixtures"][case["fixture"]], case["policy_override"])
    got, reasons = scores_for(oracle, nodes, case["pod"], "LeastNUMANodes")
    # the batch driver scores only nodes that pass Filter; these fixtures use best-effort policies
    # (no filter handler) so every node is scored
    assert not reasons[:len(nodes)].any()
    assert got == case["want"]


PARTIAL = [  # TestNodeResourcePartialDataScorePlugin, score_test.go:483-600: (strategy, nodes that keep an NRT, wanted)
    ("MostAllocated", (), {}), ("LeastAllocated", (), {}), ("BalancedAllocation", (), {}),
    ("MostAllocated", ("Node1",), {"Node1": 27}), ("LeastAllocated", ("Node1",), {"Node1": 73}),
    ("BalancedAllocation", ("Node1",), {"Node1": 89}),
]


@pytest.mark.parametrize("strategy,keep,want", PARTIAL, ids=lambda v: str(v))
def test_partial_nrt_data_scores(oracle, strategy, keep, want):
    """Nodes whose NRT object is missing score 0 (score.go:83-86); the node that has one is elected with the expected
    score.  Fixture and pod: defaultNUMANodes + Pod1 of score_test.go (cpu 2, memory 20 MiB, Guaranteed)."""
    from oracle import pyoracle_nrt

    s = load()["suites"][0]
    fixture = with_policy(s["nodes"], s["policy_override"])
    pod = s["cases"][0]["pod"]
    nodes, nrts = node_objects({"nodes": fixture})
    nrts = [t if n["name"] in keep else None for n, t in zip(fixture, nrts)]
    names = F.build_dictionary([pod])
    sc, feas, reasons = pyoracle_nrt.nrt_batch(F.flatten_nrt_nodes(nodes, nrts, names), F.flatten_nrt_pods([pod], names),
                                               STRATEGY[strategy])
    got = {n["name"]: int(sc[0, i]) for i, n in enumerate(fixture)}
    assert not reasons[0][:len(fixture)].any()                # a node without NRT data passes Filter (filter.go:198-200)
    for name in got:
        if name not in keep:
            assert got[name] == 0
    for name, score in want.items():
        assert got[name] == score and max(got.values()) == score

"""Property tests of the oracle itself (hypothesis): invariants the reference's arithmetic must satisfy for ANY input,
complementing the fixed golden vectors.  They guard the checker -- a broken oracle would silently bless a broken
kernel."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from scheduler_plugins_b200 import synth

i63 = st.integers(min_value=0, max_value=(1 << 40))


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 256_000), st.integers(0, 1 << 38)), min_size=1, max_size=40), st.sampled_from([0, 1]))
def test_allocatable_normalize_properties(oracle, nodes, mode):
    """allocatable.go:143-168: scores land in [0, 100], keep the order of the raw scores, the best node gets 100 unless
    every node ties (then all get 0)."""
    w = [1 << 20, 1]
    raw = [oracle.alloc_score(list(n), w, mode) for n in nodes]
    got = list(oracle.alloc_normalize(raw))
    assert all(0 <= s <= 100 for s in got)
    if max(raw) == min(raw):
        assert set(got) == {0}
    else:
        assert got[int(np.argmax(raw))] == 100 and got[int(np.argmin(raw))] == 0
        order = np.argsort(raw, kind="stable")
        assert all(got[a] <= got[b] for a, b in zip(order, order[1:]))
    # the batch driver agrees with score + normalize
    cols = [np.array([n[r] for n in nodes], dtype=np.int64) for r in range(2)]
    assert list(oracle.alloc_batch(cols, w, mode, 1)[0]) == got


@settings(max_examples=200, deadline=None)
@given(st.floats(0, 100), st.integers(0, 256_000), st.integers(0, 64_000), st.integers(0, 16_000), st.integers(1, 99))
def test_tlp_score_properties(oracle, util, cap, missing, pod_cpu, target):
    """targetloadpacking.go:146-186: the score is in [0, 100]; at or below the target it never decreases with the
    predicted utilisation, above it it never increases; a node without metrics scores 0."""
    s = oracle.tlp_score(util, cap, missing, 3, pod_cpu, target)
    assert 0 <= s <= 100
    assert oracle.tlp_score(util, cap, missing, 0, pod_cpu, target) == 0
    if cap > 0:
        pred = lambda extra: 100 * ((util / 100) * cap + pod_cpu + missing + extra) / cap  # noqa: E731
        s2 = oracle.tlp_score(util, cap, missing + 100, 3, pod_cpu, target)
        if pred(100) <= target:
            assert s2 >= s
        elif pred(0) > target:
            assert s2 <= s


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(-(1 << 40), 1 << 40), min_size=0, max_size=30))
def test_peaks_normalize_properties(oracle, scores):
    """peaks.go:152-168: lowest raw score -> 100, highest -> 0 (unless all equal), order reversed, all-zero untouched."""
    got = list(oracle.peaks_normalize(scores))
    if not scores:
        return
    if min(scores) == 0 and max(scores) == 0:
        assert got == scores
    elif min(scores) == max(scores):
        assert set(got) == {100}
    else:
        assert got[int(np.argmin(scores))] == 100 and got[int(np.argmax(scores))] == 0
        assert all(0 <= s <= 100 for s in got)
        order = np.argsort(scores, kind="stable")
        assert all(got[a] >= got[b] for a, b in zip(order, order[1:]))


@settings(max_examples=120, deadline=None)
@given(st.floats(0.1, 99.9), st.floats(0.0, 40.0), st.integers(100, 64_000), st.integers(0, 64_000), st.integers(0, 200_000),
       st.integers(1, 20))
def test_lowrisk_risk_load_properties(oracle, util, std, cap, req, lim, window):
    """lowriskovercommitment.go:213-249: a probability complement -- always in [0, 1]; zero deviation gives a step."""
    r = oracle.lowrisk_risk_load(True, util, std, float(cap), cap, min(req, cap), lim, window)
    assert 0.0 <= r <= 1.0
    step = oracle.lowrisk_risk_load(True, util, 0.0, float(cap), cap, min(req, cap), 10 ** 9, window)
    assert step in (0.0, 1.0)
    assert step == (0.0 if util / 100 <= min(req, cap) / cap else 1.0)


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(0, 3))
def test_nrt_pod_scope_more_capacity_never_rejects(oracle, seed, strategy):
    """singleNUMAPodLevelHandler (filter.go:162-173) is ONE resourcesAvailableInAnyNUMANodes call on the
    pod-effective request, and availability only enters it through `available >= request` (:121-136): on a
    pod-scope node raising every zone's availability can only turn rejects into passes."""
    from oracle import pyoracle_nrt

    N, P = 40, 12
    nodes, pods = synth.gen_nrt(seed, N, P, Z=4)
    _, f0, r0 = pyoracle_nrt.nrt_batch(nodes, pods, strategy, None, None, pitch=N)
    roomy = dict(nodes)
    roomy["avail"] = nodes["avail"] * 2 + (nodes["avail"] > 0) * 1000
    _, f1, r1 = pyoracle_nrt.nrt_batch(roomy, pods, strategy, None, None, pitch=N)
    pod_scope = (nodes["node_flags"] & 8) != 0
    passed0 = (r0[:, :N] == 0) & pod_scope[None, :]
    assert (r1[:, :N][passed0] == 0).all()


def test_nrt_container_scope_is_not_monotone_in_capacity(oracle):
    """The container-scope handler (filter.go:39-78) places each app container greedily on the LOWEST fitting
    NUMA id (:154) and subtracts it there (:68-74), so more capacity can move an early container onto the zone a
    later one needed: seed 272, pod 3 / node 1 passes on the original node and is rejected on the doubled one.
    Pins the counter-example that refuted the (wrong) monotonicity property this file used to assert."""
    from oracle import pyoracle_nrt

    N, P = 40, 12
    nodes, pods = synth.gen_nrt(272, N, P, Z=4)
    _, _, r0 = pyoracle_nrt.nrt_batch(nodes, pods, 0, None, None, pitch=N)
    roomy = dict(nodes)
    roomy["avail"] = nodes["avail"] * 2 + (nodes["avail"] > 0) * 1000
    _, _, r1 = pyoracle_nrt.nrt_batch(roomy, pods, 0, None, None, pitch=N)
    flipped = (r0[:, :N] == 0) & (r1[:, :N] != 0)
    assert flipped.any()
    assert not (flipped & ((nodes["node_flags"] & 8) != 0)[None, :]).any()  # only container-scope nodes flip

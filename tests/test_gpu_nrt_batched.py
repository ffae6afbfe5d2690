"""GPU parity of the BATCHED NodeResourceTopologyMatch path (csrc/nrt2.cu: one score / pod-scope-filter table per
distinct request vector, gcd-scaled 32-bit arithmetic, coalesced natural-order expansion) against the oracle and
against the direct per-(pod, node) kernel -- including the cases where it must decline and hand over to the direct
kernel, and the BASELINE config-4 shape (5 000 pods x 20 000 nodes x 4 zones)."""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth
from test_gpu_nrt import run_nrt

pytestmark = pytest.mark.gpu


def oracle_rows(nodes, pods, rows, strategy, w, feas, pitch):
    from oracle import pyoracle_nrt

    sub = {k: (v[rows] if isinstance(v, np.ndarray) else v) for k, v in pods.items()}
    return pyoracle_nrt.nrt_batch(nodes, sub, strategy, w, None if feas is None else feas[rows], pitch=pitch)


@pytest.mark.parametrize("strategy", [0, 1, 2])
def test_batched_equals_oracle_many_shapes(eng, engine_mod, strategy):
    """ragged shapes: N not a multiple of the 256-node tile, P not a multiple of the 32-pod tile, 1..4 zones."""
    from oracle import pyoracle_nrt

    E = engine_mod
    for seed, P, N, Z in ((11, 32, 1, 1), (12, 65, 129, 2), (13, 100, 255, 4), (14, 37, 513, 4), (15, 260, 3000, 2),
                          (16, 64, 300, 3)):
        nodes, pods = synth.gen_nrt(seed, N, P, Z=Z)
        feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N)) if seed % 2 else None
        w = [1, 2, 3, 4]
        gs, gf, gr = run_nrt(eng, E, nodes, pods, strategy, w, feas, path=E.NRT_PATH_BATCHED)
        if Z == 3:  # a third of a MiB-granular node is byte-granular: the common unit is 1 byte and 2^38 does not fit
            assert eng.nrt_last_path() == E.NRT_PATH_DIRECT and "does not fit" in eng.nrt_path_note()
        else:
            assert eng.nrt_last_path() == E.NRT_PATH_BATCHED, eng.nrt_path_note()
        ws, wf, wr = pyoracle_nrt.nrt_batch(nodes, pods, strategy, w, feas, pitch=eng.Npad)
        assert np.array_equal(gr, wr), (seed, np.argwhere(gr != wr)[:5])
        assert np.array_equal(gf, wf), seed
        assert np.array_equal(gs, ws), (seed, np.argwhere(gs != ws)[:5])


@pytest.mark.parametrize("strategy", [0, 2])
def test_batched_wide_value_ratios(eng, engine_mod, strategy):
    """page-granular (4 KiB) memory: zone capacities in pages exceed 2^32 / 100, the table kernel switches to the
    64-bit numerator form of (cv - rv) * 100 / cv -- same scores."""
    from oracle import pyoracle_nrt

    E = engine_mod
    P, N = 80, 1100
    nodes, pods = synth.gen_nrt(23, N, P, Z=4)
    nodes = dict(nodes, avail=nodes["avail"].copy())
    listed = ((nodes["zone_res_mask"] >> 1) & 1).astype(np.int64)
    rng = np.random.default_rng(5)
    nodes["avail"][:, 1] = (nodes["avail"][:, 1] * 8 + rng.integers(0, 256, listed.shape) * 4096 * 1000) * listed
    w = [1, 3, 1, 2]
    gs, gf, gr = run_nrt(eng, E, nodes, pods, strategy, w, None, path=E.NRT_PATH_BATCHED)
    assert eng.nrt_last_path() == E.NRT_PATH_BATCHED and "64-bit" in eng.nrt_path_note(), eng.nrt_path_note()
    ws, wf, wr = pyoracle_nrt.nrt_batch(nodes, pods, strategy, w, None, pitch=eng.Npad)
    assert np.array_equal(gr, wr) and np.array_equal(gf, wf) and np.array_equal(gs, ws)


@pytest.mark.parametrize("strategy", [0, 2])  # MostAllocated, LeastAllocated
def test_batched_quotient_tables_and_per_vector_tables_agree_with_oracle(eng, engine_mod, strategy):
    """Least / MostAllocated tables are built from per-(request value, cell) quotient tables when the weight sum keeps
    100 x sum inside a 16-bit lane (<= 655); larger weights keep the per-vector table kernel.  Both forms, ragged
    row / slot counts, with and without an upstream mask -- against the oracle."""
    from oracle import pyoracle_nrt

    E = engine_mod
    seen = set()
    for seed, P, N, Z, w in ((41, 130, 700, 4, [1, 1, 1, 1]), (42, 97, 391, 2, [5, 3, 2, 1]), (43, 130, 700, 4, [300, 400, 1, 1]),
                             (44, 64, 129, 4, [160, 160, 160, 175]), (45, 64, 129, 4, [164, 164, 164, 164])):
        nodes, pods = synth.gen_nrt(seed, N, P, Z=Z)
        feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N)) if seed % 2 else None
        gs, gf, gr = run_nrt(eng, E, nodes, pods, strategy, w, feas, path=E.NRT_PATH_BATCHED)
        assert eng.nrt_last_path() == E.NRT_PATH_BATCHED, eng.nrt_path_note()
        want_q = max(sum(w[r] for r in range(4) if (int(m) >> r) & 1) for m in np.unique(pods["req_mask"])) <= 655
        assert ("quotient tables" in eng.nrt_path_note()) == want_q, (seed, eng.nrt_path_note())
        seen.add(want_q)
        ws, wf, wr = pyoracle_nrt.nrt_batch(nodes, pods, strategy, w, feas, pitch=eng.Npad)
        assert np.array_equal(gr, wr), (seed, np.argwhere(gr != wr)[:5])
        assert np.array_equal(gf, wf), seed
        assert np.array_equal(gs, ws), (seed, np.argwhere(gs != ws)[:5])
    assert seen == {True, False}


def test_batched_declines_what_it_cannot_encode(eng, engine_mod):
    """Quantities that do not fit the scaled 32-bit encoding, absurd weights, a Guaranteed pod that names a single
    resource under BalancedAllocation (NaN variance) and LeastNUMANodes all keep the direct kernel -- with the
    same results as the oracle."""
    from oracle import pyoracle_nrt

    E = engine_mod
    P, N = 64, 700
    nodes, pods = synth.gen_nrt(31, N, P, Z=4)

    def check(nd, pd, strategy, w, want_path):
        gs, gf, gr = run_nrt(eng, E, nd, pd, strategy, w, None, path=E.NRT_PATH_BATCHED)
        assert eng.nrt_last_path() == want_path
        ws, wf, wr = pyoracle_nrt.nrt_batch(nd, pd, strategy, w, None, pitch=eng.Npad)
        assert np.array_equal(gr, wr) and np.array_equal(gf, wf) and np.array_equal(gs, ws)

    check(nodes, pods, 2, [1, 1, 1, 1], E.NRT_PATH_BATCHED)
    odd = dict(nodes, avail=nodes["avail"].copy())
    n4 = int(np.argmax(nodes["n_zones_node"] == 4))
    odd["avail"][1, 1, n4] += 1  # one zone's memory is off by a milli-byte: gcd 1, 2^40-ish values do not fit 32 bits
    check(odd, pods, 2, [1, 1, 1, 1], E.NRT_PATH_DIRECT)
    check(nodes, pods, 2, [1 << 40, 1, 1, 1], E.NRT_PATH_DIRECT)
    check(nodes, pods, 3, [1, 1, 1, 1], E.NRT_PATH_DIRECT)
    lone = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pods.items()}
    g = int(np.argmax((pods["qos"] == 0) & (pods["flags"] == 0)))
    lone["req_mask"][g, :] &= 1  # a Guaranteed pod whose containers name cpu only
    lone["req"][g, :, 1:] = 0
    check(nodes, lone, 1, None, E.NRT_PATH_DIRECT)
    check(nodes, lone, 2, None, E.NRT_PATH_BATCHED)
    # odd (non power-of-two) units still scale exactly: 1 000 000-byte blocks
    dec = dict(nodes, avail=nodes["avail"].copy())
    dec["avail"][:, 1] = (dec["avail"][:, 1] // (1 << 20) // 1000) * 1_000_000 * 1000
    pdec = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in pods.items()}
    pdec["req"][:, :, 1] = (pdec["req"][:, :, 1] // (1 << 20) // 1000) * 1_000_000 * 1000
    check(dec, pdec, 0, [2, 1, 1, 1], E.NRT_PATH_BATCHED)


def test_batched_after_patch_and_deduct(eng, engine_mod):
    """the gcd bookkeeping follows b200s_snapshot_patch_nrt / _nrt_deduct (rows with new units shrink the scale)."""
    from oracle import pyoracle_nrt

    E = engine_mod
    P, N, Z = 48, 900, 4
    na, pods = synth.gen_nrt(41, N, P, Z=Z)
    eng.config_nrt_path(E.NRT_PATH_BATCHED)
    eng.snapshot_begin(N)
    eng.snapshot_nrt(na)
    eng.snapshot_commit()
    eng.config_nrt(2, [1, 1, 1, 1])
    eng.pods_upload(P, nrt=pods)
    eng.eval(E.PLUGIN_NRT)
    assert eng.nrt_last_path() == E.NRT_PATH_BATCHED
    idx = np.array([3, 77, 500, 899], dtype=np.int32)
    target = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in na.items()}
    target["avail"][:, 0, idx] += 333          # cpu: odd milli values
    target["avail"][:, 1, idx] += 4096 * 1000  # memory: page-sized units instead of MiB
    target["node_flags"][idx[0]] ^= 8          # scope flips: the class lists are rebuilt
    rows = dict(target)
    for k in ("node_flags", "max_numa", "n_zones_node", "node_res_mask"):
        rows[k] = target[k][idx]
    rows["zone_res_mask"] = target["zone_res_mask"][:, idx]
    rows["avail"] = target["avail"][:, :, idx]
    rows["cost"] = target["cost"][:, :, idx]
    eng.snapshot_patch_begin()
    eng.snapshot_patch_nrt(idx, rows)
    ded_idx = np.array([10, 20], dtype=np.int32)
    deduct = np.zeros((4, 2), np.int64)
    deduct[0] = [1500, 250]
    deduct[1] = [3 * 4096 * 1000, 0]
    eng.snapshot_patch_nrt_deduct(ded_idx, np.array([3, 1], np.uint8), deduct)
    eng.snapshot_commit()
    for i, n in enumerate(ded_idx):
        for r in range(2):
            if (3, 1)[i] >> r & 1:
                a = target["avail"][:, r, n]
                listed = (target["zone_res_mask"][:, n] >> r) & 1
                target["avail"][:, r, n] = np.where(listed == 1, np.where(a < deduct[r, i], 0, a - deduct[r, i]), a)
    eng.eval(E.PLUGIN_NRT)
    assert eng.nrt_last_path() == E.NRT_PATH_BATCHED, eng.nrt_path_note()
    ws, wf, wr = pyoracle_nrt.nrt_batch(target, pods, 2, [1, 1, 1, 1], None, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), wr)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NRT), ws)


@pytest.mark.parametrize("strategy", [2, 1])
def test_config4_full_size(eng, engine_mod, strategy):
    """BASELINE config 4 at full size (5 000 pods x 20 000 nodes x 4 zones): sampled pods against the oracle, and the
    whole matrix of the batched path against the direct kernel (both dtypes carry the same 0..100 scores)."""
    E = engine_mod
    P, N = 5_000, 20_000
    seed = synth.BASE_SEED + 4
    nodes, pods = synth.gen_nrt(seed, N, P, Z=4)
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    w = [1, 1, 1, 1]
    eng.config_nrt_path(E.NRT_PATH_BATCHED)
    eng.snapshot_begin(N)
    eng.snapshot_nrt(nodes)
    eng.snapshot_commit()
    eng.config_nrt(strategy, w)
    eng.pods_upload(P, feasible=feas, nrt=pods)
    eng.eval(E.PLUGIN_NRT, E.OUT_U8)
    assert eng.nrt_last_path() == E.NRT_PATH_BATCHED
    bs = eng.fetch_scores(E.PLUGIN_NRT, E.OUT_U8)
    bf, br = eng.fetch_feasible(E.PLUGIN_NRT), eng.fetch_reasons(E.PLUGIN_NRT)
    rows = np.random.default_rng(4).choice(P, 48, replace=False)
    ws, wf, wr = oracle_rows(nodes, pods, rows, strategy, w, feas, eng.Npad)
    assert np.array_equal(br[rows], wr)
    assert np.array_equal(bf[rows], wf)
    assert np.array_equal(bs[rows].astype(np.int64), ws)
    eng.config_nrt_path(E.NRT_PATH_DIRECT)
    eng.eval(E.PLUGIN_NRT, E.OUT_U8)
    assert eng.nrt_last_path() == E.NRT_PATH_DIRECT
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), br)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NRT), bf)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NRT, E.OUT_U8), bs)
    # size-independent properties: scores in 0..100, infeasible -> 0, reasons only from the plugin's code set
    assert bs.max() <= 100 and set(np.unique(br)).issubset({0, 1, 2, 3, 4, 5, 8, 9})
    bits = np.unpackbits(bf.view(np.uint8), axis=1, bitorder="little")[:, :eng.Npad].astype(bool)
    assert not bs[~bits].any() and np.array_equal(bits[:, :N], br[:, :N] == 0) and not bits[:, N:].any()

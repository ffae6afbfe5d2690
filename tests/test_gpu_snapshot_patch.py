"""Incremental snapshot: rewriting a few node rows through b200s_snapshot_patch_* must give
exactly what a full re-flattening gives.  Each case uploads snapshot A in full, evaluates once (so every derived
array exists), patches a random subset of rows with snapshot B's rows (indices repeat: the last row wins),
and checks the CUDA result against the oracle run on the merged snapshot."""
import numpy as np
import pytest

from scheduler_plugins_b200 import synth
from test_gpu_parity import W_DEFAULT, netoh_setup, trimaran_snapshot

pytestmark = pytest.mark.gpu


def pick_rows(seed, N, frac=0.07):
    """Patch list with repeats; returns (node_idx, src) where src[j] = which copy of B's row travels in slot j
    (a repeated index first carries a decoy row, then the real one)."""
    g = np.random.default_rng(seed)
    uniq = g.choice(N, size=max(1, int(N * frac)), replace=False).astype(np.int32)
    dup = uniq[: max(1, len(uniq) // 5)]
    idx = np.concatenate([dup, uniq])              # `dup` rows appear twice: first occurrence must lose
    decoy = np.concatenate([np.ones(len(dup), bool), np.zeros(len(uniq), bool)])
    return idx, decoy, uniq


def merged(a, b, uniq):
    out = np.array(a, copy=True)
    out[..., uniq] = np.asarray(b)[..., uniq]
    return out


def rows(b, idx, decoy, fill=0):
    """Rows of B in patch order; decoy slots get a wrong value that must be overwritten."""
    r = np.array(np.asarray(b)[..., idx], copy=True)
    r[..., decoy] = fill
    return r


@pytest.mark.parametrize("P,N,mode", [(64, 1000, 0), (33, 4097, 1), (4, 130, 1)])
def test_patch_allocatable(eng, engine_mod, oracle, P, N, mode):
    E = engine_mod
    seed = synth.BASE_SEED + 61
    a, b = synth.gen_nodes(seed, N), synth.gen_nodes(seed + 1, N)
    cols_a = [a["alloc_cpu_milli"], a["alloc_mem_bytes"]]
    cols_b = [b["alloc_cpu_milli"], b["alloc_mem_bytes"]]
    eng.snapshot_begin(N)
    eng.snapshot_allocatable(cols_a)
    eng.snapshot_commit()
    eng.config_allocatable(mode, W_DEFAULT)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad)
    eng.pods_upload(P, feasible=feas)
    eng.eval(E.PLUGIN_ALLOCATABLE)  # builds the sorted raw scores of A
    idx, decoy, uniq = pick_rows(seed, N)
    eng.snapshot_patch_begin(2)
    eng.snapshot_patch_allocatable(idx, [rows(c, idx, decoy, fill=1) for c in cols_b])
    eng.snapshot_commit()
    eng.eval(E.PLUGIN_ALLOCATABLE)  # the pod batch survives a patch
    got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)
    cols_t = [merged(ca, cb, uniq) for ca, cb in zip(cols_a, cols_b)]
    want = oracle.alloc_batch(cols_t, W_DEFAULT, mode, P, feas, pitch=eng.Npad)
    assert np.array_equal(got, want)
    assert not np.array_equal(want, oracle.alloc_batch(cols_a, W_DEFAULT, mode, P, feas, pitch=eng.Npad))


def test_patch_trimaran(eng, engine_mod, oracle):
    E = engine_mod
    P, N, seed = 48, 2049, synth.BASE_SEED + 62
    na, nb = synth.gen_nodes(seed, N), synth.gen_nodes(seed + 1, N)
    ta, tb = synth.gen_trimaran(seed, na), synth.gen_trimaran(seed + 1, nb)
    pods = synth.gen_pods(seed, P)
    trimaran_snapshot(eng, na, ta)
    eng.config_tlp(40)
    eng.config_lvrb(1.0, 1.0)
    eng.pods_upload(P, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=pods["req_cpu_milli"],
                    lvrb_req_mem_bytes=pods["req_mem_bytes"])
    idx, decoy, uniq = pick_rows(seed, N)
    eng.snapshot_patch_begin(2)
    eng.snapshot_patch_tlp(idx, rows(tb["cpu_avg"], idx, decoy), rows(nb["cap_cpu_milli"], idx, decoy),
                           rows(tb["missing_milli"], idx, decoy), rows(tb["tlp_flags"], idx, decoy))
    eng.snapshot_patch_lvrb(idx, *[rows(tb[k], idx, decoy) for k in ("cpu_avg", "cpu_std", "mem_avg", "mem_std")],
                            rows(nb["alloc_cpu_milli"], idx, decoy), rows(nb["alloc_mem_bytes"], idx, decoy),
                            rows(tb["lvrb_flags"], idx, decoy))
    eng.snapshot_commit()
    m = lambda x, y: merged(x, y, uniq)  # noqa: E731
    eng.eval(E.PLUGIN_TLP)
    want = oracle.tlp_batch(m(ta["cpu_avg"], tb["cpu_avg"]), m(na["cap_cpu_milli"], nb["cap_cpu_milli"]),
                            m(ta["missing_milli"], tb["missing_milli"]), m(ta["tlp_flags"], tb["tlp_flags"]),
                            pods["tlp_pod_cpu_milli"], 40, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_TLP), want)
    eng.eval(E.PLUGIN_LVRB)
    want = oracle.lvrb_batch(*[m(ta[k], tb[k]) for k in ("cpu_avg", "cpu_std", "mem_avg", "mem_std")],
                             m(na["alloc_cpu_milli"], nb["alloc_cpu_milli"]),
                             m(na["alloc_mem_bytes"], nb["alloc_mem_bytes"]), m(ta["lvrb_flags"], tb["lvrb_flags"]),
                             pods["req_cpu_milli"], pods["req_mem_bytes"], 1.0, 1.0, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_LVRB), want)


@pytest.mark.parametrize("strategy", [2, 3])
def test_patch_nrt(eng, engine_mod, strategy):
    from oracle import pyoracle_nrt

    E = engine_mod
    P, N, Z, seed = 40, 1500, 4, synth.BASE_SEED + 63
    na, pods = synth.gen_nrt(seed, N, P, Z=Z)
    nb, _ = synth.gen_nrt(seed + 1, N, P, Z=Z)
    feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    eng.snapshot_begin(N)
    eng.snapshot_nrt(na)
    eng.snapshot_commit()
    eng.config_nrt(strategy, [3, 1, 2, 1])
    eng.pods_upload(P, feasible=feas, nrt=pods)
    eng.eval(E.PLUGIN_NRT)
    idx, decoy, uniq = pick_rows(seed, N)
    keys = ("node_flags", "max_numa", "n_zones_node", "node_res_mask", "zone_res_mask", "avail", "cost")
    patch = {k: rows(nb[k], idx, decoy) for k in keys}
    patch.update(n_zones=na["n_zones"], n_res=na["n_res"])
    eng.snapshot_patch_begin(2)
    eng.snapshot_patch_nrt(idx, patch)
    eng.snapshot_commit()
    eng.eval(E.PLUGIN_NRT)
    target = dict(na)
    for k in keys:
        target[k] = merged(na[k], nb[k], uniq)
    ws, wf, wr = pyoracle_nrt.nrt_batch(target, pods, strategy, [3, 1, 2, 1], feas, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), wr)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NRT), wf)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NRT), ws)
    # the patch moved nodes between control-flow classes (scope / flags differ between the two seeds)
    assert (na["node_flags"][uniq] != nb["node_flags"][uniq]).any()


@pytest.mark.parametrize("want_counts", [False, True])
def test_patch_network_overhead(eng, engine_mod, oracle, want_counts):
    E = engine_mod
    P, N, seed = 64, 3001, synth.BASE_SEED + 64
    net = synth.gen_netoh(seed, N, P)
    netoh_setup(eng, E, net, N)
    eng.config_network_overhead(want_counts=want_counts)
    feas = synth.gen_feasible_words(seed, P, N, eng.Npad)
    eng.pods_upload(P, feasible=feas, netoh=net)
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
    # move nodes to other (region, zone) labels; one of them to a pair no node had before
    g = np.random.default_rng(seed)
    idx, decoy, uniq = pick_rows(seed, N)
    region_b = np.array(net["region_all"][:N], copy=True)
    zone_b = np.array(net["zone_all"][:N], copy=True)
    perm = g.permutation(N)
    region_b[uniq], zone_b[uniq] = net["region_all"][perm[uniq]], net["zone_all"][perm[uniq]]
    region_b[uniq[0]], zone_b[uniq[0]] = net["region_all"][perm[0]], 0  # region label without a zone label
    eng.snapshot_patch_begin(2)
    eng.snapshot_patch_network_overhead(idx, rows(region_b, idx, decoy), rows(zone_b, idx, decoy))
    eng.snapshot_commit()
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
    ws, wf, wr = oracle.netoh_batch(net["zone_cost"], net["region_cost"], region_b, zone_b, net["score_equally"],
                                    net["dep_offset"], net["deps"], feas, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NETWORK_OVERHEAD), wf)
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NETWORK_OVERHEAD), wr)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD), ws)


def test_patch_state_errors(eng, engine_mod):
    E = engine_mod
    nodes = synth.gen_nodes(synth.BASE_SEED, 300)
    with pytest.raises(E.B200SError):  # nothing committed yet on a fresh snapshot
        eng.snapshot_begin(300)
        eng.snapshot_patch_begin(1)
    eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
    eng.snapshot_commit()
    one = np.zeros(1, np.int64)
    with pytest.raises(E.B200SError):  # patch call outside a patch session
        eng.snapshot_patch_allocatable([0], [one, one])
    eng.snapshot_patch_begin(2)
    with pytest.raises(E.B200SError):  # index out of range
        eng.snapshot_patch_allocatable([300], [one, one])
    with pytest.raises(E.B200SError):  # resource count differs
        eng.snapshot_patch_allocatable([0], [one])
    with pytest.raises(E.B200SError):  # TLP columns were never uploaded
        eng.snapshot_patch_tlp([0], np.zeros(1), one, one, np.zeros(1, np.uint8))
    with pytest.raises(E.B200SError):  # no eval while the snapshot is open
        eng.pods_upload(1)
        eng.eval(E.PLUGIN_ALLOCATABLE)
    eng.snapshot_patch_allocatable(np.zeros(0, np.int32), [np.zeros(0, np.int64)] * 2)  # empty patch is a no-op
    eng.snapshot_commit()
    eng.config_allocatable(0, W_DEFAULT)
    eng.pods_upload(1)
    eng.eval(E.PLUGIN_ALLOCATABLE)


def test_patch_trimaran2(eng, engine_mod, oracle):
    """Peaks / LowRiskOverCommitment rows (a bind changes a node's request and limit sums): patched == re-flattened.
    LowRisk compares two CUDA results (full vs patched), so it is exact despite the libm/CUDA tolerance."""
    E = engine_mod
    P, N, seed = 24, 1500, synth.BASE_SEED + 65
    na, nb = synth.gen_nodes(seed, N), synth.gen_nodes(seed + 1, N)
    ta, tb = synth.gen_trimaran(seed, na), synth.gen_trimaran(seed + 1, nb)
    a2, b2 = synth.gen_trimaran2(seed, na, P), synth.gen_trimaran2(seed + 1, nb, P)
    idx, decoy, uniq = pick_rows(seed, N)
    m = lambda x, y: merged(x, y, uniq)  # noqa: E731
    peaks_cols = lambda n, t, t2: (t["cpu_avg"], n["cap_cpu_milli"], t["tlp_flags"], t2["k1"], t2["k2"])  # noqa: E731
    lr_cols = lambda n, t, t2: (t["cpu_avg"], t["cpu_std"], t["mem_avg"], t["mem_std"], n["alloc_cpu_milli"],  # noqa: E731
                                n["alloc_mem_bytes"], t["lvrb_flags"], t2["node_req_cpu"], t2["node_req_mem"],
                                t2["node_lim_cpu"], t2["node_lim_mem"])
    pa, pb, la, lb = peaks_cols(na, ta, a2), peaks_cols(nb, tb, b2), lr_cols(na, ta, a2), lr_cols(nb, tb, b2)
    pods = dict(peaks_pod_cpu_milli=a2["peaks_pod_cpu_milli"], low_risk_pod=a2["low_risk_pod"])

    def evaluate():
        eng.pods_upload(P, **pods)
        eng.eval(E.PLUGIN_PEAKS)
        eng.eval(E.PLUGIN_LOW_RISK)
        return eng.fetch_scores(E.PLUGIN_PEAKS), eng.fetch_scores(E.PLUGIN_LOW_RISK)

    eng.config_low_risk(5, 0.5, 0.5)
    eng.snapshot_begin(N)  # reference run: the merged snapshot uploaded in full
    eng.snapshot_peaks(*[m(x, y) for x, y in zip(pa, pb)])
    eng.snapshot_low_risk(*[m(x, y) for x, y in zip(la, lb)])
    eng.snapshot_commit()
    want_peaks, want_lr = evaluate()
    eng.snapshot_begin(N)  # A in full, evaluated (derives the risk columns of A), then patched with B's rows
    eng.snapshot_peaks(*pa)
    eng.snapshot_low_risk(*la)
    eng.snapshot_commit()
    before_peaks, before_lr = evaluate()
    eng.snapshot_patch_begin(3)
    eng.snapshot_patch_peaks(idx, *[rows(x, idx, decoy) for x in pb])
    eng.snapshot_patch_low_risk(idx, *[rows(x, idx, decoy) for x in lb])
    eng.snapshot_commit()
    got_peaks, got_lr = evaluate()
    assert np.array_equal(got_peaks, want_peaks) and np.array_equal(got_lr, want_lr)
    assert not np.array_equal(before_peaks, want_peaks) and not np.array_equal(before_lr, want_lr)
    assert np.array_equal(got_peaks, oracle.peaks_batch(*[m(x, y) for x, y in zip(pa, pb)], a2["peaks_pod_cpu_milli"],
                                                        None, pitch=eng.Npad))


def test_patch_nrt_overreserve_deduct(eng, engine_mod):
    """OverReserve cache (cache/store.go:129-160): assumed pods are taken off every zone of their node on the device;
    Filter + Score afterwards equal the oracle run on the NRT the reference's GetCachedNRTCopy would hand out."""
    import ctypes as C

    from oracle import pyoracle as orc
    from oracle import pyoracle_nrt

    E = engine_mod
    P, N, Z, seed = 32, 900, 4, synth.BASE_SEED + 66
    nodes, pods = synth.gen_nrt(seed, N, P, Z=Z)
    R = nodes["n_res"]
    eng.snapshot_begin(N)
    eng.snapshot_nrt(nodes)
    eng.snapshot_commit()
    eng.config_nrt(2, [1, 1, 1, 1])
    eng.pods_upload(P, nrt=pods)
    eng.eval(E.PLUGIN_NRT)
    before = eng.fetch_reasons(E.PLUGIN_NRT)
    g = np.random.default_rng(seed)
    idx = np.sort(g.choice(N, size=120, replace=False)).astype(np.int32)
    # 1..3 assumed pods per node; summed per resource for the engine, applied pod by pod for the oracle
    target = dict(nodes)
    target["avail"] = np.array(nodes["avail"], copy=True)
    res_mask = np.zeros(len(idx), np.uint8)
    deduct = np.zeros((R, len(idx)), np.int64)
    for j, n in enumerate(idx):
        a = np.ascontiguousarray(target["avail"][:, :, n])
        zm = np.ascontiguousarray(nodes["zone_res_mask"][:, n])
        for _ in range(int(g.integers(1, 4))):
            m = int(g.integers(1, 16))
            q = np.array([g.integers(0, 9) * 1000, g.integers(0, 17) << 30, g.integers(0, 3) << 29, g.integers(0, 5) * 1000],
                         dtype=np.int64) * np.array([1, 1000, 1000, 1])
            orc.lib().orc_nrt_overreserve_deduct(C.c_void_p(a.ctypes.data), C.c_void_p(zm.ctypes.data), C.c_int(Z),
                                                 C.c_int(R), C.c_uint8(m), C.c_void_p(q.ctypes.data))
            res_mask[j] |= m
            for r in range(R):
                if (m >> r) & 1:
                    deduct[r, j] += q[r]
        target["avail"][:, :, n] = a
    eng.snapshot_patch_begin(2)
    with pytest.raises(E.B200SError):  # a repeated node would race
        eng.snapshot_patch_nrt_deduct([5, 5], [1, 1], np.zeros((R, 2), np.int64))
    eng.snapshot_patch_nrt_deduct(idx, res_mask, deduct)
    eng.snapshot_commit()
    eng.eval(E.PLUGIN_NRT)
    ws, wf, wr = pyoracle_nrt.nrt_batch(target, pods, 2, [1, 1, 1, 1], None, pitch=eng.Npad)
    assert np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT), wr)
    assert np.array_equal(eng.fetch_feasible(E.PLUGIN_NRT), wf)
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_NRT), ws)
    assert (wr != before).any()  # the deduction turned fits into rejects

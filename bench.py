#!/usr/bin/env python
"""bench.py — the measurement contract of the hot path.

A "step" is ONE pass of the hot path over one batch of synthetic input.  `--config` picks the BASELINE.json
workload (default c2, the configuration the headline metric is quoted on):

  c2  NodeResourcesAllocatable, mode Most + NormalizeScore, 10k pods x 50k nodes PER GPU (weak scaling; one per-pod
      min/max all-reduce per step across the shards)
  c3  TargetLoadPacking + LoadVariationRiskBalancing, 10k pods x 50k nodes PER GPU (weak scaling, no collective)
  c4  NodeResourceTopologyMatch Filter + Score (LeastAllocated), 5k pods x 20k nodes x 4 NUMA zones, the 20k nodes
      SPLIT over the GPUs (strong scaling, no collective)
  c5  combined profile: all five plugins, weighted sum, per-pod top-1; 50k pods x 200k nodes, the nodes SPLIT over
      the GPUs (strong scaling): per step and pod chunk one min/max all-reduce per normalising plugin and ONE
      ncclAllGather of the per-pod winners + fold.  The 50k pods go through the engine in chunks of --chunk pods.

  value     pod x node Score evals/s with inputs resident in HBM, CUDA events on the engine's stream, max over ranks
            (c3 / c5 count one eval per plugin per pair; `config.pairs_per_s` is the plugin-independent figure)
  roofline  the dominant kernel of the config — algorithmic bytes / its mean launch time (events inside the timed
            region) against the measured HBM copy peak of MEASURED_PEAKS.json
  e2e       the same metric through the C-ABI call the Go shim makes, HOST buffers in and out, copies inside the timed
            region (c2/c3: b200s_score_batch, u8 matrix back; c4: + feasibility words and reason codes; c5: TOPK mode —
            pod columns in, [P] winners out)
  parity    after the timed region every rank compares sampled pods of ITS shard (and, c5, the folded global top-k)
            with the UNSHARDED oracle; `parity_checked` = number of ranks that did
  cpu_baseline  the CPU oracle (a port of the Go path; Go is not installed) on a bounded sample, 1 thread (N = 1 only)
  cycle_latency (c2, N = 1) BASELINE metric 2: P = 1 wall-clock latency of one cycle through the C-ABI, and the cost
            of refreshing the snapshot between two cycles (16-row patch vs full re-upload)

`--impl reference` times the reference's CPU path alone (rank 0 only; never loads the CUDA library): the Go-faithful
restatement (c2) / the C ports of the oracle (c3..c5) on all host threads, each step a bounded sample of the config.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODE_MOST = 1
WEIGHTS = [1 << 20, 1]
METRIC = "pod_x_node_score_evals_per_sec"
UNIT = "evals/s"
BASE_SEED = 0xB2005EED
PROFILE_WEIGHTS = [1, 1, 1, 1, 5, 0, 0]  # NetworkOverhead weight 5 as in manifests/networktopology/scheduler-config.yaml

CONFIGS = {
    "c2": dict(no=2, pods=10_000, nodes=50_000, scaling="weak", dtype="int64",
               label="configs[1]: NodeResourcesAllocatable Most + NormalizeScore, 10k pods x 50k nodes/GPU"),
    "c3": dict(no=3, pods=10_000, nodes=50_000, scaling="weak", dtype="f64",
               label="configs[2]: TargetLoadPacking + LoadVariationRiskBalancing, 10k pods x 50k nodes/GPU"),
    "c4": dict(no=4, pods=5_000, nodes=20_000, scaling="strong", dtype="int64",
               label="configs[3]: NodeResourceTopologyMatch single-numa-node Filter + Score, 5k pods x 20k nodes x 4 zones"),
    "c5": dict(no=5, pods=50_000, nodes=200_000, scaling="strong", dtype="int64",
               label="configs[4]: combined profile (Allocatable + TLP + LVRB + NRT + NetworkOverhead), 50k pods x 200k nodes"),
}


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def npad_of(n):
    return (max(n, 1) + 127) // 128 * 128


def shard_bounds(n_nodes, world):
    """contiguous, 128-aligned starts, sizes differ by at most one block (scheduler-plugins_b200/sharding.py)"""
    blocks = (n_nodes + 127) // 128
    base, extra = divmod(blocks, world)
    out, off = [], 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        cnt = max(0, min(nb * 128, n_nodes - off))
        out.append((off, cnt))
        off += cnt
    return out


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------------
# synthetic inputs: the GLOBAL snapshot of a config (every rank generates it from the seed and slices its shard;
# the oracle check after the timed region needs the unsharded columns anyway)
def gen_inputs(synth, cfg, seed, P, N_global):
    d = dict(nodes=synth.gen_nodes(seed, N_global), pods=synth.gen_pods(seed, P))
    if cfg in ("c3", "c5"):
        d["tri"] = synth.gen_trimaran(seed, d["nodes"])
    if cfg in ("c4", "c5"):
        d["nrt_nodes"], d["nrt_pods"] = synth.gen_nrt(seed, N_global, P, Z=4)
    if cfg == "c5":
        d["net"] = synth.gen_netoh(seed, N_global, P)
    return d


def slice_nrt_nodes(nn, sl):
    return dict(nn, node_flags=nn["node_flags"][sl], max_numa=nn["max_numa"][sl], n_zones_node=nn["n_zones_node"][sl],
                node_res_mask=nn["node_res_mask"][sl], zone_res_mask=np.ascontiguousarray(nn["zone_res_mask"][:, sl]),
                avail=np.ascontiguousarray(nn["avail"][:, :, sl]),
                cost=None if nn.get("cost") is None else np.ascontiguousarray(nn["cost"][:, :, sl]))


def slice_pods(d, rows):
    """pod-side columns of a config for a subset of pods (oracle samples / pod chunks)"""
    out = dict(pods={k: (v[rows] if isinstance(v, np.ndarray) else v) for k, v in d["pods"].items()})
    if "nrt_pods" in d:
        out["nrt_pods"] = {k: (v[rows] if isinstance(v, np.ndarray) else v) for k, v in d["nrt_pods"].items()}
    if "net" in d:
        net = d["net"]
        offs, deps, cur = [0], [], 0
        for p in rows:
            a, b = int(net["dep_offset"][p]), int(net["dep_offset"][p + 1])
            deps.append(net["deps"][a:b])
            cur += b - a
            offs.append(cur)
        out["net"] = dict(net, score_equally=net["score_equally"][rows], dep_offset=np.array(offs, dtype=np.int32),
                          deps=np.concatenate(deps) if cur else net["deps"][:0])
    return out


def upload_snapshot(E, eng, cfg, d, off, cnt, n_global):
    sl = slice(off, off + cnt)
    nodes = d["nodes"]
    eng.snapshot_begin(cnt, generation=1, node_offset=off, n_nodes_global=n_global)
    if cfg in ("c2", "c5"):
        eng.snapshot_allocatable([nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl]])
    if cfg in ("c3", "c5"):
        tri = d["tri"]
        eng.snapshot_tlp(tri["cpu_avg"][sl], nodes["cap_cpu_milli"][sl], tri["missing_milli"][sl], tri["tlp_flags"][sl])
        eng.snapshot_lvrb(tri["cpu_avg"][sl], tri["cpu_std"][sl], tri["mem_avg"][sl], tri["mem_std"][sl],
                          nodes["alloc_cpu_milli"][sl], nodes["alloc_mem_bytes"][sl], tri["lvrb_flags"][sl])
    if cfg in ("c4", "c5"):
        eng.snapshot_nrt(slice_nrt_nodes(d["nrt_nodes"], sl))
    if cfg == "c5":
        net = d["net"]
        eng.snapshot_network_overhead(net["region_all"][sl], net["zone_all"][sl], net["zone_cost"], net["region_cost"])
    eng.snapshot_commit()
    if cfg in ("c2", "c5"):
        eng.config_allocatable(MODE_MOST, WEIGHTS)
    if cfg in ("c3", "c5"):
        eng.config_tlp(40)
        eng.config_lvrb(1.0, 1.0)
    if cfg in ("c4", "c5"):
        eng.config_nrt(E.NRT_LEAST_ALLOCATED, [1, 1, 1, 1])


def pod_columns(cfg, dp, feas=None):
    """keyword arguments of Engine.pods_upload / make_batch for the pods in dp (slice_pods output)"""
    kw = {}
    if feas is not None:
        kw["feasible"] = feas
    if cfg in ("c3", "c5"):
        kw.update(tlp_pod_cpu_milli=dp["pods"]["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=dp["pods"]["req_cpu_milli"],
                  lvrb_req_mem_bytes=dp["pods"]["req_mem_bytes"])
    if cfg in ("c4", "c5"):
        kw["nrt"] = dp["nrt_pods"]
    if cfg == "c5":
        kw["netoh"] = dp["net"]
    return kw


# ------------------------------------------------------------------------------------------------------------------
# oracle: the unsharded CPU restatement on a few sampled pods (checker and CPU baseline; never the product path)
def oracle_rows(cfg, d, rows, feas_global, N_global, threads=1):
    """per-plugin score rows [len(rows)][npad(N_global)] of the sampled pods over ALL nodes + (c5) the top-1"""
    from oracle import pyoracle as orc

    dp = slice_pods(d, rows)
    nodes = d["nodes"]
    pitch = npad_of(N_global)
    fw = None if feas_global is None else np.ascontiguousarray(feas_global[rows])
    out = {}
    if cfg == "c2":
        out["alloc"] = orc.alloc_batch([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]], WEIGHTS, MODE_MOST, len(rows), fw,
                                       pitch=pitch)
    if cfg == "c3":
        tri = d["tri"]
        out["tlp"] = orc.tlp_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"],
                                   dp["pods"]["tlp_pod_cpu_milli"], 40, pitch=pitch)
        out["lvrb"] = orc.lvrb_batch(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                                     nodes["alloc_mem_bytes"], tri["lvrb_flags"], dp["pods"]["req_cpu_milli"],
                                     dp["pods"]["req_mem_bytes"], 1.0, 1.0, pitch=pitch)
    if cfg == "c4":
        from oracle import pyoracle_nrt

        s, f, r = pyoracle_nrt.nrt_batch(d["nrt_nodes"], dp["nrt_pods"], 2, [1, 1, 1, 1], fw, pitch=pitch)
        out.update(nrt=s, nrt_feas=f, nrt_reasons=r)
    if cfg == "c5":
        from oracle import combined as OC

        tri, net = d["tri"], dp["net"]
        total, feas, topk = OC.combined(
            len(rows), N_global, pitch, fw, PROFILE_WEIGHTS[:5], 1,
            alloc=dict(cols=[nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]], weights=WEIGHTS, mode=MODE_MOST),
            tlp=dict(util=tri["cpu_avg"], cap=nodes["cap_cpu_milli"], missing=tri["missing_milli"], flags=tri["tlp_flags"],
                     pod_cpu=dp["pods"]["tlp_pod_cpu_milli"], target=40),
            lvrb=dict(node_cols=[tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                                 nodes["alloc_mem_bytes"], tri["lvrb_flags"]], req_cpu=dp["pods"]["req_cpu_milli"],
                      req_mem=dp["pods"]["req_mem_bytes"], margin=1.0, sens=1.0),
            nrt=dict(nodes=d["nrt_nodes"], pods=dp["nrt_pods"], strategy=2, weights=[1, 1, 1, 1]),
            netoh=dict(zone_cost=net["zone_cost"], region_cost=net["region_cost"], region_id=net["region_all"],
                       zone_id=net["zone_all"], score_equally=net["score_equally"], dep_offset=net["dep_offset"],
                       deps=net["deps"]))
        out.update(total=total, total_feas=feas, topk=topk)
    return out


def cpu_threads_sample(cfg, d, N_global, sample_pods, threads, feas_global):
    """Times the oracle port on sample_pods x N_global with `threads` host threads (pods split across threads; the C
    functions release the GIL).  Returns (evals/s counted like `value`, seconds)."""
    from concurrent.futures import ThreadPoolExecutor

    chunks = [c for c in np.array_split(np.arange(sample_pods), max(1, threads)) if len(c)]
    t0 = time.perf_counter()
    if len(chunks) == 1:
        oracle_rows(cfg, d, chunks[0], feas_global, N_global)
    else:
        with ThreadPoolExecutor(len(chunks)) as ex:
            list(ex.map(lambda c: oracle_rows(cfg, d, c, feas_global, N_global), chunks))
    dt = time.perf_counter() - t0
    return sample_pods * N_global * plugins_per_pair(cfg) / dt, dt


def plugins_per_pair(cfg):
    return {"c2": 1, "c3": 2, "c4": 1, "c5": 5}[cfg]


def cpu_gofaithful(orc, cols, feas, pods, sample_pods, threads):
    """Times the Go-faithful restatement (oracle/gofaithful.cpp: per-call maps, string switches, NodeScoreList —
    the structure the Go scheduler actually executes) on sample_pods x N, `threads` cycles in parallel."""
    N = len(cols[0])
    _, dt = orc.gofaithful_alloc_batch(cols, ["cpu", "memory"], WEIGHTS, MODE_MOST, pods["req_cpu_milli"][:sample_pods],
                                       pods["req_mem_bytes"][:sample_pods], np.ascontiguousarray(feas[:sample_pods]),
                                       pitch=N, threads=threads, return_seconds=True)  # cycles only, snapshot pre-built
    return sample_pods * N / dt, dt


# ------------------------------------------------------------------------------------------------------------------
def cycle_latency(E, synth, device, N, cycles=1000):
    """Scheduling-cycle latency (BASELINE.json metric 2): P = 1 pod, snapshot already resident; per cycle the
    pod columns go host->device, every enabled plugin is evaluated over all N nodes, and the result comes back.
    Wall clock around the C-ABI calls (that is what the scheduler goroutine waits for)."""
    seed = BASE_SEED + 5
    nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, 1)
    tri = synth.gen_trimaran(seed, nodes)
    nn, npods = synth.gen_nrt(seed, N, 1, Z=4)
    net = synth.gen_netoh(seed, N, 1)
    net["score_equally"][:] = 0
    if net["dep_offset"][1] == 0:  # make sure the one pod has dependencies
        net = synth.gen_netoh(seed + 1, N, 1)
        net["score_equally"][:] = 0
    eng = E.Engine(device)
    eng.snapshot_begin(N)
    eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
    eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
    eng.snapshot_lvrb(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                      nodes["alloc_mem_bytes"], tri["lvrb_flags"])
    eng.snapshot_nrt(nn)
    eng.snapshot_network_overhead(net["region_all"], net["zone_all"], net["zone_cost"], net["region_cost"])
    eng.snapshot_commit()
    eng.config_allocatable(MODE_MOST, WEIGHTS)
    eng.config_tlp(40)
    eng.config_lvrb(1.0, 1.0)
    eng.config_nrt(E.NRT_LEAST_ALLOCATED, [1, 1, 1, 1])
    cols = dict(tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=pods["req_cpu_milli"],
                lvrb_req_mem_bytes=pods["req_mem_bytes"], nrt=npods, netoh=net)
    batch, keep = eng.make_batch(1, **cols)
    out = eng.pinned(eng.Npad)
    row = out.view(np.uint8, (1, eng.Npad))

    def p50(fn):
        for _ in range(20):
            fn()
        ts = []
        for _ in range(cycles):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        return {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99))}

    res = {"nodes": N, "pods_per_cycle": 1, "cycles": cycles}
    res["NodeResourcesAllocatable"] = p50(lambda: eng.score_batch(E.PLUGIN_ALLOCATABLE, batch, E.OUT_U8, row))
    w = PROFILE_WEIGHTS

    topk1 = np.empty((1, 1), dtype=E.TOPK_DTYPE)

    # ONE C-ABI call: pod columns in, winner out (b200s_schedule_batch); ctypes arguments marshalled once, as cgo's are
    combined = eng.prepare_schedule_batch(batch, 0b11111, w, 1, topk1)

    res["all_five_plugins_top1"] = p50(combined)
    res["all_five_plugins_top1"]["path"] = ("b200s_schedule_batch -> cycle.cu: copy of the pod columns + two kernels as one "
                                            "graph launch, winner written to a mapped host page, no per-plugin matrix")
    fused_winner = (int(topk1[0, 0]["score"]), int(topk1[0, 0]["node"]))
    eng.config_fused_cycle(2)  # the same two kernels issued as plain launches + copies
    res["all_five_plugins_top1_plain_launches"] = p50(combined)
    assert (int(topk1[0, 0]["score"]), int(topk1[0, 0]["node"])) == fused_winner, "graph and plain-launch cycle disagree"
    eng.config_fused_cycle(False)  # the plugin-by-plugin path of round 1 (13 launches) for comparison
    res["all_five_plugins_top1_plugin_by_plugin"] = p50(combined)
    assert (int(topk1[0, 0]["score"]), int(topk1[0, 0]["node"])) == fused_winner, "fused cycle and plugin-by-plugin path disagree"
    eng.config_fused_cycle(True)

    # snapshot refresh between two cycles: a node event changes a few NodeInfos.  Either the host
    # re-uploads every column of all five plugins, or it patches the 16 rows that changed; both leave the engine
    # ready for the next cycle (the patch re-sorts Allocatable's raw scores at the next eval, so one eval is inside).
    g = np.random.default_rng(seed)
    idx = np.sort(g.choice(N, size=16, replace=False)).astype(np.int32)
    nrt_rows = {k: np.ascontiguousarray(np.asarray(nn[k])[..., idx])
                for k in ("node_flags", "max_numa", "n_zones_node", "node_res_mask", "zone_res_mask", "avail", "cost")}
    nrt_rows.update(n_zones=nn["n_zones"], n_res=nn["n_res"])
    gen = [2]

    def refresh_patch():
        gen[0] += 1
        eng.snapshot_patch_begin(gen[0])
        eng.snapshot_patch_allocatable(idx, [nodes["alloc_cpu_milli"][idx], nodes["alloc_mem_bytes"][idx]])
        eng.snapshot_patch_tlp(idx, tri["cpu_avg"][idx], nodes["cap_cpu_milli"][idx], tri["missing_milli"][idx],
                               tri["tlp_flags"][idx])
        eng.snapshot_patch_lvrb(idx, tri["cpu_avg"][idx], tri["cpu_std"][idx], tri["mem_avg"][idx], tri["mem_std"][idx],
                                nodes["alloc_cpu_milli"][idx], nodes["alloc_mem_bytes"][idx], tri["lvrb_flags"][idx])
        eng.snapshot_patch_nrt(idx, nrt_rows)
        eng.snapshot_patch_network_overhead(idx, net["region_all"][idx], net["zone_all"][idx])
        eng.snapshot_commit()
        eng.score_batch(E.PLUGIN_ALLOCATABLE, batch, E.OUT_U8, row)

    def refresh_full():
        gen[0] += 1
        eng.snapshot_begin(N, generation=gen[0])
        eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
        eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
        eng.snapshot_lvrb(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                          nodes["alloc_mem_bytes"], tri["lvrb_flags"])
        eng.snapshot_nrt(nn)
        eng.snapshot_network_overhead(net["region_all"], net["zone_all"], net["zone_cost"], net["region_cost"])
        eng.snapshot_commit()
        eng.score_batch(E.PLUGIN_ALLOCATABLE, batch, E.OUT_U8, row)

    full_cycles, cycles = cycles, max(20, cycles // 10)
    res["snapshot_refresh_16_rows_patch_plus_cycle"] = p50(refresh_patch)
    res["snapshot_refresh_full_upload_plus_cycle"] = p50(refresh_full)
    cycles = full_cycles
    out.free()
    eng.close()
    # the same cycle on the CPU through the Go-faithful restatement of NodeResourcesAllocatable in UPSTREAM's shape:
    # one pod at a time, Score fanned out over a 16-worker Parallelizer, NormalizeScore serial
    try:
        from oracle import pyoracle as orc

        cols = [nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]]
        feas = synth.gen_feasible_words(seed, 48, N, npad_of(N))
        p48 = synth.gen_pods(seed, 48)
        _, secs = orc.gofaithful_alloc_cycles16(cols, ["cpu", "memory"], WEIGHTS, MODE_MOST, p48["req_cpu_milli"],
                                                p48["req_mem_bytes"], feas, pitch=N, workers=16)
        _, dt1 = cpu_gofaithful(orc, cols, feas, p48, 1, 1)
        res["cpu_gofaithful_NodeResourcesAllocatable"] = {
            "parallelizer_16_workers_p50_us": float(np.percentile(secs[8:], 50) * 1e6),
            "one_cycle_1_thread_us": dt1 * 1e6, "host_cores": os.cpu_count(),
            "note": "pods one at a time, chunkSizeFor(n, 16) nodes per work piece (upstream Parallelizer shape)"}
    except Exception as e:  # the oracle is only the checker; its absence must not break the bench
        res["cpu_gofaithful_NodeResourcesAllocatable"] = {"unavailable": str(e)}
    return res


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path alone (rank 0).  Builds and loads the ORACLE only -- the CUDA
    library is never loaded by this process."""
    if rank != 0:
        return
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        print(json.dumps({"impl": "reference", "unavailable": "oracle build failed: " + r.stderr[-200:]}), flush=True)
        return
    from oracle import pyoracle as orc
    from scheduler_plugins_b200 import synth  # pure numpy generator; importing it does not load libb200sched.so

    cfg = args.config
    spec = CONFIGS[cfg]
    seed = BASE_SEED + spec["no"]
    N = args.nodes or spec["nodes"]
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if cfg == "c2":
        sample = max(threads * 24, 256)
        nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, sample)
        feas = synth.gen_feasible_words(seed, sample, N, npad_of(N))
        cols = [nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]]
        run = lambda: cpu_gofaithful(orc, cols, feas, pods, sample, threads)[1]  # noqa: E731
        kind_note = ("Go-faithful C++ restatement of allocatable.go:63-168 + resource_allocation.go:49-131 (per-call small "
                     "maps, string switches, NodeScoreList), one scheduling cycle per pinned thread")
    else:
        per_thread = {"c3": 6, "c4": 3, "c5": 1}[cfg]
        sample = max(threads * per_thread, 16)
        d = gen_inputs(synth, cfg, seed, sample, N)
        feas = synth.gen_feasible_words(seed, sample, N, npad_of(N))
        run = lambda: cpu_threads_sample(cfg, d, N, sample, threads, feas)[1]  # noqa: E731
        kind_note = "C port of the plugin arithmetic on flat columns (oracle/*.c), pods split over the host threads"
    for _ in range(args.warmup):
        run()
    dts = [run() for _ in range(args.steps)]
    dt = float(np.mean(dts))
    val = sample * N * plugins_per_pair(cfg) / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": spec["scaling"],
        "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
        "config": {"workload": spec["label"], "config": cfg,
                   "note": f"each step = bounded sample of {sample} pods x {N} nodes of that workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} pods x {N} nodes per step on {threads} threads; {kind_note}; Go toolchain "
                                   "absent, the reference itself cannot run"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if cfg == "c2":
        line["cpu_baseline"]["soa_port_value"] = cpu_threads_sample("c2", dict(nodes=nodes, pods=pods), N, sample, threads, feas)[0]
        line["cpu_baseline"]["soa_port_note"] = "flat-column C port (oracle/alloc.c), same threads — the layout-only speed-up"
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="0 = the config's default (20; c5: 3)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--pods", type=int, default=0)
    ap.add_argument("--nodes", type=int, default=0, help="c2/c3: nodes per GPU; c4/c5: nodes over all GPUs")
    ap.add_argument("--chunk", type=int, default=5000, help="c5: pods per engine batch")
    ap.add_argument("--cycles", type=int, default=1000, help="P=1 scheduling cycles for the latency leg")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 10)")
    ap.add_argument("--parity-pods", type=int, default=0, help="pods sampled for the oracle check (0 = per config)")
    ap.add_argument("--no-peer", action="store_true", help="keep NCCL for the per-pod exchanges (A/B against the peer-memory path)")
    ap.add_argument("--kernel-only", action="store_true", help="profiling runs: skip the e2e, parity and CPU legs")
    args = ap.parse_args()
    args.steps = args.steps or (3 if args.config == "c5" else 20)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    if rank == 0:
        g.build()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import synth

    torch.cuda.set_device(local)
    cfg = args.config
    spec = CONFIGS[cfg]
    P = args.pods or spec["pods"]
    seed = BASE_SEED + spec["no"]
    if spec["scaling"] == "weak":
        n_local = args.nodes or spec["nodes"]
        N_global, off, cnt = n_local * world, rank * n_local, n_local
    else:
        N_global = args.nodes or spec["nodes"]
        off, cnt = shard_bounds(N_global, world)[rank]
    d = gen_inputs(synth, cfg, seed, P, N_global)
    # the caller's upstream feasibility (what in-tree filters left), density 0.875; global, sliced per shard
    feas_global = synth.gen_feasible_words(seed, P, N_global, npad_of(N_global)) if cfg != "c3" else None
    feas_local = None
    if feas_global is not None:
        feas_local = E.pack_bits(E.unpack_bits(feas_global, N_global)[:, off:off + cnt], npad_of(cnt)) if world > 1 else feas_global

    eng = E.Engine(local)
    if world > 1:
        uid = [eng.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(uid[0], rank, world)
        if not args.no_peer:  # symmetric buffers over CUDA IPC: the per-pod exchanges become peer-memory stores
            handles = [None] * world
            dist.all_gather_object(handles, eng.peer_export())
            eng.peer_import(handles)
    upload_snapshot(E, eng, cfg, d, off, cnt, N_global)
    npad = eng.Npad
    words = npad // 64
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device("cuda", local))

    def barrier():
        eng.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- per-config step functions -------------------------------------------------------------------------------
    chunk = min(args.chunk, P) if cfg == "c5" else P
    chunks = [np.arange(a, min(P, a + chunk)) for a in range(0, P, chunk)]
    pins = []  # pinned host staging (what the cgo shim would hold)

    def pinned_copy(a):
        buf = eng.pinned(a.nbytes)
        v = buf.view(a.dtype, a.shape)
        v[...] = a
        pins.append(buf)
        return v

    feas_pin = pinned_copy(feas_local) if feas_local is not None else None
    plugins = {"c2": [E.PLUGIN_ALLOCATABLE], "c3": [E.PLUGIN_TLP, E.PLUGIN_LVRB], "c4": [E.PLUGIN_NRT],
               "c5": [E.PLUGIN_NRT, E.PLUGIN_NETWORK_OVERHEAD, E.PLUGIN_ALLOCATABLE, E.PLUGIN_TLP, E.PLUGIN_LVRB]}[cfg]
    dom = {"c2": (E.PLUGIN_ALLOCATABLE, "alloc_norm_kernel<int64>"), "c3": (E.PLUGIN_LVRB, "lvrb_kernel<u8> over the distinct (cpu, mem) keys + expand_rows_kernel<int64>"),
           "c4": (E.PLUGIN_NRT, "nrt2_q_kernel + nrt2_tableq_kernel x2 + nrt2_expand_kernel<int64>"),
           "c5": (E.PLUGIN_NRT, "nrt2_q_kernel + nrt2_tableq_kernel x2 + nrt2_expand_kernel<u8>")}[cfg]
    topk_rows = {}  # c5: chunk index -> fetched [chunk][1] winners of the last step

    if cfg == "c5":
        batches = []
        for ci, rows in enumerate(chunks):
            dp = slice_pods(d, rows)
            b, keep = eng.make_batch(len(rows), **pod_columns(cfg, dp, pinned_copy(np.ascontiguousarray(feas_local[rows]))))
            batches.append((b, keep, len(rows)))

        # the host prepares chunk i + 1 (request-vector dictionary, staging) while the device evaluates chunk i: uploads
        # queue without synchronising; the batches' arrays stay alive and unchanged in `batches`
        eng.config_async_upload(True)

        def step(fetch=False):
            for ci, (b, _keep, n) in enumerate(batches):
                eng._chk(eng.lib.b200s_pods_upload(eng.ctx, C.byref(b)))  # pod columns + upstream mask of the chunk
                eng.P = n
                eng.eval_combined(0b11111, PROFILE_WEIGHTS, k=1, write_total=False)
                if fetch:
                    topk_rows[ci] = eng.fetch_topk().copy()

        def step_resident():
            step(False)

        def e2e_step():
            step(True)
    else:
        dp_all = slice_pods(d, np.arange(P))
        eng.pods_upload(P, **pod_columns(cfg, dp_all, feas_pin))
        out_dtype = E.OUT_I64

        def step_resident():
            for pl in plugins:
                eng.eval(pl, out_dtype)

        batch, keep = eng.make_batch(P, **pod_columns(cfg, dp_all, feas_pin))
        out8 = [pinned_copy(np.zeros((P, npad), np.uint8)) for _ in plugins]
        feas_out = pinned_copy(np.zeros((P, words), np.uint64)) if cfg == "c4" else None
        reas_out = pinned_copy(np.zeros((P, npad), np.uint8)) if cfg == "c4" else None

        def e2e_step():
            for pl, o in zip(plugins, out8):
                eng.score_batch(pl, batch, E.OUT_U8, o, feas_out, reas_out)

    # ---------------- value: inputs resident in HBM ----------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    eng.sync()
    for pl in plugins:
        eng.kernel_time(pl)
    for ph in range(3):
        eng.phase_time(ph)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    eng.set_profiling(True)
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        step_resident()
    e1.record(ext)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = eng.launches - l0
    k_times = {pl: eng.kernel_time(pl) for pl in plugins}
    phases = {name: eng.phase_time(ph) for name, ph in (("allreduce", 0), ("allgather", 1), ("combine_topk_fold", 2))}
    eng.set_profiling(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    pairs_per_step = P * N_global
    evals_per_step = pairs_per_step * plugins_per_pair(cfg)
    value = evals_per_step / (ms_step * 1e-3)

    if args.kernel_only:
        if rank == 0:
            sampler.stop()
            print(json.dumps({"kernel_only": True, "config": cfg, "value": value, "ms_per_step": ms_step}), flush=True)
        eng.close()
        return

    # ---------------- e2e: host buffers through the C-ABI ------------------------------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(ext)
    for _ in range(e2e_steps):
        e2e_step()
    b.record(ext)
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    tt = torch.tensor([max(wall, a.elapsed_time(b))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_e2e = float(tt.item()) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    if cfg == "c5":
        # upstream mask + TLP / LVRB pod columns + NRT pod records (qos, flags, counts, kinds, masks, 9 x 4 requests) +
        # NetworkOverhead (scoreEqually, CSR offsets, 16-byte dependency entries)
        h2d = P * words * 8 + P * 24 + P * (4 + 8 + 9 + 9 * 4 * 8) + P + (P + len(chunks)) * 4 + len(d["net"]["deps"]) * 16
        d2h = P * 16
        e2e_mode = "TOPK: pod columns + upstream mask in, [P] {score, node} winners out (no matrix D2H)"
    else:
        h2d = (P * words * 8 if feas_pin is not None else 0) + (P * 24 if cfg == "c3" else 0) + (P * (9 * 4 * 8 + 21) if cfg == "c4" else 0)
        d2h = len(plugins) * P * npad + (P * words * 8 + P * npad if cfg == "c4" else 0)
        e2e_mode = "MATRIX: u8 [P][Npad] per plugin back" + (" + feasibility words + reason codes" if cfg == "c4" else "")

    # ---------------- parity: sampled pods of this rank's shard vs the UNSHARDED oracle ------------------------------
    n_par = args.parity_pods or {"c2": 8, "c3": 8, "c4": 8, "c5": 4}[cfg]
    rows = np.sort(np.random.default_rng(seed + 17).choice(P, size=min(n_par, P), replace=False))
    parity_ok, parity_err = True, ""
    try:
        want = oracle_rows(cfg, d, rows, feas_global, N_global)
        sl = slice(off, off + cnt)
        if cfg == "c5":
            e2e_step()  # fresh winners of every chunk
            for r_i, p in enumerate(rows):
                ci, pi = int(p) // chunk, int(p) % chunk
                got = topk_rows[ci][pi][0]
                w = want["topk"][r_i][0]
                if (int(got["score"]), int(got["node"])) != (int(w[0]), int(w[1])):
                    parity_ok, parity_err = False, f"pod {p}: top-1 {(int(got['score']), int(got['node']))} != oracle {w}"
            # and this shard's total matrix for the chunk holding the first sampled pod
            ci = int(rows[0]) // chunk
            b_, _k, n_ = batches[ci]
            eng._chk(eng.lib.b200s_pods_upload(eng.ctx, C.byref(b_)))
            eng.P = n_
            eng.eval_combined(0b11111, PROFILE_WEIGHTS, k=1, write_total=True)
            tot = eng.fetch_total()
            for r_i, p in enumerate(rows):
                if int(p) // chunk == ci and not np.array_equal(tot[int(p) % chunk, :cnt], want["total"][r_i, sl]):
                    parity_ok, parity_err = False, f"pod {p}: total score row differs from the oracle on this shard"
        else:
            names = {"c2": ["alloc"], "c3": ["tlp", "lvrb"], "c4": ["nrt"]}[cfg]
            eng.pods_upload(P, **pod_columns(cfg, dp_all, feas_pin))  # a pipelined score_batch leaves no resident batch
            for pl, nm in zip(plugins, names):
                eng.eval(pl, E.OUT_I64)
                got = eng.fetch_scores(pl)[rows, :cnt]
                if not np.array_equal(got, want[nm][:, sl]):
                    parity_ok, parity_err = False, f"{nm}: sampled rows differ from the oracle on this shard"
                if not np.array_equal(out8[plugins.index(pl)][rows, :cnt].astype(np.int64), want[nm][:, sl]):
                    parity_ok, parity_err = False, f"{nm}: e2e (u8) rows differ from the oracle on this shard"
            if cfg == "c4":
                if not np.array_equal(eng.fetch_reasons(E.PLUGIN_NRT)[rows, :cnt], want["nrt_reasons"][:, sl]):
                    parity_ok, parity_err = False, "nrt: reason codes differ from the oracle on this shard"
    except Exception as e:  # noqa: BLE001 -- reported in the line, never hidden
        parity_ok, parity_err = False, f"{type(e).__name__}: {e}"
    pc = torch.tensor([1 if parity_ok else 0], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)
    parity_checked = int(pc.item())
    errs = [parity_err]
    if world > 1:
        errs = [None] * world
        dist.all_gather_object(errs, parity_err)

    # ---------------- roofline of the dominant kernel -----------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    n_chunks = len(chunks)
    per_launch_pods = chunk
    if cfg == "c2":  # int64 score matrix + mask bits + raw + params
        alg_bytes = P * cnt * 8 + P * ((cnt + 7) // 8) + cnt * 16 + P * 32
    elif cfg == "c3":  # int64 score matrix + node columns (49 B) + pod columns (16 B)
        alg_bytes = P * cnt * 8 + cnt * 49 + P * 16
    elif cfg == "c4":  # int64 score + reason code + feasibility bit; node columns ~166 B, pod records ~140 B
        alg_bytes = P * cnt * (8 + 1) + P * ((cnt + 7) // 8) * 2 + cnt * 166 + P * 140
    else:  # u8 score + reason code + feasibility bit per pair of one chunk
        alg_bytes = per_launch_pods * cnt * (1 + 1) + per_launch_pods * ((cnt + 7) // 8) * 2 + cnt * 166 + per_launch_pods * 140
    # the engine may split one pass into several launches of the same kernel (the sharded Allocatable path goes in
    # pod chunks to hide the all-reduce): time per PASS = summed kernel time / passes, not / launches
    passes = args.steps * n_chunks
    k_ms, k_n = k_times[dom[0]]
    k_avg_ms = k_ms / max(passes, 1)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom[1], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes,
                "kernel_ms": k_avg_ms,
                "kernel_share_of_step": k_avg_ms * (n_chunks if cfg == "c5" else 1) / ms_step if ms_step else None,
                "per_plugin_kernel_ms": {E.PLUGIN_NAMES[pl]: k_times[pl][0] / max(passes, 1) for pl in plugins},
                "kernel_launches_per_pass": k_n / max(passes, 1),
                "phase_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()}}
    traffic_file = {"c2": "r01_alloc_norm_traffic.json", "c4": "r02_nrt2_traffic.json", "c5": "r02_nrt2_traffic.json"}.get(cfg)
    if traffic_file and os.path.exists(os.path.join(ROOT, "profiles", traffic_file)):
        try:
            roofline["traffic"] = json.load(open(os.path.join(ROOT, "profiles", traffic_file))).get("dram_bytes_per_launch")
        except Exception:
            pass

    # ---------------- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        sample = {"c2": 96, "c3": 24, "c4": 12, "c5": 2}[cfg]
        v1, dt1 = cpu_threads_sample(cfg, d, N_global, sample, 1, feas_global)
        cpu = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{sample} pods x {N_global} nodes, scalar single-thread C port of the plugin arithmetic "
                         f"(oracle/*.c, {dt1:.2f} s); Go toolchain absent so the reference itself cannot run",
               "host_cores": os.cpu_count()}
        if cfg == "c2":
            from oracle import pyoracle as orc

            cols = [d["nodes"]["alloc_cpu_milli"], d["nodes"]["alloc_mem_bytes"]]
            vg, dtg = cpu_gofaithful(orc, cols, feas_global, synth.gen_pods(seed, 32), 32, 1)
            cpu["gofaithful_value"] = vg
            cpu["gofaithful_note"] = (f"Go-faithful per-call restatement (oracle/gofaithful.cpp), 32 pods x {N_global} nodes, "
                                      f"1 thread ({dtg:.2f} s)")

    cycle = None
    if rank == 0 and world == 1 and cfg == "c2":
        cycle = cycle_latency(E, synth, local, cnt, args.cycles)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": spec["scaling"],
            "vs_baseline": None, "dtype": spec["dtype"], "data": "synthetic",
            "config": {"workload": spec["label"] + f"; feasibility density 0.875; {P} pods, {N_global} nodes over {world} GPU(s)",
                       "config": cfg, "pods": P, "nodes_global": N_global, "nodes_this_rank": cnt,
                       "parallelism": f"node-sharded x{world}",
                       "exchange": None if world == 1 else ("NCCL" if args.no_peer else "peer memory (CUDA IPC over NVLink), NCCL above 4 MiB"), "plugins_per_pair": plugins_per_pair(cfg),
                       "pairs_per_s": pairs_per_step / (ms_step * 1e-3), "pod_chunk": chunk if cfg == "c5" else None,
                       "value_out": ("int64 [P][Npad] (8 B/eval)" if cfg != "c5" else "per-pod top-1 {score, node} (TOPK mode)"),
                       "l2": f"{P * npad * (8 if cfg != 'c5' else 1) / 1e9:.2f} GB of scores per plugin and step >> 126 MB L2: "
                             "every step streams past L2"},
            "e2e": {"value": evals_per_step / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": ms_e2e, "mode": e2e_mode},
            "gpu_launches": int(launches),
            "parity_checked": parity_checked, "parity_pods": [int(x) for x in rows],
            "parity_errors": [e for e in errs if e],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "cycle_latency": cycle,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    for b_ in pins:
        b_.free()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

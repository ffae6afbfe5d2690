#!/usr/bin/env python
"""Measures every plugin kernel at the BASELINE.json config shapes on ONE GPU (inputs resident in
HBM, CUDA events around the dominant kernel via the engine's profiling hooks) and prints one JSON
line per (config, plugin, dtype).  Not the headline bench (bench.py is); this fills BASELINE.md §4
and DESIGN.md's per-kernel roofline table.

  c1  128 x 1 000     Allocatable Least
  c2  10k x 50k       Allocatable Most + NormalizeScore
  c3  10k x 50k       TargetLoadPacking, LoadVariationRiskBalancing
  c3b 10k x 50k       Peaks (2 passes), LowRiskOverCommitment
  c4  5k x 20k x 4    NodeResourceTopologyMatch Filter + Score (all four strategies)
  c5s 50k x 25k       the per-GPU shard of c5 (50k x 200k over 8 GPUs): all five + combined top-1
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--configs", default="c1,c2,c3,c3b,c4,c5s")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import __graft_entry__ as g

    g.build()
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import synth
    from test_gpu_combined import build_inputs, load_engine

    peak = 6650.0
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    lines = []

    def emit(cfg, plugin, dtype, P, N, ms, nl, alg_bytes, extra=None):
        k_ms = ms / max(nl, 1)
        line = dict(config=cfg, plugin=plugin, out=dtype, pods=P, nodes=N, kernel_ms=round(k_ms, 4),
                    evals_per_s=P * N / (k_ms * 1e-3) if k_ms else None,
                    alg_gbps=alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None,
                    hbm_frac=(alg_bytes / (k_ms * 1e-3) / 1e9) / peak if k_ms else None, launches=nl)
        if extra:
            line.update(extra)
        lines.append(line)
        print(json.dumps(line), flush=True)

    def time_plugin(eng, cfg, plugin, name, P, N, bytes_per_eval_in=0.0):
        for dtype, dn, ob in ((E.OUT_I64, "i64", 8), (E.OUT_U8, "u8", 1)):
            for _ in range(3):
                eng.eval(plugin, dtype)
            eng.sync()
            eng.kernel_time(plugin)
            eng.set_profiling(True)
            t0 = time.perf_counter()
            for _ in range(args.steps):
                eng.eval(plugin, dtype)
            eng.sync()
            wall = (time.perf_counter() - t0) * 1e3 / args.steps
            ms, nl = eng.kernel_time(plugin)
            eng.set_profiling(False)
            alg = P * N * (ob + bytes_per_eval_in)
            emit(cfg, name, dn, P, N, ms, nl, alg, dict(step_ms_wall=round(wall, 4)))

    cfgs = args.configs.split(",")
    eng = E.Engine(0)
    if "c1" in cfgs or "c2" in cfgs:
        for cfg, P, N, mode in (("c1", 128, 1000, 0), ("c2", 10_000, 50_000, 1)):
            if cfg not in cfgs:
                continue
            nodes = synth.gen_nodes(synth.BASE_SEED + int(cfg[1]), N)
            eng.snapshot_begin(N)
            eng.snapshot_allocatable([nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]])
            eng.snapshot_commit()
            eng.config_allocatable(mode, [1 << 20, 1])
            feas = synth.gen_feasible_words(synth.BASE_SEED + int(cfg[1]), P, N, eng.Npad)
            eng.pods_upload(P, feasible=feas)
            time_plugin(eng, cfg, E.PLUGIN_ALLOCATABLE, "NodeResourcesAllocatable", P, N, 1 / 8)
    if "c3" in cfgs:
        P, N = 10_000, 50_000
        seed = synth.BASE_SEED + 3
        nodes, pods = synth.gen_nodes(seed, N), synth.gen_pods(seed, P)
        tri = synth.gen_trimaran(seed, nodes)
        eng.snapshot_begin(N)
        eng.snapshot_tlp(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["missing_milli"], tri["tlp_flags"])
        eng.snapshot_lvrb(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                          nodes["alloc_mem_bytes"], tri["lvrb_flags"])
        eng.snapshot_commit()
        eng.config_tlp(40)
        eng.config_lvrb(1.0, 1.0)
        eng.pods_upload(P, tlp_pod_cpu_milli=pods["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=pods["req_cpu_milli"],
                        lvrb_req_mem_bytes=pods["req_mem_bytes"])
        time_plugin(eng, "c3", E.PLUGIN_TLP, "TargetLoadPacking", P, N)
        time_plugin(eng, "c3", E.PLUGIN_LVRB, "LoadVariationRiskBalancing", P, N)
    if "c3b" in cfgs:  # the rest of the Trimaran family on the c3 shape
        P, N = 10_000, 50_000
        seed = synth.BASE_SEED + 3
        nodes = synth.gen_nodes(seed, N)
        tri, t2 = synth.gen_trimaran(seed, nodes), synth.gen_trimaran2(seed, nodes, P)
        eng.snapshot_begin(N)
        eng.snapshot_peaks(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["tlp_flags"], t2["k1"], t2["k2"])
        eng.snapshot_low_risk(tri["cpu_avg"], tri["cpu_std"], tri["mem_avg"], tri["mem_std"], nodes["alloc_cpu_milli"],
                              nodes["alloc_mem_bytes"], tri["lvrb_flags"], t2["node_req_cpu"], t2["node_req_mem"],
                              t2["node_lim_cpu"], t2["node_lim_mem"])
        eng.snapshot_commit()
        eng.config_low_risk(5, 0.5, 0.5)
        eng.pods_upload(P, feasible=synth.gen_feasible_words(seed, P, N, eng.Npad),
                        peaks_pod_cpu_milli=t2["peaks_pod_cpu_milli"], low_risk_pod=t2["low_risk_pod"])
        time_plugin(eng, "c3b", E.PLUGIN_PEAKS, "Peaks", P, N)
        time_plugin(eng, "c3b", E.PLUGIN_LOW_RISK, "LowRiskOverCommitment", P, N)
    if "c4" in cfgs:
        P, N = 5_000, 20_000
        seed = synth.BASE_SEED + 4
        nn, npods = synth.gen_nrt(seed, N, P, Z=4)
        eng.snapshot_begin(N)
        eng.snapshot_nrt(nn)
        eng.snapshot_commit()
        eng.pods_upload(P, nrt=npods)
        for strat, sn in ((2, "LeastAllocated"), (0, "MostAllocated"), (1, "BalancedAllocation"), (3, "LeastNUMANodes")):
            eng.config_nrt(strat, [1, 1, 1, 1])
            time_plugin(eng, "c4", E.PLUGIN_NRT, f"NodeResourceTopologyMatch/{sn}", P, N, 1 + 1 / 8)
    if "c5s" in cfgs:
        P, N = 50_000, 25_000
        seed = synth.BASE_SEED + 5
        d = build_inputs(seed, P, N)
        feas = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
        load_engine(eng, E, d, N, P, feas, node_offset=0, n_global=N)
        time_plugin(eng, "c5s", E.PLUGIN_NETWORK_OVERHEAD, "NetworkOverhead(raw pass)", P, N, 8 + 1 + 1 / 8)
        w = [1, 1, 1, 1, 5]
        for _ in range(2):
            eng.eval_combined(0b11111, w, k=1, write_total=False)
        eng.sync()
        t0 = time.perf_counter()
        for _ in range(max(2, args.steps // 2)):
            eng.eval_combined(0b11111, w, k=1, write_total=False)
        eng.sync()
        ms = (time.perf_counter() - t0) * 1e3 / max(2, args.steps // 2)
        line = dict(config="c5s", plugin="combined(5 plugins, top-1, no total matrix)", pods=P, nodes=N,
                    step_ms_wall=round(ms, 3), evals_per_s=P * N / (ms * 1e-3), plugin_evals_per_s=5 * P * N / (ms * 1e-3))
        lines.append(line)
        print(json.dumps(line), flush=True)
    eng.close()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(lines, f, indent=1)


if __name__ == "__main__":
    main()

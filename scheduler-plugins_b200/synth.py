"""Seeded synthetic snapshot generator (SURVEY.md §8d "Synthetic inputs").

One generator, three consumers: the CUDA engine, the CPU oracle and the CPU baseline all read
the SAME columns, so "same synthetic snapshot" holds by construction.  Everything is drawn from
numpy's PCG64 with seed 0xB2005EED + config#, one independent child stream per column group
(SeedSequence.spawn order: nodes, pods, trimaran, feasible, nrt, netoh) so adding a consumer
never shifts another group's draws.

Units follow the reference: cpu in milli-cores, memory in bytes (resource_allocation.go:79-100),
Trimaran utilisation in percent quantised to 1/1024 (exact in float64).
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 0xB2005EED
GiB = 1 << 30
MiB = 1 << 20


def _streams(seed: int):
    ss = np.random.SeedSequence(seed)
    return [np.random.Generator(np.random.PCG64(s)) for s in ss.spawn(6)]


def gen_nodes(seed: int, N: int) -> dict:
    """Node capacity/allocatable columns."""
    g = _streams(seed)[0]
    cores = g.choice(np.array([8, 16, 32, 64, 96, 128]), size=N, p=[0.10, 0.20, 0.30, 0.25, 0.10, 0.05])
    alloc_cpu = cores.astype(np.int64) * 1000
    cap_cpu = alloc_cpu + g.choice(np.array([0, 500, 1000]), size=N).astype(np.int64)
    alloc_mem = cores.astype(np.int64) * g.choice(np.array([2, 4, 8]), size=N).astype(np.int64) * GiB \
        - g.integers(0, 2049, size=N).astype(np.int64) * MiB
    eph = g.integers(100, 2001, size=N).astype(np.int64) * GiB
    return dict(N=N, alloc_cpu_milli=alloc_cpu, cap_cpu_milli=cap_cpu, alloc_mem_bytes=alloc_mem,
                alloc_ephemeral_bytes=eph)


def gen_pods(seed: int, P: int) -> dict:
    """Pod-level effective requests (cpu milli, memory bytes)."""
    g = _streams(seed)[1]
    cpu_choices = np.array([100, 250, 500, 1000, 2000, 4000, 8000])
    zipf = 1.0 / np.arange(1, len(cpu_choices) + 1)
    cpu = g.choice(cpu_choices, size=P, p=zipf / zipf.sum()).astype(np.int64)
    mem = (np.int64(128) << g.integers(0, 9, size=P).astype(np.int64)) * MiB  # 128Mi .. 32Gi
    # TLP: predicted utilisation = limits if set else round(requests*1.5) (targetloadpacking.go:198-205)
    has_limit = g.random(P) < 0.5
    limit = cpu * g.choice(np.array([1, 2]), size=P).astype(np.int64)
    tlp_cpu = np.where(has_limit, limit, np.round(cpu * 1.5).astype(np.int64))
    return dict(P=P, req_cpu_milli=cpu, req_mem_bytes=mem.astype(np.int64), tlp_pod_cpu_milli=tlp_cpu.astype(np.int64))


def gen_trimaran(seed: int, nodes: dict) -> dict:
    """load-watcher metrics flattened per GetResourceData / TLP's CPU-metric scan."""
    g = _streams(seed)[2]
    N = nodes["N"]
    q = 1024.0
    cpu_avg = np.round(g.beta(2, 3, N) * 100 * q) / q
    cpu_std = np.round(g.uniform(0, 25, N) * q) / q
    mem_avg = np.round(g.beta(3, 3, N) * 100 * q) / q
    mem_std = np.round(g.uniform(0, 15, N) * q) / q
    u = g.random(N)
    no_metrics = u < 0.02
    no_mem = (u >= 0.02) & (u < 0.03)
    miss_sel = g.random(N) >= 0.90
    missing = np.where(miss_sel, g.integers(100, 4001, N), 0).astype(np.int64)
    tlp_flags = np.where(no_metrics, 0, 3).astype(np.uint8)
    lvrb_flags = np.where(no_metrics, 0, np.where(no_mem, 1 | 2, 1 | 2 | 4)).astype(np.uint8)
    return dict(cpu_avg=cpu_avg, cpu_std=cpu_std, mem_avg=mem_avg, mem_std=mem_std, missing_milli=missing,
                tlp_flags=tlp_flags, lvrb_flags=lvrb_flags)


def gen_trimaran2(seed: int, nodes: dict, P: int) -> dict:
    """Peaks power models and LowRiskOverCommitment request/limit sums, on top of gen_nodes / gen_trimaran."""
    g = np.random.default_rng([seed, 77])
    N = nodes["N"]
    # power = k0 + k1 * exp(k2 * util%): the shape of the reference's sample model (peaks_test.go:81-87), per node
    k1 = -np.round(g.uniform(40, 140, N) * 4096) / 4096
    k2 = -np.round(g.uniform(0.01, 0.12, N) * 65536) / 65536
    no_model = g.random(N) < 0.03  # getPowerModel: nodes without an entry score with {0, 0, 0}
    k1[no_model] = 0
    k2[no_model] = 0
    cap_c, cap_m = nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]
    # pods already on the node: requests between 10 % and 130 % of capacity (the cap branch), limits >= requests,
    # some nodes not overcommitted at all (limit < capacity: the conditioning branch), some empty
    fr = g.uniform(0.1, 1.3, (2, N))
    over = g.uniform(1.0, 2.5, (2, N))
    req_c = (cap_c * fr[0]).astype(np.int64) // 50 * 50
    req_m = (cap_m * fr[1]).astype(np.int64) >> 20 << 20
    lim_c = (req_c * over[0]).astype(np.int64) // 50 * 50
    lim_m = (req_m * over[1]).astype(np.int64) >> 20 << 20
    empty = g.random(N) < 0.05
    for a in (req_c, req_m, lim_c, lim_m):
        a[empty] = 0
    pr_c = g.choice(np.array([0, 100, 250, 500, 1000, 2000, 4000]), size=P).astype(np.int64)
    pr_m = (g.choice(np.array([0, 128, 256, 1024, 4096]), size=P) << 20).astype(np.int64)
    pl_c = np.maximum(pr_c, g.choice(np.array([0, 500, 2000, 8000]), size=P)).astype(np.int64)  # SetMaxLimits
    pl_m = np.maximum(pr_m, (g.choice(np.array([0, 512, 8192]), size=P) << 20)).astype(np.int64)
    be = g.random(P) < 0.1  # best-effort pods score 0 everywhere
    for a in (pr_c, pr_m, pl_c, pl_m):
        a[be] = 0
    return dict(k1=k1, k2=k2, node_req_cpu=req_c, node_req_mem=req_m, node_lim_cpu=lim_c, node_lim_mem=lim_m,
                low_risk_pod=np.stack([pr_c, pr_m, pl_c, pl_m]), peaks_pod_cpu_milli=pr_c.copy())


def gen_feasible_words(seed: int, P: int, N: int, npad: int, k_or: int = 3) -> np.ndarray:
    """Upstream feasibility (what the filters before the Score phase left) as packed words
    [P][npad/64] uint64, bit j of word w = node 64*w+j.  Each bit is the OR of k_or fair bits:
    density 1 - 2^-k_or (0.875 for the default).  Padding bits (node >= N) are 0."""
    g = _streams(seed)[3]
    words = npad // 64
    m = np.zeros((P, words), dtype=np.uint64)
    for _ in range(k_or):
        m |= g.integers(0, np.iinfo(np.uint64).max, size=(P, words), dtype=np.uint64, endpoint=True)
    full, rem = divmod(N, 64)
    if rem:
        m[:, full] &= np.uint64((1 << rem) - 1)
        m[:, full + 1:] = 0
    else:
        m[:, full:] = 0
    return m


def gen_netoh(seed: int, N: int, P: int, n_regions: int = 8, zones_per_region: int = 8, n_global: int | None = None,
              max_deps: int = 8) -> dict:
    """NetworkOverhead inputs: 3-tier topology (same host 0 / same zone 1 / zone matrix inside a
    region / region matrix / MaxCost 100), AppGroup dependencies already matched against the
    placed pods.  Name dictionary: id 0 = empty label, 1..R = regions, R+1.. = zones."""
    g = _streams(seed)[5]
    n_global = n_global or N
    R, Z = n_regions, n_regions * zones_per_region
    K = 1 + R + Z
    MISSING = -(2**63)
    region_cost = np.full((K, K), MISSING, dtype=np.int64)
    zone_cost = np.full((K, K), MISSING, dtype=np.int64)
    rc = g.integers(20, 101, size=(R, R))
    miss = g.random((R, R)) < 0.05
    for a in range(R):
        for b in range(R):
            if a != b and not miss[a, b]:
                region_cost[1 + a, 1 + b] = rc[a, b]
    zc = g.integers(2, 20, size=(Z, Z))
    for a in range(Z):
        for b in range(Z):
            if a != b and a // zones_per_region == b // zones_per_region:
                zone_cost[1 + R + a, 1 + R + b] = zc[a, b]
    zone_of = g.integers(0, Z, size=n_global)
    unl = g.random(n_global) < 0.005
    region_all = np.where(unl, 0, 1 + zone_of // zones_per_region).astype(np.uint16)
    zone_all = np.where(unl, 0, 1 + R + zone_of).astype(np.uint16)
    # a few nodes with a zone label but no region label ("same region" incl. the empty string, :540)
    odd = g.random(n_global) < 0.002
    region_all = np.where(odd, 0, region_all).astype(np.uint16)
    score_equally = (g.random(P) < 0.2).astype(np.uint8)
    nd = np.where(score_equally == 1, 0, g.integers(1, max_deps + 1, size=P))
    off = np.zeros(P + 1, dtype=np.int32)
    off[1:] = np.cumsum(nd)
    total = int(off[-1])
    deps = np.zeros(total, dtype=[("host_node", "<i4"), ("host_region", "<u2"), ("host_zone", "<u2"),
                                  ("max_network_cost", "<i8")])
    hosts = g.integers(0, n_global, size=total)
    deps["host_node"] = hosts
    deps["host_region"] = region_all[hosts]
    deps["host_zone"] = zone_all[hosts]
    deps["max_network_cost"] = g.choice(np.array([0, 10, 30, 100]), size=total)
    return dict(K=K, zone_cost=zone_cost, region_cost=region_cost, region_all=region_all, zone_all=zone_all,
                score_equally=score_equally, dep_offset=off, deps=deps)


def gen_nrt(seed: int, N: int, P: int, Z: int = 4, with_cost: bool = True) -> tuple:
    """NodeResourceTopologyMatch inputs directly in the dense encoding (include/b200sched.h):
    resource slots 0 cpu, 1 memory, 2 hugepages-2Mi, 3 vendor/nic1 (device).  Returns (nodes, pods)."""
    g = _streams(seed)[4]
    R, C = 4, 8
    base = gen_nodes(seed, N)
    cores = base["alloc_cpu_milli"] // 1000
    res_flags = np.array([1, 1, 1, 2], dtype=np.uint8)  # affine, affine, affine, host-level (non-native)
    node_flags = np.full(N, 1 | 2 | 4, dtype=np.uint8)  # hasNRT | fresh | single-numa-node
    node_flags |= (g.random(N) < 0.5).astype(np.uint8) * 8  # scope pod for half of the nodes
    u = g.random(N)
    node_flags[u < 0.01] &= ~np.uint8(1)       # 1 %: no NRT object
    node_flags[(u >= 0.01) & (u < 0.015)] &= ~np.uint8(2)  # 0.5 %: stale
    node_flags[(u >= 0.015) & (u < 0.02)] &= ~np.uint8(4)  # 0.5 %: policy none
    node_flags[(u >= 0.02) & (u < 0.022)] |= 16  # unsupported shape -> host fallback
    nz = np.full(N, Z, dtype=np.uint8)
    nz[g.random(N) < 0.05] = max(1, Z // 2)
    avail = np.zeros((Z, R, N), dtype=np.int64)
    zmask = np.zeros((Z, N), dtype=np.uint8)
    for z in range(Z):
        has = z < nz
        cpu = np.maximum(cores // Z - g.integers(0, 5, N), 0) * 1000
        frac = g.random(N) < 0.10
        cpu = np.where(frac, cpu + g.integers(1, 1000, N), cpu)
        mem = np.maximum(base["alloc_mem_bytes"] // Z - g.integers(0, 4, N) * GiB, 0) * 1000
        hp = g.choice(np.array([0, 512 * MiB, GiB]), size=N) * 1000
        dev = np.where((z % 2) == 0, g.integers(0, 9, N), 0) * 1000
        avail[z, 0], avail[z, 1], avail[z, 2], avail[z, 3] = cpu, mem, hp, dev
        m = np.full(N, 1 | 2 | 4, dtype=np.uint8)
        m |= np.where((z % 2) == 0, 8, 0).astype(np.uint8)  # device listed by 2 of 4 zones
        zmask[z] = np.where(has, m, 0)
        avail[z][:, ~has] = 0
    node_res_mask = np.full(N, 1 | 2 | 4 | 8, dtype=np.uint8)
    node_res_mask[g.random(N) < 0.02] &= ~np.uint8(8)  # device missing at node level
    cost = np.full((Z, Z, N), -1, dtype=np.int32)
    for a in range(Z):
        for b in range(Z):
            cost[a, b] = 10 if a == b else (12 if a // 2 == b // 2 else 20)
    cost[:, :, g.random(N) < 0.01] = -1  # no SLIT data
    nodes = dict(n_zones=Z, n_res=R, res_flags=res_flags, node_flags=node_flags,
                 max_numa=np.where(g.random(N) < 0.9, 8, 4).astype(np.uint16), n_zones_node=nz,
                 node_res_mask=node_res_mask, zone_res_mask=zmask, avail=avail, cost=cost if with_cost else None)

    qos = g.choice(np.array([0, 1, 2]), size=P, p=[0.6, 0.3, 0.1]).astype(np.uint8)
    n_app = g.integers(1, 5, P).astype(np.uint8)
    n_init = g.integers(0, 3, P).astype(np.uint8)
    kind = np.zeros((P, C), dtype=np.uint8)
    req = np.zeros((P, C + 1, R), dtype=np.int64)
    rmask = np.zeros((P, C + 1), dtype=np.uint8)
    flags = np.zeros(P, dtype=np.uint8)
    cpu_ch = np.array([100, 250, 500, 1000, 2000, 4000, 8000])
    for p in range(P):
        ni, na = int(n_init[p]), int(n_app[p])
        non_native = False
        for c in range(ni + na):
            if c < ni:
                kind[p, c] = 2 if g.random() < 0.3 else 1
            if qos[p] != 2:
                cpu = int(g.choice(cpu_ch)) if qos[p] == 1 or g.random() < 0.1 else int(g.integers(1, 9)) * 1000
                req[p, c, 0], req[p, c, 1] = cpu, int(128 << g.integers(0, 8)) * MiB * 1000
                rmask[p, c] |= 3
                if g.random() < 0.2:
                    req[p, c, 2] = int(g.choice([0, 64, 256, 512])) * MiB * 1000
                    rmask[p, c] |= 4
            if g.random() < (0.25 if qos[p] != 2 else 0.5):
                req[p, c, 3] = int(g.integers(0, 5)) * 1000
                rmask[p, c] |= 8
                non_native = True
        if qos[p] == 2 and not non_native:
            flags[p] |= 1  # BestEffort with only native resources: Filter passes (filter.go:181)
        # GetPodEffectiveRequest: max(sum app, max init) per resource (pkg/util/resource.go:51-85)
        app_sum = req[p, ni:ni + na].sum(axis=0)
        app_mask = np.bitwise_or.reduce(rmask[p, ni:ni + na]) if na else 0
        init_max = req[p, :ni].max(axis=0) if ni else np.zeros(R, dtype=np.int64)
        init_mask = np.bitwise_or.reduce(rmask[p, :ni]) if ni else 0
        for r in range(R):
            in_app, in_init = (app_mask >> r) & 1, (init_mask >> r) & 1
            if in_app and in_init:
                req[p, C, r] = max(app_sum[r], init_max[r])
            elif in_app:
                req[p, C, r] = app_sum[r]
            elif in_init:
                req[p, C, r] = init_max[r]
        rmask[p, C] = app_mask | init_mask
    flags[g.random(P) < 0.002] |= 2  # unsupported pods
    pods = dict(qos=qos, flags=flags, n_init=n_init, n_app=n_app, cont_kind=kind, req_mask=rmask, req=req)
    return nodes, pods

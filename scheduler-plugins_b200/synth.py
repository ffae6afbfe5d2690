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

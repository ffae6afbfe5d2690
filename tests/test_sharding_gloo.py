"""CPU, world_size 2, gloo: the host-side logic of the node-sharded path — shard bounds, global
node offsets, the min/max exchange for NormalizeScore and the top-k fold — against the unsharded
oracle.  (The device path does the same exchanges with NCCL; tests/mgpu_worker.py covers it.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

P, N, K = 24, 1000, 3
W = [1 << 20, 1]


def _inputs():
    import sys

    sys.path.insert(0, ROOT)
    from scheduler_plugins_b200 import synth

    nodes = synth.gen_nodes(123, N)
    return nodes, [nodes["alloc_cpu_milli"], nodes["alloc_mem_bytes"]]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, ROOT)
    from oracle import pyoracle as orc
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import sharding

    _, cols = _inputs()
    off, cnt = sharding.shard_bounds(N, world)[rank]
    local = [c[off:off + cnt] for c in cols]
    # local raw scores and per-pod local min/max (every node feasible here)
    raw = np.array([orc.alloc_score([local[0][i], local[1][i]], W, 1) for i in range(cnt)], dtype=np.int64)
    lo = torch.tensor([raw.min()] * P)
    hi = torch.tensor([raw.max()] * P)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    rng = hi - lo
    norm = ((torch.tensor(raw)[None, :] - lo[:, None]) * 100) // rng[:, None]  # non-negative: floor == trunc
    # local top-k with GLOBAL node indices, then all-gather + fold
    cand = np.zeros((P, K), dtype=E.TOPK_DTYPE)
    for p in range(P):
        order = np.lexsort((np.arange(cnt), -norm[p].numpy()))[:K]
        cand[p]["score"] = norm[p].numpy()[order]
        cand[p]["node"] = off + order
    t = torch.from_numpy(cand.view(np.uint8).reshape(-1).copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    per_rank = [g.numpy().view(E.TOPK_DTYPE).reshape(P, K) for g in gathered]
    folded = sharding.fold_topk(per_rank, K)
    if rank == 0:
        q.put((norm.numpy(), folded, off, cnt))
    else:
        q.put((norm.numpy(), None, off, cnt))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    import sys

    sys.path.insert(0, ROOT)
    from scheduler_plugins_b200 import sharding

    for n, w in [(1000, 2), (50_000, 8), (200_000, 8), (100, 4), (128, 3), (1, 2)]:
        b = sharding.shard_bounds(n, w)
        assert sum(c for _, c in b) == n and all(o % 128 == 0 for o, c in b if c)
        assert [o for o, _ in b] == sorted(o for o, _ in b)


def test_two_rank_sharded_normalize_and_topk(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    _, cols = _inputs()
    want = oracle.alloc_batch(cols, W, 1, P, None, pitch=N)
    got = np.zeros((P, N), dtype=np.int64)
    folded = None
    for norm, f, off, cnt in res:
        got[:, off:off + cnt] = norm
        folded = f if f is not None else folded
    assert np.array_equal(got, want)  # sharded NormalizeScore == unsharded
    for p in range(P):
        order = np.lexsort((np.arange(N), -want[p]))[:K]
        assert list(folded[p]["node"]) == list(order) and list(folded[p]["score"]) == list(want[p][order])


def _peaks_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys

    sys.path.insert(0, ROOT)
    from oracle import pyoracle as orc
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import sharding, synth

    nodes = synth.gen_nodes(321, N)
    tri, t2 = synth.gen_trimaran(321, nodes), synth.gen_trimaran2(321, nodes, P)
    feas = E.unpack_bits(synth.gen_feasible_words(321, P, N, E.npad_of(N)), N)
    off, cnt = sharding.shard_bounds(N, world)[rank]
    sl = slice(off, off + cnt)
    # raw Peaks scores of this shard, local min/max over the feasible nodes, the exchange, then NormalizeScore
    raw = np.array([[orc.peaks_score(tri["cpu_avg"][n], int(nodes["cap_cpu_milli"][n]), int(tri["tlp_flags"][n]),
                                     t2["k1"][n], t2["k2"][n], int(t2["peaks_pod_cpu_milli"][p]))
                     for n in range(off, off + cnt)] for p in range(P)], dtype=np.int64)
    f = feas[:, sl].astype(bool)
    big = np.iinfo(np.int64)
    lo = torch.from_numpy(np.where(f, raw, big.max).min(axis=1))
    hi = torch.from_numpy(np.where(f, raw, big.min).max(axis=1))
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out = np.zeros_like(raw)
    for p in range(P):
        l, h = int(lo[p]), int(hi[p])
        for i in np.nonzero(f[p])[0]:
            s = int(raw[p, i])
            if l == 0 and h == 0:
                out[p, i] = s
            elif h != l:
                out[p, i] = 100 - int(100.0 * float(s - l) / float(h - l))
            else:
                out[p, i] = 100
    q.put((out, off, cnt))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_peaks_normalize(oracle):
    """Peaks.NormalizeScore (peaks.go:152-168) over a node-sharded list: per-shard min/max + one exchange == unsharded."""
    import sys

    sys.path.insert(0, ROOT)
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_peaks_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    nodes = synth.gen_nodes(321, N)
    tri, t2 = synth.gen_trimaran(321, nodes), synth.gen_trimaran2(321, nodes, P)
    want = oracle.peaks_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["tlp_flags"], t2["k1"], t2["k2"],
                              t2["peaks_pod_cpu_milli"], synth.gen_feasible_words(321, P, N, E.npad_of(N)), pitch=N)
    got = np.zeros((P, N), dtype=np.int64)
    for out, off, cnt in res:
        got[:, off:off + cnt] = out
    assert np.array_equal(got, want)

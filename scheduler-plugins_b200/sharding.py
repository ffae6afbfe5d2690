"""Node-axis sharding helpers (host side of SURVEY §8e).

The node axis is split into `world` contiguous shards whose sizes are multiples of 128 (the
engine's row alignment) except the last; pod columns and the cost matrices are replicated.
`fold_topk` is the host restatement of the device fold that follows the NCCL all-gather — used
when a caller gathers the per-shard winners itself (and by the gloo tests).
"""
from __future__ import annotations

import numpy as np

ALIGN = 128


def shard_bounds(n_nodes: int, world: int):
    """[(offset, count)] per rank: contiguous, 128-aligned starts, sizes differ by at most one block."""
    blocks = (n_nodes + ALIGN - 1) // ALIGN
    base, extra = divmod(blocks, world)
    out, off = [], 0
    for r in range(world):
        nb = base + (1 if r < extra else 0)
        cnt = max(0, min(nb * ALIGN, n_nodes - off))
        out.append((off, cnt))
        off += cnt
    assert off == n_nodes
    return out


def fold_minmax(los, his):
    """per-rank per-pod (lo, hi) -> global (min of lo, max of hi): the ncclMin/ncclMax all-reduce."""
    return np.min(np.stack(los), axis=0), np.max(np.stack(his), axis=0)


def fold_topk(per_rank, k: int):
    """per_rank: list over ranks of [P][k] arrays with fields score (i8) / node (i4, -1 = none).
    Returns [P][k] under (score desc, node asc)."""
    allc = np.concatenate(per_rank, axis=1)  # [P][world*k]
    P = allc.shape[0]
    out = np.zeros((P, k), dtype=allc.dtype)
    out["node"] = -1
    for p in range(P):
        c = allc[p][allc[p]["node"] >= 0]
        order = np.lexsort((c["node"], -c["score"]))[:k]
        out[p, :len(order)] = c[order]
    return out

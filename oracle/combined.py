"""Oracle: the upstream scheduling cycle over the five plugins (numpy restatement).

TEST INFRASTRUCTURE (see oracle/oracle.h).  Restates k8s.io/kubernetes pkg/scheduler
findNodesThatFitPod / prioritizeNodes / selectHost [upstream, not in the reference tree]:
filters are ANDed, Score plugins run on the feasible nodes, NormalizeScore over that list,
total = sum weight*score, host = arg-max.  Upstream breaks ties at random; parity is defined on
the score vectors plus a deterministic order (total desc, node index asc) — SURVEY §8c(iii).
"""
from __future__ import annotations

import numpy as np

from . import pyoracle as orc
from . import pyoracle_nrt


def unpack(words, n):
    P = words.shape[0]
    return np.unpackbits(words.view(np.uint8).reshape(P, -1), axis=1, bitorder="little")[:, :n].astype(bool)


def combined(P, N, pitch, upstream_words, weights, k, alloc=None, tlp=None, lvrb=None, nrt=None, netoh=None,
             node_offset=0):
    """Each plugin argument is a dict of that plugin's oracle-batch inputs or None (disabled).
    Returns (total [P][pitch] int64, feasible words [P][pitch/64], topk list of [(score,node)]*k per pod)."""
    cur = upstream_words
    scores = {}
    if nrt is not None:
        s, f, _ = pyoracle_nrt.nrt_batch(nrt["nodes"], nrt["pods"], nrt["strategy"], nrt.get("weights"), cur, pitch=pitch)
        scores[3], cur = s, f
    if netoh is not None:
        s, f, _ = orc.netoh_batch(netoh["zone_cost"], netoh["region_cost"], netoh["region_id"], netoh["zone_id"],
                                  netoh["score_equally"], netoh["dep_offset"], netoh["deps"], cur, pitch=pitch,
                                  node_offset=node_offset)
        scores[4], cur = s, f
    if alloc is not None:
        scores[0] = orc.alloc_batch(alloc["cols"], alloc["weights"], alloc["mode"], P, cur, pitch=pitch)
    if tlp is not None:
        scores[1] = orc.tlp_batch(tlp["util"], tlp["cap"], tlp["missing"], tlp["flags"], tlp["pod_cpu"], tlp["target"],
                                  pitch=pitch)
    if lvrb is not None:
        scores[2] = orc.lvrb_batch(*lvrb["node_cols"], lvrb["req_cpu"], lvrb["req_mem"], lvrb["margin"], lvrb["sens"],
                                   pitch=pitch)
    if cur is None:
        feas = np.zeros((P, pitch), dtype=bool)
        feas[:, :N] = True
        cur = np.packbits(feas, axis=1, bitorder="little").view(np.uint64).reshape(P, pitch // 64)
    fb = np.zeros((P, pitch), dtype=bool)
    fb[:, :N] = unpack(cur, N)
    total = np.zeros((P, pitch), dtype=np.int64)
    for j, s in scores.items():
        total += np.int64(weights[j]) * s
    total[~fb] = 0
    topk = []
    for p in range(P):
        idx = np.nonzero(fb[p])[0]
        order = np.lexsort((idx, -total[p, idx]))[:k]
        row = [(int(total[p, idx[i]]), int(node_offset + idx[i])) for i in order]
        row += [(0, -1)] * (k - len(row))
        topk.append(row)
    return total, cur, topk

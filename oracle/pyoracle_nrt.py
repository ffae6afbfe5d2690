"""ctypes wrappers for the NodeResourceTopologyMatch oracle (TEST INFRASTRUCTURE, see oracle.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .pyoracle import _c, _p, lib


class _NodesSoA(C.Structure):
    _fields_ = [("n_zones", C.c_int32), ("n_res", C.c_int32), ("res_flags", C.c_void_p), ("node_flags", C.c_void_p),
                ("max_numa", C.c_void_p), ("n_zones_node", C.c_void_p), ("node_res_mask", C.c_void_p),
                ("zone_res_mask", C.c_void_p), ("avail", C.c_void_p), ("cost", C.c_void_p)]


class _PodsSoA(C.Structure):
    _fields_ = [("qos", C.c_void_p), ("flags", C.c_void_p), ("n_init", C.c_void_p), ("n_app", C.c_void_p),
                ("cont_kind", C.c_void_p), ("req_mask", C.c_void_p), ("req", C.c_void_p)]


def nrt_batch(nodes: dict, pods: dict, strategy: int, weights=None, feasible_words=None, pitch=None):
    """nodes / pods: the SoA dicts of oracle.flatten.flatten_nrt_nodes / flatten_nrt_pods (also what
    the engine's snapshot_nrt / pods_upload take).  Returns (scores, feasible words, reasons)."""
    Z, R = int(nodes["n_zones"]), int(nodes["n_res"])
    N, P = len(nodes["node_flags"]), len(pods["qos"])
    pitch = pitch or (max(N, 1) + 127) // 128 * 128
    keep = dict(
        res_flags=_c(nodes["res_flags"], np.uint8), node_flags=_c(nodes["node_flags"], np.uint8),
        max_numa=_c(nodes["max_numa"], np.uint16), n_zones_node=_c(nodes["n_zones_node"], np.uint8),
        node_res_mask=_c(nodes["node_res_mask"], np.uint8), zone_res_mask=_c(nodes["zone_res_mask"], np.uint8),
        avail=_c(nodes["avail"], np.int64),
        cost=None if nodes.get("cost") is None else _c(nodes["cost"], np.int32),
        qos=_c(pods["qos"], np.uint8), flags=_c(pods["flags"], np.uint8), n_init=_c(pods["n_init"], np.uint8),
        n_app=_c(pods["n_app"], np.uint8), cont_kind=_c(pods["cont_kind"], np.uint8),
        req_mask=_c(pods["req_mask"], np.uint8), req=_c(pods["req"], np.int64))
    assert keep["avail"].shape == (Z, R, N) and keep["req"].shape == (P, 9, R)
    ns = _NodesSoA(Z, R, *[(keep[k].ctypes.data if keep[k] is not None else None) for k in
                           ("res_flags", "node_flags", "max_numa", "n_zones_node", "node_res_mask", "zone_res_mask",
                            "avail", "cost")])
    ps = _PodsSoA(*[keep[k].ctypes.data for k in ("qos", "flags", "n_init", "n_app", "cont_kind", "req_mask", "req")])
    w = _c(weights if weights is not None else [1] * 8, np.int64)
    if len(w) < 8:
        w = np.concatenate([w, np.ones(8 - len(w), np.int64)])
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    scores = np.zeros((P, pitch), np.int64)
    feas = np.zeros((P, pitch // 64), np.uint64)
    reasons = np.zeros((P, pitch), np.uint8)
    lib().orc_nrt_batch(C.byref(ns), C.c_int(N), C.byref(ps), C.c_int(P), C.c_int(strategy), _p(w), _p(fw),
                        C.c_int(words), _p(scores), _p(feas), _p(reasons), C.c_int(pitch))
    return scores, feas, reasons

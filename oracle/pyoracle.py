"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_alloc_score.restype = C.c_int64
        _lib.orc_tlp_score.restype = C.c_int64
        _lib.orc_tlp_score.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64]
        _lib.orc_lvrb_score.restype = C.c_int64
        _lib.orc_lvrb_score.argtypes = [C.c_double] * 4 + [C.c_int64, C.c_int64, C.c_uint8, C.c_int64, C.c_int64,
                                                           C.c_double, C.c_double]
        _lib.orc_lvrb_compute_score.restype = C.c_double
        _lib.orc_lvrb_compute_score.argtypes = [C.c_double] * 6
        _lib.orc_go_exp.restype = C.c_double
        _lib.orc_go_exp.argtypes = [C.c_double]
        _lib.orc_peaks_score.restype = C.c_int64
        _lib.orc_peaks_score.argtypes = [C.c_double, C.c_int64, C.c_uint8, C.c_double, C.c_double, C.c_int64]
        _lib.orc_incbet.restype = C.c_double
        _lib.orc_incbet.argtypes = [C.c_double] * 3
        _lib.orc_lowrisk_risk_load.restype = C.c_double
        _lib.orc_lowrisk_risk_load.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                               C.c_int64, C.c_int64]
        _lib.orc_lowrisk_compute_risk.restype = C.c_double
        _lib.orc_lowrisk_compute_risk.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_int64] * 6 + [C.c_double]
        _lib.orc_lowrisk_score.restype = C.c_int64
        _lib.orc_lowrisk_score.argtypes = [C.c_double] * 4 + [C.c_int64, C.c_int64, C.c_uint8] + [C.c_int64] * 9 + [C.c_double] * 2
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def alloc_score(alloc, weights, mode) -> int:
    a = _c(alloc, np.int64); w = _c(weights, np.int64)
    return int(lib().orc_alloc_score(_p(a), _p(w), C.c_int(len(a)), C.c_int(mode)))


def alloc_normalize(scores):
    s = _c(scores, np.int64).copy()
    lib().orc_alloc_normalize(_p(s), C.c_int(len(s)))
    return s


def alloc_batch(cols, weights, mode, P, feasible_words=None, pitch=None):
    cols = [_c(c, np.int64) for c in cols]
    N = len(cols[0])
    pitch = pitch or N
    w = _c(weights, np.int64)
    arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    out = np.zeros((P, pitch), dtype=np.int64)
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    lib().orc_alloc_batch(arr, C.c_int(len(cols)), C.c_int(N), _p(w), C.c_int(mode), C.c_int(P), _p(fw),
                          C.c_int(words), _p(out), C.c_int(pitch))
    return out


def gofaithful_alloc_batch(cols, res_names, weights, mode, pod_cpu_milli, pod_mem_bytes, feasible_words=None,
                           pitch=None, threads=1, return_seconds=False):
    """NodeResourcesAllocatable through the reference's per-call structure (oracle/gofaithful.cpp)."""
    cols = [_c(c, np.int64) for c in cols]
    N, P = len(cols[0]), len(pod_cpu_milli)
    pitch = pitch or N
    w = _c(weights, np.int64)
    arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    names = (C.c_char_p * len(cols))(*[n.encode() for n in res_names])
    pc = _c(pod_cpu_milli, np.int64); pm = _c(pod_mem_bytes, np.int64)
    out = np.zeros((P, pitch), dtype=np.int64)
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    secs = C.c_double()
    lib().orc_gofaithful_alloc_batch(arr, names, C.c_int(len(cols)), C.c_int(N), _p(w), C.c_int(mode), C.c_int(P),
                                     _p(pc), _p(pm), _p(fw), C.c_int(words), _p(out), C.c_int(pitch), C.c_int(threads),
                                     C.byref(secs))
    return (out, secs.value) if return_seconds else out


def gofaithful_alloc_cycles16(cols, res_names, weights, mode, pod_cpu_milli, pod_mem_bytes, feasible_words=None,
                              pitch=None, workers=16):
    """The same restatement in upstream's shape: pods one at a time, each cycle's Score calls over a `workers`-wide
    Parallelizer (chunks of chunkSizeFor(n, workers) nodes), NormalizeScore serial.  Returns (scores, seconds per cycle)."""
    cols = [_c(c, np.int64) for c in cols]
    N, P = len(cols[0]), len(pod_cpu_milli)
    pitch = pitch or N
    w = _c(weights, np.int64)
    arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
    names = (C.c_char_p * len(cols))(*[n.encode() for n in res_names])
    pc = _c(pod_cpu_milli, np.int64); pm = _c(pod_mem_bytes, np.int64)
    out = np.zeros((P, pitch), dtype=np.int64)
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    secs = np.zeros(P, dtype=np.float64)
    lib().orc_gofaithful_alloc_cycles16(arr, names, C.c_int(len(cols)), C.c_int(N), _p(w), C.c_int(mode), C.c_int(P),
                                        _p(pc), _p(pm), _p(fw), C.c_int(words), _p(out), C.c_int(pitch), C.c_int(workers),
                                        _p(secs))
    return out, secs


def tlp_score(util, cap, missing, flags, pod_cpu, target=40) -> int:
    return int(lib().orc_tlp_score(util, cap, missing, flags, pod_cpu, target))


def tlp_batch(util, cap, missing, flags, pod_cpu, target=40, pitch=None):
    util = _c(util, np.float64); cap = _c(cap, np.int64); missing = _c(missing, np.int64)
    flags = _c(flags, np.uint8); pod_cpu = _c(pod_cpu, np.int64)
    N, P = len(util), len(pod_cpu)
    pitch = pitch or N
    out = np.zeros((P, pitch), dtype=np.int64)
    lib().orc_tlp_batch(_p(util), _p(cap), _p(missing), _p(flags), C.c_int(N), _p(pod_cpu), C.c_int(P),
                        C.c_int64(target), _p(out), C.c_int(pitch))
    return out


def lvrb_compute_score(used_avg, used_std, req, capacity, margin, sensitivity) -> float:
    return float(lib().orc_lvrb_compute_score(used_avg, used_std, req, capacity, margin, sensitivity))


def lvrb_mu_sigma(used_avg, used_std, req, capacity):
    mu, sg = C.c_double(), C.c_double()
    lib().orc_lvrb_mu_sigma(C.c_double(used_avg), C.c_double(used_std), C.c_double(req), C.c_double(capacity),
                            C.byref(mu), C.byref(sg))
    return mu.value, sg.value


def lvrb_score(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, req_cpu, req_mem, margin=1.0,
               sens=1.0) -> int:
    return int(lib().orc_lvrb_score(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, req_cpu,
                                    req_mem, margin, sens))


def lvrb_batch(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, req_cpu, req_mem, margin=1.0,
               sens=1.0, pitch=None):
    f = [_c(x, np.float64) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
    i = [_c(x, np.int64) for x in (alloc_cpu, alloc_mem)]
    fl = _c(flags, np.uint8); rc = _c(req_cpu, np.int64); rm = _c(req_mem, np.int64)
    N, P = len(fl), len(rc)
    pitch = pitch or N
    out = np.zeros((P, pitch), dtype=np.int64)
    lib().orc_lvrb_batch(*[_p(x) for x in f], *[_p(x) for x in i], _p(fl), C.c_int(N), _p(rc), _p(rm), C.c_int(P),
                         C.c_double(margin), C.c_double(sens), _p(out), C.c_int(pitch))
    return out


# ---- Trimaran Peaks + LowRiskOverCommitment -----------------------------------------------------
def go_exp(x) -> float:
    return float(lib().orc_go_exp(x))


def peaks_score(util, cap, flags, k1, k2, pod_cpu) -> int:
    return int(lib().orc_peaks_score(util, cap, flags, k1, k2, pod_cpu))


def peaks_normalize(scores):
    s = _c(scores, np.int64).copy()
    lib().orc_peaks_normalize(_p(s), C.c_int(len(s)))
    return s


def peaks_batch(util, cap, flags, k1, k2, pod_cpu, feasible_words=None, pitch=None):
    util = _c(util, np.float64); cap = _c(cap, np.int64); flags = _c(flags, np.uint8)
    k1 = _c(k1, np.float64); k2 = _c(k2, np.float64); pod_cpu = _c(pod_cpu, np.int64)
    N, P = len(util), len(pod_cpu)
    pitch = pitch or N
    out = np.zeros((P, pitch), dtype=np.int64)
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    lib().orc_peaks_batch(_p(util), _p(cap), _p(flags), _p(k1), _p(k2), C.c_int(N), _p(pod_cpu), C.c_int(P), _p(fw),
                          C.c_int(words), _p(out), C.c_int(pitch))
    return out


def incbet(a, b, x) -> float:
    return float(lib().orc_incbet(a, b, x))


def lowrisk_risk_load(stats_ok, util, std, capacity_f, capacity, req_minus_pod, lim_minus_pod, window=5) -> float:
    return float(lib().orc_lowrisk_risk_load(int(stats_ok), util, std, capacity_f, capacity, req_minus_pod,
                                             lim_minus_pod, window))


def lowrisk_compute_risk(stats_ok, util, std, capacity_f, capacity, node_req, node_lim, pod_req, pod_lim, window=5,
                         weight=0.5) -> float:
    return float(lib().orc_lowrisk_compute_risk(int(stats_ok), util, std, capacity_f, capacity, node_req, node_lim,
                                                pod_req, pod_lim, window, weight))


def lowrisk_score(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, node_req_cpu, node_req_mem,
                  node_lim_cpu, node_lim_mem, pod_req_cpu, pod_req_mem, pod_lim_cpu, pod_lim_mem, window=5, w_cpu=0.5,
                  w_mem=0.5) -> int:
    return int(lib().orc_lowrisk_score(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, node_req_cpu,
                                       node_req_mem, node_lim_cpu, node_lim_mem, pod_req_cpu, pod_req_mem,
                                       pod_lim_cpu, pod_lim_mem, window, w_cpu, w_mem))


def lowrisk_batch(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, flags, node_req_cpu, node_req_mem,
                  node_lim_cpu, node_lim_mem, pod_req_cpu, pod_req_mem, pod_lim_cpu, pod_lim_mem, window=5, w_cpu=0.5,
                  w_mem=0.5, pitch=None):
    f = [_c(x, np.float64) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
    i = [_c(x, np.int64) for x in (alloc_cpu, alloc_mem)]
    fl = _c(flags, np.uint8)
    nd = [_c(x, np.int64) for x in (node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem)]
    pd = [_c(x, np.int64) for x in (pod_req_cpu, pod_req_mem, pod_lim_cpu, pod_lim_mem)]
    N, P = len(fl), len(pd[0])
    pitch = pitch or N
    out = np.zeros((P, pitch), dtype=np.int64)
    lib().orc_lowrisk_batch(*[_p(x) for x in f], *[_p(x) for x in i], _p(fl), *[_p(x) for x in nd], C.c_int(N),
                            *[_p(x) for x in pd], C.c_int(P), C.c_int64(window), C.c_double(w_cpu), C.c_double(w_mem),
                            _p(out), C.c_int(pitch))
    return out


# ---- NetworkOverhead -------------------------------------------------------------------------
NETOH_DEP_DTYPE = np.dtype([("host_node", "<i4"), ("host_region", "<u2"), ("host_zone", "<u2"),
                            ("max_network_cost", "<i8")])
NETOH_MISSING = -(2**63)


class _NetohTopo(C.Structure):
    _fields_ = [("n_names", C.c_int), ("zone_cost", C.c_void_p), ("region_cost", C.c_void_p)]


def _topo(zone_cost, region_cost):
    zc = _c(zone_cost, np.int64); rc = _c(region_cost, np.int64)
    return _NetohTopo(zc.shape[0], zc.ctypes.data, rc.ctypes.data), (zc, rc)


def netoh_node(zone_cost, region_cost, node_global, region, zone, deps):
    t, keep = _topo(zone_cost, region_cost)
    d = np.ascontiguousarray(deps, dtype=NETOH_DEP_DTYPE)
    s, v, c = C.c_int64(), C.c_int64(), C.c_int64()
    lib().orc_netoh_node(C.byref(t), C.c_int(node_global), C.c_int(region), C.c_int(zone), _p(d) if len(d) else None,
                         C.c_int(len(d)), C.byref(s), C.byref(v), C.byref(c))
    return s.value, v.value, c.value


def netoh_normalize(scores):
    s = _c(scores, np.int64).copy()
    lib().orc_netoh_normalize(_p(s), C.c_int(len(s)))
    return s


def netoh_batch(zone_cost, region_cost, region_id, zone_id, score_equally, dep_offset, deps, feasible_words=None,
                pitch=None, node_offset=0):
    t, keep = _topo(zone_cost, region_cost)
    rid = _c(region_id, np.uint16); zid = _c(zone_id, np.uint16)
    eq = _c(score_equally, np.uint8); off = _c(dep_offset, np.int32)
    d = np.ascontiguousarray(deps, dtype=NETOH_DEP_DTYPE)
    N, P = len(rid), len(eq)
    pitch = pitch or (N + 127) // 128 * 128
    fw = None if feasible_words is None else _c(feasible_words, np.uint64)
    words = 0 if fw is None else fw.shape[1]
    scores = np.zeros((P, pitch), dtype=np.int64)
    feas = np.zeros((P, pitch // 64), dtype=np.uint64)
    reasons = np.zeros((P, pitch), dtype=np.uint8)
    lib().orc_netoh_batch(C.byref(t), _p(rid), _p(zid), C.c_int(N), C.c_int(node_offset), _p(eq), _p(off),
                          _p(d) if len(d) else None, C.c_int(P), _p(fw), C.c_int(words), _p(scores), _p(feas),
                          _p(reasons), C.c_int(pitch))
    return scores, feas, reasons

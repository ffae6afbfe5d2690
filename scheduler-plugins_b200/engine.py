"""ctypes binding over the C-ABI of libb200sched.so (include/b200sched.h).

This is harness glue — the role the cgo stub plays for the Go plugins.
It owns no arithmetic: every score comes from the CUDA kernels behind the C-ABI, and the
module raises at load time when the library is missing (no CPU fallback exists).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200S_LIB selects another build of the same library (A/B experiments with tuning knobs); never a fallback
LIB_PATH = os.environ.get("B200S_LIB") or os.path.join(_HERE, "lib", "libb200sched.so")

(PLUGIN_ALLOCATABLE, PLUGIN_TLP, PLUGIN_LVRB, PLUGIN_NRT, PLUGIN_NETWORK_OVERHEAD, PLUGIN_PEAKS,
 PLUGIN_LOW_RISK) = range(7)
PLUGIN_COUNT = 7
PLUGIN_NAMES = ["NodeResourcesAllocatable", "TargetLoadPacking", "LoadVariationRiskBalancing", "NodeResourceTopologyMatch",
                "NetworkOverhead", "Peaks", "LowRiskOverCommitment"]
OUT_I64, OUT_U8 = 0, 1
ALLOC_LEAST, ALLOC_MOST = 0, 1
NRT_MOST_ALLOCATED, NRT_BALANCED_ALLOCATION, NRT_LEAST_ALLOCATED, NRT_LEAST_NUMA_NODES = range(4)
NODE_ALIGN = 128
NRT_MAX_ZONES = NRT_MAX_RES = NRT_MAX_CONT = 8
NETOH_MISSING = -(2**63)
NRT_PATH_AUTO, NRT_PATH_DIRECT, NRT_PATH_BATCHED = 0, 1, 2
PHASE_ALLREDUCE, PHASE_ALLGATHER, PHASE_COMBINE = 0, 1, 2
OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_NCCL, ERR_NOMEM = 0, -1, -2, -3, -4, -5, -6

EXPORTS = [
    "b200s_version", "b200s_init", "b200s_shutdown", "b200s_last_error", "b200s_stream", "b200s_sync",
    "b200s_launch_count", "b200s_comm_unique_id", "b200s_comm_init", "b200s_comm_rank", "b200s_comm_world",
    "b200s_snapshot_begin", "b200s_snapshot_allocatable", "b200s_snapshot_tlp", "b200s_snapshot_lvrb",
    "b200s_snapshot_nrt", "b200s_snapshot_network_overhead", "b200s_snapshot_commit",
    "b200s_snapshot_peaks", "b200s_snapshot_low_risk", "b200s_config_low_risk",
    "b200s_snapshot_patch_begin", "b200s_snapshot_patch_allocatable", "b200s_snapshot_patch_tlp",
    "b200s_snapshot_patch_lvrb", "b200s_snapshot_patch_nrt", "b200s_snapshot_patch_network_overhead",
    "b200s_snapshot_patch_peaks", "b200s_snapshot_patch_low_risk", "b200s_snapshot_patch_nrt_deduct",
    "b200s_config_allocatable", "b200s_config_tlp", "b200s_config_lvrb", "b200s_config_nrt",
    "b200s_config_network_overhead", "b200s_fetch_network_overhead_raw", "b200s_fetch_network_overhead_counts",
    "b200s_pods_upload", "b200s_eval", "b200s_fetch_scores", "b200s_fetch_feasible", "b200s_fetch_reasons",
    "b200s_device_scores", "b200s_device_feasible", "b200s_eval_combined", "b200s_fetch_topk",
    "b200s_fetch_total", "b200s_fetch_total_feasible", "b200s_score_batch", "b200s_alloc_pinned",
    "b200s_free_pinned", "b200s_npad", "b200s_set_profiling", "b200s_kernel_time", "b200s_debug_div_check",
    "b200s_config_nrt_path", "b200s_nrt_last_path", "b200s_nrt_path_note", "b200s_phase_time", "b200s_comm_peer_export", "b200s_comm_peer_import", "b200s_config_fused_cycle", "b200s_config_async_upload",
    "b200s_schedule_batch", "b200s_schedule_sequence",
]


class B200SError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200sched error {code}: {msg}")
        self.code = code


class NrtNodes(C.Structure):
    _fields_ = [("n_zones", C.c_int32), ("n_res", C.c_int32), ("res_flags", C.c_void_p),
                ("node_flags", C.c_void_p), ("max_numa", C.c_void_p), ("n_zones_node", C.c_void_p),
                ("node_res_mask", C.c_void_p), ("zone_res_mask", C.c_void_p), ("avail", C.c_void_p),
                ("cost", C.c_void_p)]


class NrtPods(C.Structure):
    _fields_ = [("qos", C.c_void_p), ("flags", C.c_void_p), ("n_init", C.c_void_p), ("n_app", C.c_void_p),
                ("cont_kind", C.c_void_p), ("req_mask", C.c_void_p), ("req", C.c_void_p)]


class NetohPods(C.Structure):
    _fields_ = [("score_equally", C.c_void_p), ("dep_offset", C.c_void_p), ("deps", C.c_void_p)]


class PodBatch(C.Structure):
    _fields_ = [("n_pods", C.c_int32), ("feasible", C.c_void_p), ("tlp_pod_cpu_milli", C.c_void_p),
                ("lvrb_req_cpu_milli", C.c_void_p), ("lvrb_req_mem_bytes", C.c_void_p),
                ("nrt", C.POINTER(NrtPods)), ("netoh", C.POINTER(NetohPods)),
                ("peaks_pod_cpu_milli", C.c_void_p), ("low_risk_pod", C.c_void_p)]


NETOH_DEP_DTYPE = np.dtype([("host_node", "<i4"), ("host_region", "<u2"), ("host_zone", "<u2"),
                            ("max_network_cost", "<i8")])
TOPK_DTYPE = np.dtype([("score", "<i8"), ("node", "<i4"), ("pad", "<i4")])


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the engine has no CPU fallback)")
    lib = C.CDLL(path)
    lib.b200s_last_error.restype = C.c_char_p
    lib.b200s_nrt_path_note.restype = C.c_char_p
    lib.b200s_stream.restype = C.c_void_p
    lib.b200s_launch_count.restype = C.c_uint64
    lib.b200s_device_scores.restype = C.c_void_p
    lib.b200s_device_feasible.restype = C.c_void_p
    lib.b200s_alloc_pinned.restype = C.c_void_p
    lib.b200s_alloc_pinned.argtypes = [C.c_size_t]
    lib.b200s_free_pinned.argtypes = [C.c_void_p]
    lib.b200s_shutdown.restype = None
    lib.b200s_free_pinned.restype = None
    return lib


def npad_of(n: int) -> int:
    return max(NODE_ALIGN, (n + NODE_ALIGN - 1) // NODE_ALIGN * NODE_ALIGN)


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _arr(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class PinnedBuffer:
    """Caller-owned pinned host memory from b200s_alloc_pinned, viewed as a numpy array."""

    def __init__(self, lib, nbytes):
        self._lib = lib
        self.nbytes = int(nbytes)
        self.ptr = lib.b200s_alloc_pinned(C.c_size_t(max(self.nbytes, 1)))
        if not self.ptr:
            raise MemoryError("b200s_alloc_pinned failed")

    def view(self, dtype, shape):
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self._lib.b200s_free_pinned(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    """One engine context = one GPU = one shard of the node axis."""

    def __init__(self, device: int = 0, lib: C.CDLL | None = None):
        self.lib = lib or load_library()
        self.ctx = C.c_void_p()
        rc = self.lib.b200s_init(C.c_int(device), C.byref(self.ctx))
        if rc != OK:
            msg = self.lib.b200s_last_error(None)
            raise B200SError(rc, msg.decode() if msg else "b200s_init failed")
        self.N = self.Npad = self.P = 0
        self._keep = []

    # -- plumbing -------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != OK:
            raise B200SError(rc, self.lib.b200s_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.lib.b200s_shutdown(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._chk(self.lib.b200s_sync(self.ctx))

    @property
    def stream(self) -> int:
        return int(self.lib.b200s_stream(self.ctx) or 0)

    @property
    def launches(self) -> int:
        return int(self.lib.b200s_launch_count(self.ctx))

    def set_profiling(self, on: bool):
        self._chk(self.lib.b200s_set_profiling(self.ctx, C.c_int(1 if on else 0)))

    def kernel_time(self, plugin):
        """(summed ms, launches) of the plugin's dominant kernel since the last call."""
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self.lib.b200s_kernel_time(self.ctx, C.c_int(plugin), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def phase_time(self, phase):
        """(summed ms, count) of PHASE_ALLREDUCE / PHASE_ALLGATHER / PHASE_COMBINE since the last call."""
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self.lib.b200s_phase_time(self.ctx, C.c_int(phase), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def debug_div_check(self, x, d) -> int:
        x = _arr(x, np.float64); d = _arr(d, np.float64)
        out = C.c_uint64()
        self._chk(self.lib.b200s_debug_div_check(self.ctx, _ptr(x), _ptr(d), C.c_int32(len(x)), C.byref(out)))
        return int(out.value)

    def pinned(self, nbytes) -> PinnedBuffer:
        return PinnedBuffer(self.lib, nbytes)

    # -- multi-GPU ------------------------------------------------------------------------
    def peer_export(self) -> bytes:
        """CUDA IPC handle of this rank's symmetric exchange buffer (after comm_init)."""
        buf = C.create_string_buffer(64)
        self._chk(self.lib.b200s_comm_peer_export(self.ctx, buf))
        return buf.raw

    def peer_import(self, handles):
        """handles: list of every rank's peer_export() bytes, in rank order; switches the per-pod exchanges from NCCL
        to stores into peer memory over NVLink."""
        blob = b"".join(handles)
        self._chk(self.lib.b200s_comm_peer_import(self.ctx, C.c_char_p(blob)))

    def unique_id(self) -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = self.lib.b200s_comm_unique_id(buf)
        if rc != OK:
            raise B200SError(rc, "ncclGetUniqueId failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def comm_init(self, uid: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._chk(self.lib.b200s_comm_init(self.ctx, buf, C.c_int(rank), C.c_int(world)))

    # -- snapshot -------------------------------------------------------------------------
    def snapshot_begin(self, n_nodes, generation=1, node_offset=0, n_nodes_global=None):
        if n_nodes_global is None:
            n_nodes_global = node_offset + n_nodes
        self._chk(self.lib.b200s_snapshot_begin(self.ctx, C.c_uint64(generation), C.c_int32(n_nodes),
                                                C.c_int32(node_offset), C.c_int32(n_nodes_global)))
        self.N = n_nodes
        self.Npad = int(self.lib.b200s_npad(self.ctx))

    def snapshot_allocatable(self, cols):
        cols = [_arr(c, np.int64, (self.N,)) for c in cols]
        arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        self._chk(self.lib.b200s_snapshot_allocatable(self.ctx, C.c_int32(len(cols)), arr))

    def snapshot_tlp(self, cpu_util_pct, cap_milli, missing_milli, flags):
        a = _arr(cpu_util_pct, np.float64, (self.N,)); b = _arr(cap_milli, np.int64, (self.N,))
        c = _arr(missing_milli, np.int64, (self.N,)); d = _arr(flags, np.uint8, (self.N,))
        self._chk(self.lib.b200s_snapshot_tlp(self.ctx, _ptr(a), _ptr(b), _ptr(c), _ptr(d)))

    def snapshot_lvrb(self, cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu_milli, alloc_mem_bytes, flags):
        f = [_arr(x, np.float64, (self.N,)) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
        i = [_arr(x, np.int64, (self.N,)) for x in (alloc_cpu_milli, alloc_mem_bytes)]
        fl = _arr(flags, np.uint8, (self.N,))
        self._chk(self.lib.b200s_snapshot_lvrb(self.ctx, *[_ptr(x) for x in f], *[_ptr(x) for x in i], _ptr(fl)))

    def snapshot_nrt(self, nodes: dict):
        """nodes: dict with n_zones, n_res, res_flags[R], node_flags[N], max_numa[N], n_zones_node[N],
        node_res_mask[N], zone_res_mask[Z][N], avail[Z][R][N], cost[Z][Z][N] or None."""
        Z, R, N = int(nodes["n_zones"]), int(nodes["n_res"]), self.N
        keep = dict(
            res_flags=_arr(nodes["res_flags"], np.uint8, (R,)),
            node_flags=_arr(nodes["node_flags"], np.uint8, (N,)),
            max_numa=_arr(nodes["max_numa"], np.uint16, (N,)),
            n_zones_node=_arr(nodes["n_zones_node"], np.uint8, (N,)),
            node_res_mask=_arr(nodes["node_res_mask"], np.uint8, (N,)),
            zone_res_mask=_arr(nodes["zone_res_mask"], np.uint8, (Z, N)),
            avail=_arr(nodes["avail"], np.int64, (Z, R, N)),
            cost=None if nodes.get("cost") is None else _arr(nodes["cost"], np.int32, (Z, Z, N)),
        )
        s = NrtNodes(Z, R, *[_ptr(keep[k]) for k in ("res_flags", "node_flags", "max_numa", "n_zones_node",
                                                     "node_res_mask", "zone_res_mask", "avail", "cost")])
        self._chk(self.lib.b200s_snapshot_nrt(self.ctx, C.byref(s)))
        self.nrt_R = R

    def snapshot_network_overhead(self, region_id, zone_id, zone_cost, region_cost):
        K = int(np.asarray(zone_cost).shape[0])
        a = _arr(region_id, np.uint16, (self.N,)); b = _arr(zone_id, np.uint16, (self.N,))
        zc = _arr(zone_cost, np.int64, (K, K)); rc_ = _arr(region_cost, np.int64, (K, K))
        self._chk(self.lib.b200s_snapshot_network_overhead(self.ctx, _ptr(a), _ptr(b), C.c_int32(K), _ptr(zc),
                                                           _ptr(rc_)))

    def snapshot_peaks(self, cpu_util_pct, cap_milli, flags, k1, k2):
        n = (self.N,)
        a = _arr(cpu_util_pct, np.float64, n); b = _arr(cap_milli, np.int64, n); c = _arr(flags, np.uint8, n)
        d = _arr(k1, np.float64, n); e = _arr(k2, np.float64, n)
        self._chk(self.lib.b200s_snapshot_peaks(self.ctx, _ptr(a), _ptr(b), _ptr(c), _ptr(d), _ptr(e)))

    def snapshot_low_risk(self, cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu_milli, alloc_mem_bytes, flags,
                          node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem):
        n = (self.N,)
        f = [_arr(x, np.float64, n) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
        i = [_arr(x, np.int64, n) for x in (alloc_cpu_milli, alloc_mem_bytes)]
        fl = _arr(flags, np.uint8, n)
        nd = [_arr(x, np.int64, n) for x in (node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem)]
        self._chk(self.lib.b200s_snapshot_low_risk(self.ctx, *[_ptr(x) for x in f], *[_ptr(x) for x in i], _ptr(fl),
                                                   *[_ptr(x) for x in nd]))

    def config_low_risk(self, smoothing_window_size=5, w_cpu=0.5, w_mem=0.5):
        self._chk(self.lib.b200s_config_low_risk(self.ctx, C.c_int64(smoothing_window_size), C.c_double(w_cpu),
                                                 C.c_double(w_mem)))

    def snapshot_commit(self):
        self._chk(self.lib.b200s_snapshot_commit(self.ctx))

    # -- incremental snapshot: rewrite a few node rows in place (same node list) -------------
    def snapshot_patch_begin(self, generation=0):
        self._chk(self.lib.b200s_snapshot_patch_begin(self.ctx, C.c_uint64(generation)))

    def _idx(self, node_idx):
        return _arr(node_idx, np.int32)

    def snapshot_patch_allocatable(self, node_idx, cols):
        idx = self._idx(node_idx)
        cols = [_arr(c, np.int64, (len(idx),)) for c in cols]
        arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        self._chk(self.lib.b200s_snapshot_patch_allocatable(self.ctx, C.c_int32(len(idx)), _ptr(idx),
                                                            C.c_int32(len(cols)), arr))

    def snapshot_patch_tlp(self, node_idx, cpu_util_pct, cap_milli, missing_milli, flags):
        idx = self._idx(node_idx); m = (len(idx),)
        a = _arr(cpu_util_pct, np.float64, m); b = _arr(cap_milli, np.int64, m)
        c = _arr(missing_milli, np.int64, m); d = _arr(flags, np.uint8, m)
        self._chk(self.lib.b200s_snapshot_patch_tlp(self.ctx, C.c_int32(len(idx)), _ptr(idx), _ptr(a), _ptr(b),
                                                    _ptr(c), _ptr(d)))

    def snapshot_patch_lvrb(self, node_idx, cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu_milli, alloc_mem_bytes,
                            flags):
        idx = self._idx(node_idx); m = (len(idx),)
        f = [_arr(x, np.float64, m) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
        i = [_arr(x, np.int64, m) for x in (alloc_cpu_milli, alloc_mem_bytes)]
        fl = _arr(flags, np.uint8, m)
        self._chk(self.lib.b200s_snapshot_patch_lvrb(self.ctx, C.c_int32(len(idx)), _ptr(idx),
                                                     *[_ptr(x) for x in f], *[_ptr(x) for x in i], _ptr(fl)))

    def snapshot_patch_peaks(self, node_idx, cpu_util_pct, cap_milli, flags, k1, k2):
        idx = self._idx(node_idx); m = (len(idx),)
        a = _arr(cpu_util_pct, np.float64, m); b = _arr(cap_milli, np.int64, m); c = _arr(flags, np.uint8, m)
        d = _arr(k1, np.float64, m); e = _arr(k2, np.float64, m)
        self._chk(self.lib.b200s_snapshot_patch_peaks(self.ctx, C.c_int32(len(idx)), _ptr(idx), _ptr(a), _ptr(b),
                                                      _ptr(c), _ptr(d), _ptr(e)))

    def snapshot_patch_low_risk(self, node_idx, cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu_milli, alloc_mem_bytes,
                                flags, node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem):
        idx = self._idx(node_idx); m = (len(idx),)
        f = [_arr(x, np.float64, m) for x in (cpu_avg, cpu_std, mem_avg, mem_std)]
        i = [_arr(x, np.int64, m) for x in (alloc_cpu_milli, alloc_mem_bytes)]
        fl = _arr(flags, np.uint8, m)
        nd = [_arr(x, np.int64, m) for x in (node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem)]
        self._chk(self.lib.b200s_snapshot_patch_low_risk(self.ctx, C.c_int32(len(idx)), _ptr(idx),
                                                         *[_ptr(x) for x in f], *[_ptr(x) for x in i], _ptr(fl),
                                                         *[_ptr(x) for x in nd]))

    def snapshot_patch_nrt(self, node_idx, rows: dict):
        """rows: the dict of snapshot_nrt with every [..][N] array cut down to [..][len(node_idx)]."""
        idx = self._idx(node_idx); m = len(idx)
        Z, R = int(rows["n_zones"]), int(rows["n_res"])
        keep = dict(
            node_flags=_arr(rows["node_flags"], np.uint8, (m,)),
            max_numa=_arr(rows["max_numa"], np.uint16, (m,)),
            n_zones_node=_arr(rows["n_zones_node"], np.uint8, (m,)),
            node_res_mask=_arr(rows["node_res_mask"], np.uint8, (m,)),
            zone_res_mask=_arr(rows["zone_res_mask"], np.uint8, (Z, m)),
            avail=_arr(rows["avail"], np.int64, (Z, R, m)),
            cost=None if rows.get("cost") is None else _arr(rows["cost"], np.int32, (Z, Z, m)),
        )
        s = NrtNodes(Z, R, None, *[_ptr(keep[k]) for k in ("node_flags", "max_numa", "n_zones_node",
                                                           "node_res_mask", "zone_res_mask", "avail", "cost")])
        self._chk(self.lib.b200s_snapshot_patch_nrt(self.ctx, C.c_int32(m), _ptr(idx), C.byref(s)))

    def snapshot_patch_nrt_deduct(self, node_idx, res_mask, deduct):
        """OverReserve: deduct[R][count] off every zone of the listed nodes that reports the resource."""
        idx = self._idx(node_idx); m = len(idx)
        rm = _arr(res_mask, np.uint8, (m,)); d = _arr(deduct, np.int64, (self.nrt_R, m))
        self._chk(self.lib.b200s_snapshot_patch_nrt_deduct(self.ctx, C.c_int32(m), _ptr(idx), _ptr(rm), _ptr(d)))

    def snapshot_patch_network_overhead(self, node_idx, region_id, zone_id):
        idx = self._idx(node_idx); m = (len(idx),)
        a = _arr(region_id, np.uint16, m); b = _arr(zone_id, np.uint16, m)
        self._chk(self.lib.b200s_snapshot_patch_network_overhead(self.ctx, C.c_int32(len(idx)), _ptr(idx), _ptr(a),
                                                                 _ptr(b)))

    # -- plugin args ----------------------------------------------------------------------
    def config_allocatable(self, mode, weights):
        w = _arr(weights, np.int64)
        self._chk(self.lib.b200s_config_allocatable(self.ctx, C.c_int(mode), C.c_int32(len(w)), _ptr(w)))

    def config_tlp(self, target_utilization_pct=40):
        self._chk(self.lib.b200s_config_tlp(self.ctx, C.c_int64(target_utilization_pct)))

    def config_lvrb(self, margin=1.0, sensitivity=1.0):
        self._chk(self.lib.b200s_config_lvrb(self.ctx, C.c_double(margin), C.c_double(sensitivity)))

    def config_nrt(self, strategy, weights=None):
        w = _arr(weights if weights is not None else [], np.int64)
        self._chk(self.lib.b200s_config_nrt(self.ctx, C.c_int(strategy), C.c_int32(len(w)),
                                            _ptr(w) if len(w) else None))

    def config_nrt_path(self, path):
        """NRT_PATH_AUTO / NRT_PATH_DIRECT / NRT_PATH_BATCHED (include/b200sched.h)."""
        self._chk(self.lib.b200s_config_nrt_path(self.ctx, C.c_int(path)))

    def nrt_last_path(self):
        return int(self.lib.b200s_nrt_last_path(self.ctx))

    def nrt_path_note(self):
        return (self.lib.b200s_nrt_path_note(self.ctx) or b"").decode()

    # -- pods -----------------------------------------------------------------------------
    def make_batch(self, n_pods, feasible=None, tlp_pod_cpu_milli=None, lvrb_req_cpu_milli=None,
                   lvrb_req_mem_bytes=None, nrt=None, netoh=None, peaks_pod_cpu_milli=None, low_risk_pod=None):
        """Builds the b200s_pod_batch struct; returns (struct, keepalive list)."""
        P = int(n_pods)
        keep = []
        words = self.Npad // 64

        def col(a, dtype, shape):
            if a is None:
                return None
            if isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and tuple(a.shape) == shape:
                arr = a
            else:
                arr = _arr(a, dtype, shape)
            keep.append(arr)
            return C.c_void_p(arr.ctypes.data)

        b = PodBatch()
        b.n_pods = P
        b.feasible = col(feasible, np.uint64, (P, words))
        b.tlp_pod_cpu_milli = col(tlp_pod_cpu_milli, np.int64, (P,))
        b.lvrb_req_cpu_milli = col(lvrb_req_cpu_milli, np.int64, (P,))
        b.lvrb_req_mem_bytes = col(lvrb_req_mem_bytes, np.int64, (P,))
        b.peaks_pod_cpu_milli = col(peaks_pod_cpu_milli, np.int64, (P,))
        b.low_risk_pod = col(low_risk_pod, np.int64, (4, P))  # req cpu, req mem, limit cpu, limit mem
        if nrt is not None:
            Cn, R = NRT_MAX_CONT, self.nrt_R
            s = NrtPods(col(nrt["qos"], np.uint8, (P,)), col(nrt["flags"], np.uint8, (P,)),
                        col(nrt["n_init"], np.uint8, (P,)), col(nrt["n_app"], np.uint8, (P,)),
                        col(nrt["cont_kind"], np.uint8, (P, Cn)), col(nrt["req_mask"], np.uint8, (P, Cn + 1)),
                        col(nrt["req"], np.int64, (P, Cn + 1, R)))
            keep.append(s)
            b.nrt = C.pointer(s)
        if netoh is not None:
            off = _arr(netoh["dep_offset"], np.int32, (P + 1,))
            deps = np.ascontiguousarray(netoh["deps"], dtype=NETOH_DEP_DTYPE)
            s = NetohPods(col(netoh["score_equally"], np.uint8, (P,)), col(off, np.int32, (P + 1,)),
                          col(deps, NETOH_DEP_DTYPE, deps.shape) if len(deps) else None)
            keep.append(s)
            b.netoh = C.pointer(s)
        return b, keep

    def pods_upload(self, n_pods, **cols):
        b, keep = self.make_batch(n_pods, **cols)
        self._chk(self.lib.b200s_pods_upload(self.ctx, C.byref(b)))
        self.P = int(n_pods)

    # -- evaluation -----------------------------------------------------------------------
    def eval(self, plugin, dtype=OUT_I64):
        self._chk(self.lib.b200s_eval(self.ctx, C.c_int(plugin), C.c_int(dtype)))
        self._last_dtype = dtype

    def fetch_scores(self, plugin, dtype=OUT_I64, out=None):
        npdt = np.int64 if dtype == OUT_I64 else np.uint8
        if out is None:
            out = np.empty((self.P, self.Npad), dtype=npdt)
        self._chk(self.lib.b200s_fetch_scores(self.ctx, C.c_int(plugin), _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def fetch_feasible(self, plugin):
        out = np.empty((self.P, self.Npad // 64), dtype=np.uint64)
        self._chk(self.lib.b200s_fetch_feasible(self.ctx, C.c_int(plugin), _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def fetch_reasons(self, plugin):
        out = np.empty((self.P, self.Npad), dtype=np.uint8)
        self._chk(self.lib.b200s_fetch_reasons(self.ctx, C.c_int(plugin), _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def config_network_overhead(self, want_counts=True, apply_own_filter=True):
        self._chk(self.lib.b200s_config_network_overhead(self.ctx, C.c_int(1 if want_counts else 0),
                                                         C.c_int(1 if apply_own_filter else 0)))

    def fetch_network_overhead_raw(self):
        out = np.empty((self.P, self.Npad), dtype=np.int64)
        self._chk(self.lib.b200s_fetch_network_overhead_raw(self.ctx, _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def fetch_network_overhead_counts(self):
        out = np.empty((self.P, self.Npad), dtype=np.uint32)
        self._chk(self.lib.b200s_fetch_network_overhead_counts(self.ctx, _ptr(out), C.c_size_t(out.nbytes)))
        return out & 0xffff, out >> 16

    def device_scores(self, plugin) -> int:
        return int(self.lib.b200s_device_scores(self.ctx, C.c_int(plugin)) or 0)

    def eval_combined(self, plugin_mask, weights, k=1, write_total=False):
        w = np.zeros(PLUGIN_COUNT, dtype=np.int64)  # indexed by plugin id; shorter lists cover the first plugins
        w[:len(weights)] = np.asarray(weights, dtype=np.int64)
        self._chk(self.lib.b200s_eval_combined(self.ctx, C.c_uint32(plugin_mask), _ptr(w), C.c_int32(k),
                                               C.c_int(1 if write_total else 0)))
        self._k = k

    def config_async_upload(self, on: bool):
        """pods_upload queues and returns (keep the batch's arrays alive and unchanged until the next fetch)"""
        self._chk(self.lib.b200s_config_async_upload(self.ctx, C.c_int(1 if on else 0)))

    def config_fused_cycle(self, on):
        """False / 0: plugin-by-plugin for every batch size; True / 1: fused cycle, b200s_schedule_batch as one graph
        launch; 2: fused cycle with plain launches (A/B timing)"""
        self._chk(self.lib.b200s_config_fused_cycle(self.ctx, C.c_int(2 if on == 2 else (1 if on else 0))))

    def schedule_batch(self, batch, plugin_mask, weights, k=1, out=None):
        """upload + eval_combined + winners to host in ONE call / ONE synchronisation; returns [P][k] TOPK_DTYPE"""
        w = np.zeros(PLUGIN_COUNT, dtype=np.int64)
        w[:len(weights)] = np.asarray(weights, dtype=np.int64)
        P = int(batch.n_pods)
        if out is None:
            out = np.empty((P, k), dtype=TOPK_DTYPE)
        self._chk(self.lib.b200s_schedule_batch(self.ctx, C.byref(batch), C.c_uint32(plugin_mask), _ptr(w), C.c_int32(k), _ptr(out)))
        self.P, self._k = P, k
        return out

    def prepare_schedule_batch(self, batch, plugin_mask, weights, k, out):
        """b200s_schedule_batch with every ctypes argument built once: returns a zero-argument callable whose cost is the
        C-ABI call itself (latency measurements; the cgo call of the Go shim has no marshalling either).  `batch`, `out`
        and the returned callable's captured arrays must stay alive and unchanged in layout."""
        w = np.zeros(PLUGIN_COUNT, dtype=np.int64)
        w[:len(weights)] = np.asarray(weights, dtype=np.int64)
        fn, ctx, bref = self.lib.b200s_schedule_batch, self.ctx, C.byref(batch)
        m, wp, kk, op = C.c_uint32(plugin_mask), _ptr(w), C.c_int32(k), _ptr(out)
        self.P, self._k = int(batch.n_pods), k

        def call(_keep=(w, batch, out)):
            rc = fn(ctx, bref, m, wp, kk, op)
            if rc != OK:
                self._chk(rc)

        return call

    def schedule_sequence(self, batch, plugin_mask, weights):
        """speculative placement of the batch pod by pod with the on-device assume; returns [P] TOPK_DTYPE winners"""
        w = np.zeros(PLUGIN_COUNT, dtype=np.int64)
        w[:len(weights)] = np.asarray(weights, dtype=np.int64)
        P = int(batch.n_pods)
        out = np.empty(P, dtype=TOPK_DTYPE)
        self._chk(self.lib.b200s_schedule_sequence(self.ctx, C.byref(batch), C.c_uint32(plugin_mask), _ptr(w), _ptr(out)))
        self.P = P
        return out

    def fetch_topk(self):
        out = np.empty((self.P, self._k), dtype=TOPK_DTYPE)
        self._chk(self.lib.b200s_fetch_topk(self.ctx, _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def fetch_total(self):
        out = np.empty((self.P, self.Npad), dtype=np.int64)
        self._chk(self.lib.b200s_fetch_total(self.ctx, _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def fetch_total_feasible(self):
        out = np.empty((self.P, self.Npad // 64), dtype=np.uint64)
        self._chk(self.lib.b200s_fetch_total_feasible(self.ctx, _ptr(out), C.c_size_t(out.nbytes)))
        return out

    def score_batch(self, plugin, batch: PodBatch, dtype, scores_out, feasible_out=None, reasons_out=None):
        """The per-cycle call of the Go shim: HOST buffers in, HOST buffers out."""
        self._chk(self.lib.b200s_score_batch(self.ctx, C.c_int(plugin), C.byref(batch), C.c_int(dtype),
                                             _ptr(scores_out), _ptr(feasible_out), _ptr(reasons_out)))
        self.P = int(batch.n_pods)


def pack_bits(mask_bool: np.ndarray, npad: int) -> np.ndarray:
    """[P][N] bool -> [P][npad/64] uint64 words (bit j of word w = node 64*w+j)."""
    P, N = mask_bool.shape
    padded = np.zeros((P, npad), dtype=np.uint8)
    padded[:, :N] = mask_bool
    return np.packbits(padded, axis=1, bitorder="little").view(np.uint64).reshape(P, npad // 64)


def unpack_bits(words: np.ndarray, n: int) -> np.ndarray:
    P = words.shape[0]
    return np.unpackbits(words.view(np.uint8).reshape(P, -1), axis=1, bitorder="little")[:, :n].astype(bool)

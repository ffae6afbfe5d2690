"""Oracle-side restatement of the reference's HOST-SIDE rules: API objects -> dense columns.

TEST INFRASTRUCTURE (see oracle/oracle.h): pure Python, small cases only.  It restates what the
Go plugins do to a v1.Pod / v1.Node / NodeResourceTopology before any per-node arithmetic, so that
the reference's object-level unit-test fixtures (tests/golden/*.json) can be replayed through
the C oracle — and, independently, through the product's C++ host layer + CUDA engine.

Objects are plain dicts:
  pod  = {"init": [cont...], "containers": [cont...], "overhead": {res: qty}}
  cont = {"requests": {res: qty}, "limits": {res: qty}, "restart_always": bool}
  node = {"name": str, "allocatable": {res: qty}}
  nrt  = {"policies": [str], "attributes": {name: value}, "zones": [zone...]}
  zone = {"name": "node-<id>", "type": "Node", "resources": {res: {"capacity","available"}},
          "costs": {"node-<id>": int}}
Quantities are Kubernetes quantity strings or ints; internally exact milli-units.
"""
from __future__ import annotations

import re
from fractions import Fraction

import numpy as np

Z_MAX = R_MAX = C_MAX = 8
QOS_GUARANTEED, QOS_BURSTABLE, QOS_BEST_EFFORT = 0, 1, 2
NODE_HAS_NRT, NODE_FRESH, NODE_SINGLE_NUMA, NODE_SCOPE_POD, NODE_UNSUPPORTED = 1, 2, 4, 8, 16
RES_AFFINE, RES_HOST_LEVEL = 1, 2
POD_FILTER_BYPASS, POD_UNSUPPORTED = 1, 2
CONT_APP, CONT_INIT, CONT_SIDECAR = 0, 1, 2

_BIN = {"Ki": 2**10, "Mi": 2**20, "Gi": 2**30, "Ti": 2**40, "Pi": 2**50, "Ei": 2**60}
_DEC = {"n": Fraction(1, 10**9), "u": Fraction(1, 10**6), "m": Fraction(1, 1000), "": 1, "k": 10**3, "M": 10**6,
        "G": 10**9, "T": 10**12, "P": 10**15, "E": 10**18}


def milli(q) -> int:
    """resource.Quantity -> exact milli-units [apimachinery resource.Quantity, upstream].
    MilliValue() rounds up; a quantity finer than 1m is not representable and raises."""
    if isinstance(q, (int, np.integer)):
        return int(q) * 1000
    s = str(q).strip()
    m = re.fullmatch(r"([+-]?\d+(?:\.\d*)?|\.\d+)(?:([eE][+-]?\d+)|(Ki|Mi|Gi|Ti|Pi|Ei|n|u|m|k|M|G|T|P|E)?)", s)
    if not m:
        raise ValueError(f"bad quantity {q!r}")
    num = Fraction(m.group(1))
    if m.group(2):
        num *= Fraction(10) ** int(m.group(2)[1:])
    else:
        suf = m.group(3) or ""
        num *= _BIN[suf] if suf in _BIN else _DEC[suf]
    mv = num * 1000
    if mv.denominator != 1:
        raise ValueError(f"quantity {q!r} is finer than one milli-unit (unsupported by the dense encoding)")
    return int(mv)


def value_of(mv: int) -> int:
    """Quantity.Value(): milli -> whole units, rounded up."""
    return -((-mv) // 1000)


# ---- resource name predicates [k8s.io/kubernetes pkg/apis/core/v1/helper, upstream] ------------
def is_native(name: str) -> bool:  # IsNativeResource
    return "/" not in name or "kubernetes.io/" in name


def is_hugepage(name: str) -> bool:
    return name.startswith("hugepages-")


def is_numa_affine(name: str) -> bool:  # numaresources.go:120-135
    return name in ("cpu", "memory") or is_hugepage(name)


def is_host_level(name: str) -> bool:  # numaresources.go:105-118
    return name in ("ephemeral-storage", "storage") or not is_native(name)


def is_scalar_resource_name(name: str) -> bool:  # schedutil.IsScalarResourceName [upstream]
    extended = (not is_native(name)) and not name.startswith("requests.")
    return extended or is_hugepage(name) or "kubernetes.io/" in name or name.startswith("attachable-volumes-")


def node_level_resources(node: dict) -> set:
    """util.ResourceList(nodeInfo.GetAllocatable()) keys (pkg/util/resource.go:28-44): cpu, memory,
    pods, ephemeral-storage always; plus the scalar resources NodeInfo keeps."""
    out = {"cpu", "memory", "pods", "ephemeral-storage"}
    for name in node.get("allocatable", {}):
        if is_scalar_resource_name(name):
            out.add(name)
    return out


# ---- pod rules -----------------------------------------------------------------------------------
def all_containers(pod):
    return list(pod.get("init", [])) + list(pod.get("containers", []))


def pod_qos(pod) -> int:
    """v1qos.GetPodQOS [upstream]: only cpu and memory count."""
    requests, limits = {}, {}
    guaranteed = True
    for c in list(pod.get("containers", [])) + list(pod.get("init", [])):
        for name, q in c.get("requests", {}).items():
            if name in ("cpu", "memory") and milli(q) > 0:
                requests[name] = requests.get(name, 0) + milli(q)
        found = set()
        for name, q in c.get("limits", {}).items():
            if name in ("cpu", "memory") and milli(q) > 0:
                found.add(name)
                limits[name] = limits.get(name, 0) + milli(q)
        if found != {"cpu", "memory"}:
            guaranteed = False
    if not requests and not limits:
        return QOS_BEST_EFFORT
    if guaranteed:
        for name, r in requests.items():
            if name not in limits or limits[name] != r:
                guaranteed = False
                break
    if guaranteed and len(requests) == len(limits):
        return QOS_GUARANTEED
    return QOS_BURSTABLE


def include_non_native(pod) -> bool:  # resourcerequests/exclusive.go:26-41
    return any(not is_native(r) for c in all_containers(pod) for r in c.get("requests", {}))


def pod_effective_request(pod) -> dict:
    """util.GetPodEffectiveRequest (pkg/util/resource.go:51-85), milli-units."""
    init, res = {}, {}
    for c in pod.get("init", []):
        for name, q in c.get("requests", {}).items():
            v = milli(q)
            if name in init and v <= init[name]:
                continue
            init[name] = v
    for c in pod.get("containers", []):
        for name, q in c.get("requests", {}).items():
            res[name] = res.get(name, 0) + milli(q)
    for name, v in init.items():
        if name in res and v <= res[name]:
            continue
        res[name] = v
    for name, q in (pod.get("overhead") or {}).items():
        res[name] = res.get(name, 0) + milli(q)
    return res


def tlp_pod_cpu(pod, default_requests_milli=1000, multiplier=1.5) -> int:
    """sum PredictUtilisation over app containers + overhead (targetloadpacking.go:122-129, 198-205)."""
    total = 0
    for c in pod.get("containers", []):
        if "cpu" in c.get("limits", {}):
            total += milli(c["limits"]["cpu"])
        elif "cpu" in c.get("requests", {}):
            x = float(milli(c["requests"]["cpu"])) * multiplier
            total += int(np.floor(abs(x) + 0.5) * np.sign(x))  # math.Round
        else:
            total += default_requests_milli
    if pod.get("overhead") is not None:
        total += milli(pod["overhead"].get("cpu", 0))
    return total


def lvrb_pod_request(pod):
    """trimaran.GetResourceRequested (resourcestats.go:110-146): (cpu milli, memory bytes)."""
    cpu = sum(milli(c.get("requests", {}).get("cpu", 0)) for c in pod.get("containers", []))
    mem = sum(value_of(milli(c.get("requests", {}).get("memory", 0))) for c in pod.get("containers", []))
    for c in pod.get("init", []):
        r = c.get("requests", {})
        if "cpu" in r:
            cpu = max(cpu, milli(r["cpu"]))
        if "memory" in r:
            mem = max(mem, value_of(milli(r["memory"])))
    oh = pod.get("overhead") or {}
    cpu += milli(oh.get("cpu", 0))
    mem += value_of(milli(oh.get("memory", 0)))
    return cpu, mem


def metrics_flatten(metrics):
    """load-watcher metric list of one node -> TLP and LVRB columns.
    TLP: LAST entry with Type CPU and operator Average|Latest (targetloadpacking.go:131-140).
    LVRB: GetResourceData (resourcestats.go:89-107) — Average wins, Latest/'' only before it."""
    out = dict(tlp_util=0.0, tlp_flags=0, cpu_avg=0.0, cpu_std=0.0, mem_avg=0.0, mem_std=0.0, lvrb_flags=0)
    if metrics is None:
        return out
    out["tlp_flags"] |= 1
    out["lvrb_flags"] |= 1
    for m in metrics:
        if m["type"] == "CPU" and m.get("operator", "") in ("AVG", "Latest"):
            out["tlp_util"] = float(m["value"])
            out["tlp_flags"] |= 2
    for typ, pre, bit in (("CPU", "cpu", 2), ("Memory", "mem", 4)):
        avg_found = False
        for m in metrics:
            if m["type"] != typ:
                continue
            op = m.get("operator", "")
            if op == "AVG":
                out[pre + "_avg"] = float(m["value"])
                avg_found = True
            elif op == "STD":
                out[pre + "_std"] = float(m["value"])
            elif op in ("", "Latest") and not avg_found:
                out[pre + "_avg"] = float(m["value"])
            out["lvrb_flags"] |= bit
    return out


# ---- NRT rules -----------------------------------------------------------------------------------
_POLICIES = {  # nodeconfig/topologymanager.go:130-161
    "SingleNUMANodePodLevel": ("single-numa-node", "pod"),
    "SingleNUMANodeContainerLevel": ("single-numa-node", "container"),
    "BestEffortPodLevel": ("best-effort", "pod"),
    "BestEffortContainerLevel": ("best-effort", "container"),
    "RestrictedPodLevel": ("restricted", "pod"),
    "RestrictedContainerLevel": ("restricted", "container"),
}


def topology_manager(nrt) -> tuple:
    """TopologyManagerFromNodeResourceTopology (nodeconfig/topologymanager.go:78-119)."""
    scope, policy, max_numa = "container", "none", 8  # defaults :70-76
    pols = nrt.get("policies") or []
    if pols and pols[0] in _POLICIES:
        policy, scope = _POLICIES[pols[0]]
    for name, value in (nrt.get("attributes") or {}).items():
        if name == "topologyManagerScope" and value in ("container", "pod"):
            scope = value
        elif name == "topologyManagerPolicy" and value in ("none", "best-effort", "restricted", "single-numa-node"):
            policy = value
        elif name == "topologyManagerMaxNUMANodes":
            try:
                v = int(value)
            except ValueError:
                continue
            if v > 1:
                max_numa = min(v, 1024)
    return scope, policy, max_numa


def numa_zones(nrt):
    """createNUMANodeList (pluginhelpers.go:105-134): zones of Type 'Node' named node-<id>, id <= 64.
    Returns (zones, supported): the dense encoding needs ids 0..k-1 in list order, id < 64."""
    zones, ids = [], []
    for z in nrt.get("zones", []):
        if z.get("type", "Node") != "Node":
            continue
        m = re.fullmatch(r"node-(\d+)", z["name"])
        if not m or int(m.group(1)) > 64:
            continue
        zones.append(z)
        ids.append(int(m.group(1)))
    supported = ids == list(range(len(ids))) and len(ids) <= Z_MAX
    return zones, ids, supported


def build_dictionary(pods) -> list:
    """Resource-slot dictionary of a batch: every resource name any pod requests (cpu, memory first)."""
    names = ["cpu", "memory"]
    for pod in pods:
        for c in all_containers(pod):
            for r in c.get("requests", {}):
                if r not in names:
                    names.append(r)
        for r in (pod.get("overhead") or {}):
            if r not in names:
                names.append(r)
    return names


def flatten_nrt_nodes(nodes, nrts, names, fresh=None):
    """nodes[i] + nrts[i] (None = no NRT object) -> SoA dict for b200s_snapshot_nrt / orc_nrt_batch."""
    N = len(nodes)
    R = min(len(names), R_MAX)
    zlists = []
    Z = 1
    for nrt in nrts:
        if nrt is None:
            zlists.append(([], [], True))
        else:
            zlists.append(numa_zones(nrt))
            Z = max(Z, min(len(zlists[-1][0]), Z_MAX))
    out = dict(n_zones=Z, n_res=R,
               res_flags=np.array([(RES_AFFINE if is_numa_affine(n) else 0) | (RES_HOST_LEVEL if is_host_level(n) else 0)
                                   for n in names[:R]], dtype=np.uint8),
               node_flags=np.zeros(N, np.uint8), max_numa=np.full(N, 8, np.uint16), n_zones_node=np.zeros(N, np.uint8),
               node_res_mask=np.zeros(N, np.uint8), zone_res_mask=np.zeros((Z, N), np.uint8),
               avail=np.zeros((Z, R, N), np.int64), cost=np.full((Z, Z, N), -1, np.int32))
    for i, (node, nrt) in enumerate(zip(nodes, nrts)):
        flags = NODE_FRESH if (fresh is None or fresh[i]) else 0
        level = node_level_resources(node)
        out["node_res_mask"][i] = sum(1 << r for r, n in enumerate(names[:R]) if n in level)
        if nrt is not None:
            flags |= NODE_HAS_NRT
            scope, policy, max_numa = topology_manager(nrt)
            if policy == "single-numa-node":
                flags |= NODE_SINGLE_NUMA
            if scope == "pod":
                flags |= NODE_SCOPE_POD
            out["max_numa"][i] = max_numa
            zones, ids, ok = zlists[i]
            if not ok:
                flags |= NODE_UNSUPPORTED
            else:
                out["n_zones_node"][i] = len(zones)
                for z, zone in enumerate(zones):
                    for r, n in enumerate(names[:R]):
                        if n in zone["resources"]:
                            out["zone_res_mask"][z, i] |= 1 << r
                            out["avail"][z, r, i] = milli(zone["resources"][n]["available"])
                    for cname, cval in (zone.get("costs") or {}).items():  # extractCosts :136-153
                        m = re.fullmatch(r"node-(\d+)", cname)
                        if m and int(m.group(1)) < len(zones):
                            out["cost"][z, int(m.group(1)), i] = int(cval)
        out["node_flags"][i] = flags
    return out


def flatten_nrt_pods(pods, names):
    P = len(pods)
    R = min(len(names), R_MAX)
    slot = {n: r for r, n in enumerate(names)}
    out = dict(qos=np.zeros(P, np.uint8), flags=np.zeros(P, np.uint8), n_init=np.zeros(P, np.uint8),
               n_app=np.zeros(P, np.uint8), cont_kind=np.zeros((P, C_MAX), np.uint8),
               req_mask=np.zeros((P, C_MAX + 1), np.uint8), req=np.zeros((P, C_MAX + 1, R), np.int64))
    for p, pod in enumerate(pods):
        qos = pod_qos(pod)
        out["qos"][p] = qos
        fl = 0
        if qos == QOS_BEST_EFFORT and not include_non_native(pod):
            fl |= POD_FILTER_BYPASS  # filter.go:180-183
        init, app = list(pod.get("init", [])), list(pod.get("containers", []))
        if len(init) + len(app) > C_MAX:
            fl |= POD_UNSUPPORTED
            init, app = [], []
        out["n_init"][p], out["n_app"][p] = len(init), len(app)

        def put(c, reqs):
            nonlocal fl
            for name, q in reqs.items():
                r = slot[name]
                if r >= R:
                    fl |= POD_UNSUPPORTED
                    continue
                out["req_mask"][p, c] |= 1 << r
                out["req"][p, c, r] = q if isinstance(q, (int, np.integer)) and reqs is eff else milli(q)

        eff = pod_effective_request(pod)
        for c, cont in enumerate(init + app):
            out["cont_kind"][p, c] = (CONT_SIDECAR if cont.get("restart_always") else CONT_INIT) if c < len(init) else CONT_APP
            put(c, cont.get("requests", {}))
        put(C_MAX, eff)
        out["flags"][p] = fl
    return out

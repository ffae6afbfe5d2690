"""Extracts the `scopeEqualsContainerTests` table of the reference's NodeResourceTopologyMatch integration test
(test/integration/noderesourcetopology_test.go:1742-2034, expanded by parseTestUserEntry :2158-2251) into
tests/golden/nrt_integration.json.  Run in the build container (the reference tree is not on the GPU box):
    python tests/golden/extract_nrt_integration.py
Each entry: containers' resource maps (limits for Guaranteed pods -- the API server defaults requests to limits --
or requests for Burstable ones), the expected node set (empty = unschedulable) and the failure message."""
import json
import os
import re

SRC = "/root/reference/test/integration/noderesourcetopology_test.go"
NAMES = {"cpu": "cpu", "memory": "memory", "hugepages2Mi": "hugepages-2Mi", "nicResourceName": "vendor/nic1",
         "ephemeralStorage": "ephemeral-storage", "gpuResourceName": "vendor/gpu", "v1.ResourceCPU": "cpu",
         "v1.ResourceMemory": "memory"}
STRATEGY = {None: "MostAllocated", "mostAllocatedScheduler": "MostAllocated",
            "balancedAllocationScheduler": "BalancedAllocation", "leastAllocatedScheduler": "LeastAllocated",
            "leastNUMAScheduler": "LeastNUMANodes"}


def block(text, start):
    """text[start] == '{' -> index one past the matching '}' (strings are skipped)."""
    depth, i, in_str = 0, start, False
    while i < len(text):
        ch = text[i]
        if in_str:
            if ch == "\\":
                i += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
        elif ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced braces")


def maps_of(entry, field):
    m = re.search(field + r":\s*\[\]map\[string\]string\{", entry)
    if not m:
        return []
    start = entry.index("{", m.end() - 1)
    body = entry[start + 1:block(entry, start) - 1]
    out = []
    for mm in re.finditer(r"\{([^{}]*)\}", body):
        out.append({NAMES[k]: v for k, v in re.findall(r"(\w+):\s*\"([^\"]*)\"", mm.group(1))})
    return out


def split_top(body):
    """top-level comma split of a Go composite-literal body"""
    out, depth, cur, in_str = [], 0, [], False
    for ch in body:
        if in_str:
            cur.append(ch)
            if ch == '"':
                in_str = False
            continue
        if ch == '"':
            in_str = True
        elif ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    tail = "".join(cur).strip()
    if tail:
        out.append(tail)
    return [x for x in out if x]


def res_map(txt):
    return {NAMES[k]: v for k, v in re.findall(r"([\w.]+):\s*\"([^\"]*)\"", txt)}


def parse_pod(expr):
    """containers in creation order: nested util.WithLimits(inner, map, isInit) apply inner first = text order of the
    maps; .Req(map) adds an app container with requests; .Container(image) one without resources"""
    init, app = [], []
    sched = re.search(r"\.SchedulerName\((\w+)\)", expr)
    for m in re.finditer(r"\.Req\(map\[v1\.ResourceName\]string\{([^}]*)\}\)", expr):
        app.append(dict(requests=res_map(m.group(1)), limits={}))
    for m in re.finditer(r"map\[string\]string\{([^}]*)\},\s*(true|false)\)", expr):
        (init if m.group(2) == "true" else app).append(dict(requests={}, limits=res_map(m.group(1))))
    if re.search(r"\.Container\(", expr):
        app.append(dict(requests={}, limits={}))
    assert "WithRequests" not in expr
    return dict(init=init, containers=app, strategy=STRATEGY[sched.group(1) if sched else None])


def parse_nrt(expr):
    name = re.search(r"\.Name\(\"([^\"]+)\"\)", expr).group(1)
    policies = re.findall(r"Policy\(topologyv1alpha2\.(\w+)\)", expr)
    attrs = {}
    for a, v in re.findall(r"Name:\s*nodeconfig\.(\w+),\s*Value:\s*\"([^\"]*)\"", expr):
        attrs[{"AttributePolicy": "topologyManagerPolicy", "AttributeScope": "topologyManagerScope"}[a]] = v
    zones = []
    for m in re.finditer(r"(ZoneWithCosts|Zone)\(", expr):
        start = m.end() - 1
        depth, i = 0, start
        while True:
            depth += expr[i] == "("
            depth -= expr[i] == ")"
            if depth == 0:
                break
            i += 1
        body = expr[start + 1:i]
        res = {NAMES[r]: dict(capacity=c, available=a)
               for r, c, a in re.findall(r"MakeTopologyResInfo\(([\w.]+),\s*\"([^\"]*)\",\s*\"([^\"]*)\"\)", body)}
        costs = {n: int(v) for n, v in re.findall(r"Name:\s*\"([^\"]+)\",\s*Value:\s*(\d+)", body)}
        zones.append(dict(name=f"node-{len(zones)}", resources=res, costs=costs))
    return dict(name=name, policies=policies, attributes=attrs, zones=zones)


def first_table(text):
    m = re.search(r"tests := \[\]nrtTestEntry\{", text)
    start = text.index("{", m.end() - 1)
    table = text[start + 1:block(text, start) - 1]
    cases, i = [], 0
    while True:
        j = table.find("{", i)
        if j < 0:
            break
        end = block(table, j)
        entry = table[j:end]
        i = end
        name = re.search(r"name:\s*\"((?:[^\"\\]|\\.)*)\"", entry).group(1)
        ps = entry.index("{", entry.index("pods:"))
        pods = [parse_pod(x) for x in split_top(entry[ps + 1:block(entry, ps) - 1])]
        nrts = []
        if "nodeResourceTopologies:" in entry:
            ns = entry.index("{", entry.index("nodeResourceTopologies:"))
            nrts = [parse_nrt(x) for x in split_top(entry[ns + 1:block(entry, ns) - 1])]
        exp = re.search(r"expectedNodes:\s*\[\]string\{([^}]*)\}", entry)
        err = re.search(r"errMsg:\s*\"([^\"]*)\"", entry)
        cases.append(dict(name=name, pods=pods, nrts=nrts,
                          expected_nodes=re.findall(r"\"([^\"]+)\"", exp.group(1)) if exp else [],
                          err_msg=err.group(1) if err else ""))
    return cases


def main():
    text = open(SRC).read()
    full = first_table(text)
    out2 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nrt_integration_full.json")
    with open(out2, "w") as f:
        json.dump({"source": "test/integration/noderesourcetopology_test.go:239-1741 (tests := []nrtTestEntry{...}); "
                             "node capacity :215-222; scheduler name -> scoring strategy :170-200 (default profile: "
                             "MostAllocated)",
                   "node_capacity": {"cpu": "64", "memory": "128Gi", "pods": "32", "hugepages-2Mi": "896Mi",
                                     "vendor/nic1": "48", "ephemeral-storage": "32Gi"},
                   "cases": full}, f, indent=1)
    print(f"{len(full)} cases -> {out2}")
    m = re.search(r"scopeEqualsContainerTests := \[\]nrtTestUserEntry\{", text)
    start = text.index("{", m.end() - 1)
    table = text[start + 1:block(text, start) - 1]
    cases, i = [], 0
    while True:
        j = table.find("{", i)
        if j < 0:
            break
        end = block(table, j)
        entry = table[j:end]
        i = end
        desc = re.search(r"description:\s*\"((?:[^\"\\]|\\.)*)\"", entry).group(1)
        err = re.search(r"errMsg:\s*\"([^\"]*)\"", entry)
        exp = re.search(r"expectedNodes:\s*\[\]string\{([^}]*)\}", entry)
        expected = re.findall(r"\"([^\"]+)\"", exp.group(1)) if exp else (["fake-node-1"] if not err else [])
        cases.append(dict(name=desc, init=maps_of(entry, "initCntReq"), containers=maps_of(entry, "cntReq"),
                          burstable=bool(re.search(r"isBurstable:\s*true", entry)), expected_nodes=expected,
                          err_msg=err.group(1) if err else ""))
    zone = lambda c, m_, h, n: {"cpu": c, "memory": m_, "hugepages-2Mi": h, "vendor/nic1": n}  # noqa: E731
    doc = {
        "source": "test/integration/noderesourcetopology_test.go:1742-2034 (scopeEqualsContainerTests) with the fixed "
                  "NRT objects of parseTestUserEntry :2176-2228 and the node capacity of :215-222; profile: Filter + "
                  "Score enabled, ScoringStrategy MostAllocated (:176-181)",
        "node_capacity": {"cpu": "64", "memory": "128Gi", "pods": "32", "hugepages-2Mi": "896Mi", "vendor/nic1": "48",
                          "ephemeral-storage": "32Gi"},
        "attributes": {"topologyManagerPolicy": "single-numa-node", "topologyManagerScope": "container"},
        "nrts": {"fake-node-1": [zone("30", "60Gi", "384Mi", "16"), zone("32", "64Gi", "512Mi", "32")],
                 "fake-node-2": [zone("0", "0", "0", "0"), zone("0", "0", "0", "0")]},
        "cases": cases,
    }
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nrt_integration.json")
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    print(f"{len(cases)} cases -> {out}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/nrt_filter.json from the reference's own table-driven Go tests.

Run in the build container (where /root/reference exists):
    python tests/golden/extract_nrt_golden.py
The Go toolchain is absent, so the tables are parsed textually: node fixtures
(NodeResourceTopology literals), pod constructors (makePodByResourceList & co,
pkg/noderesourcetopology/objects.go:26-125; makePod/withMultiContainers, filter_test.go:1206-1243;
testUserEntry, filter_test.go:1253-1290) and the expected status message.  Nothing from the
reference is copied into the repo except these data vectors.
"""
import json
import os
import re
import sys

REF = "/root/reference/pkg/noderesourcetopology"
OUT = os.path.dirname(os.path.abspath(__file__))

CONST = {
    "cpu": "cpu", "memory": "memory", "v1.ResourceCPU": "cpu", "v1.ResourceMemory": "memory",
    "v1.ResourceEphemeralStorage": "ephemeral-storage", "extended": "namespace/extended",
    "hugepages2Mi": "hugepages-2Mi", "nicResourceName": "vendor/nic1",
    "notExistingNICResourceName": "vendor/notexistingnic", "nicResourceNameNoNUMA": "vendor.com/old-nic-model",
    "gpuResourceName": "vendor/gpu", "corev1.ResourceCPU": "cpu", "corev1.ResourceMemory": "memory",
}


def match_brace(s, i, open_="{", close="}"):
    """index just past the brace group that opens at s[i]"""
    assert s[i] == open_, s[i:i + 20]
    depth = 0
    j = i
    in_str = False
    while j < len(s):
        ch = s[j]
        if in_str:
            if ch == "\\":
                j += 1
            elif ch == '"':
                in_str = False
        elif ch == '"':
            in_str = True
        elif ch == open_:
            depth += 1
        elif ch == close:
            depth -= 1
            if depth == 0:
                return j + 1
        j += 1
    raise ValueError("unbalanced")


def top_level_groups(body):
    """brace groups at depth 0 of `body` (text inside an outer {...})"""
    out, i = [], 0
    while i < len(body):
        if body[i] == '"':
            i = body.index('"', i + 1) + 1
            continue
        if body[i] == "{":
            j = match_brace(body, i)
            out.append(body[i + 1:j - 1])
            i = j
        else:
            i += 1
    return out


def res_name(tok):
    tok = tok.strip()
    if tok.startswith('"'):
        return tok.strip('"')
    m = re.match(r"v1\.ResourceName\((\w+)\)", tok)
    if m:
        tok = m.group(1)
    if tok not in CONST:
        raise KeyError(tok)
    return CONST[tok]


def parse_resource_list(body):
    """`KEY: *resource.NewQuantity(N, ...)` / `KEY: resource.MustParse("S")` / `KEY: "S"` entries"""
    out = {}
    for m in re.finditer(r'([\w\.\(\)"/\-]+)\s*:\s*(?:\*?resource\.NewQuantity\((\-?\d+),[^)]*\)|resource\.MustParse\("([^"]*)"\)|"([^"]*)")', body):
        key = res_name(m.group(1))
        out[key] = m.group(2) if m.group(2) is not None else (m.group(3) if m.group(3) is not None else m.group(4))
    return out


def resource_lists_after(text, marker):
    """all v1.ResourceList{...} bodies following `marker`"""
    out = []
    for m in re.finditer(re.escape(marker), text):
        i = text.index("{", m.end() - 1)
        out.append(text[i + 1:match_brace(text, i) - 1])
    return out


def cont(req, lim=None):
    c = {"requests": req}
    c["limits"] = req if lim is None else lim
    return c


def parse_pod(txt):
    txt = txt.strip()
    if txt.startswith("&v1.Pod{}"):
        return {"init": [], "containers": []}
    rls = [parse_resource_list(b) for b in resource_lists_after(txt, "v1.ResourceList{")]
    if txt.startswith("makePodByResourceListWithManyContainers"):
        n = int(re.search(r"},\s*(\d+)\)", txt).group(1))
        return {"init": [], "containers": [cont(dict(rls[0])) for _ in range(n)]}
    if txt.startswith("makePodByResourceLists"):
        i = txt.index("(")
        groups = top_level_groups(txt[i + 1:match_brace(txt, i, "(", ")") - 1])
        return {"init": [], "containers": [cont(parse_resource_list(g)) for g in groups]}
    if txt.startswith("makePodByResourceList"):
        return {"init": [], "containers": [cont(rls[0])]}
    if txt.startswith("makePodWithReqByResourceList"):
        return {"init": [], "containers": [{"requests": rls[0], "limits": {}}]}
    if txt.startswith("makePodWithReqAndLimitByResourceList"):
        return {"init": [], "containers": [cont(rls[0], rls[1])]}
    if txt.startswith("makePod("):
        pod = {"init": [], "containers": []}
        for key, field in (("withMultiInitContainers(", "init"), ("withMultiContainers(", "containers")):
            k = txt.find(key)
            if k < 0:
                continue
            i = txt.index("{", k)
            body = txt[i + 1:match_brace(txt, i) - 1]
            pod[field] = [cont(parse_resource_list(g)) for g in top_level_groups(body)]
        return pod
    raise ValueError("unknown pod constructor: " + txt[:60])


def parse_nrts(text):
    """NodeResourceTopology literals -> [{name, policies, attributes, zones, node_extra}]"""
    nodes = []
    for m in re.finditer(r"ObjectMeta:\s*metav1\.ObjectMeta\{Name:\s*\"([^\"]+)\"\}", text):
        # the enclosing literal starts at the previous '{' at the same nesting: walk back
        start = text.rfind("{", 0, m.start())
        end = match_brace(text, start)
        lit = text[start:end]
        pol = re.findall(r"TopologyPolicies:\s*\[\]string\{string\(topologyv1alpha2\.(\w+)\)\}", lit)
        attrs = dict(re.findall(r'\{Name:\s*"?([\w\.]+)"?,\s*Value:\s*"([^"]*)"\}', lit))
        zones = []
        zi = lit.find("Zones:")
        if zi >= 0:
            zb = lit.index("{", zi)
            for zt in top_level_groups(lit[zb + 1:match_brace(lit, zb) - 1]):
                zn = re.search(r'Name:\s*"([^"]+)"', zt)
                ty = re.search(r'Type:\s*"([^"]+)"', zt)
                res = {}
                for r in re.finditer(r'MakeTopologyResInfo\(([\w\.]+),\s*"([^"]*)",\s*"([^"]*)"\)', zt):
                    res[res_name(r.group(1))] = {"capacity": r.group(2), "available": r.group(3)}
                costs = {}
                ci = zt.find("Costs:")
                if ci >= 0:
                    for c in re.finditer(r'Name:\s*"([^"]+)",\s*Value:\s*(\d+)', zt[ci:]):
                        costs[c.group(1)] = int(c.group(2))
                zones.append({"name": zn.group(1), "type": ty.group(1) if ty else "Node", "resources": res, "costs": costs})
        nodes.append({"name": m.group(1), "policies": pol, "attributes": attrs, "zones": zones, "_span": (start, end)})
    return nodes


def split_cases(table_body):
    return top_level_groups(table_body)


def table_after(text, marker):
    k = text.index(marker)
    i = text.index("{", text.index("}{", k) + 1) if False else None
    # tests := []struct { ... }{  <cases> }
    j = text.index("{", k)            # struct field block
    j2 = match_brace(text, j)
    assert text[j2] == "{"
    return text[j2 + 1:match_brace(text, j2) - 1]


def want_of(case_txt):
    m = re.search(r'wantStatus:\s*fwk\.NewStatus\(fwk\.(\w+),\s*"([^"]*)"', case_txt)
    if m:
        return {"code": m.group(1), "message": m.group(2)}
    return None


def func_body(text, name):
    k = text.index("func " + name + "(")
    i = text.index("{", text.index(")", k))
    # skip the parameter list's closing paren properly
    i = text.index("{", text.index("testing.T)", k))
    return text[i:match_brace(text, i)]


def node_allocatable(nrt, extra):
    """makeResourceListFromZones (objects.go:89-101): sum of zone Available + per-test extras"""
    return {"from_zones_available": True, "extra": extra}


def main():
    src = open(os.path.join(REF, "filter_test.go")).read()
    out = {"source": "pkg/noderesourcetopology/filter_test.go (TestNodeResourceTopology :55-713, "
                     "TestNodeResourceTopologyMultiContainerPodScope :715-941, "
                     "TestNodeResourceTopologyMultiContainerContainerScope :943-1181); node allocatable = sum of "
                     "zone Available (makeResourceListFromZones, objects.go:89-101) + extras; the NRT cache is "
                     "Passthrough (filter_test.go:688-690). Statuses are compared as code + message prefix "
                     "(quasiEqualStatus :1292-1305).",
           "generated_by": "tests/golden/extract_nrt_golden.py", "suites": []}

    # ---- suite 1: TestNodeResourceTopology
    body = func_body(src, "TestNodeResourceTopology")
    nrts = parse_nrts(body[:body.index("tests := []struct")])
    # per-node extra allocatable: `node: v1.ResourceList{ v1.ResourceName(x): resource.MustParse("4") }` inside the desc
    descs_txt = body[:body.index("nodes := make(")]
    for n in nrts:
        s, e = n.pop("_span")
        tail = descs_txt[e:e + 400]
        m = re.match(r"\s*,\s*node:\s*v1\.ResourceList\{", tail)
        extra = {}
        if m:
            i = e + m.end() - 1
            extra = parse_resource_list(descs_txt[i + 1:match_brace(descs_txt, i) - 1])
        n["node_extra"] = extra
    cases = []
    for ct in split_cases(table_after(body, "tests := []struct")):
        name = re.search(r'name:\s*"((?:[^"\\]|\\.)*)"', ct).group(1)
        pod_txt = ct[ct.index("pod:") + 4:ct.rindex("node:")]
        node = int(re.search(r"node:\s*nodes\[(\d+)\]", ct).group(1))
        cases.append({"name": name, "pod": parse_pod(pod_txt.strip().rstrip(",")), "node": node, "want": want_of(ct)})
    out["suites"].append({"suite": "TestNodeResourceTopology", "nodes": nrts, "cases": cases})

    # ---- suite 2: pod scope multi-container
    body = func_body(src, "TestNodeResourceTopologyMultiContainerPodScope")
    nrts = parse_nrts(body[:body.index("nodes := make(")])
    for n in nrts:
        n.pop("_span")
        n["node_extra"] = {}
    cases = []
    for ct in split_cases(table_after(body, "tests := []struct")):
        name = re.search(r'name:\s*"((?:[^"\\]|\\.)*)"', ct).group(1)
        pod_txt = ct[ct.index("pod:") + 4:ct.index("node:")]
        cases.append({"name": name, "pod": parse_pod(pod_txt.strip().rstrip(",")), "node": 0, "want": want_of(ct)})
    out["suites"].append({"suite": "TestNodeResourceTopologyMultiContainerPodScope", "nodes": nrts, "cases": cases})

    # ---- suite 3: container scope, testUserEntry table
    body = func_body(src, "TestNodeResourceTopologyMultiContainerContainerScope")
    nrts = parse_nrts(body[:body.index("nodes := make(")])
    for n in nrts:
        n.pop("_span")
        n["node_extra"] = {}
    k = body.index("tue := []testUserEntry")
    i = body.index("{", k)
    cases = []
    for ct in top_level_groups(body[i + 1:match_brace(body, i) - 1]):
        name = re.search(r'description:\s*"((?:[^"\\]|\\.)*)"', ct).group(1)
        pod = {"init": [], "containers": []}
        for key, field in (("initCntReq:", "init"), ("cntReq:", "containers")):
            m = re.search(r"(?<![A-Za-z])" + key, ct)
            if not m:
                continue
            b = ct.index("{", m.end())
            pod[field] = [cont(parse_resource_list(g)) for g in top_level_groups(ct[b + 1:match_brace(ct, b) - 1])]
        m = re.search(r'statusErr:\s*"([^"]*)"', ct)
        cases.append({"name": name, "pod": pod, "node": 0,
                      "want": {"code": "Unschedulable", "message": m.group(1)} if m and m.group(1) else None})
    out["suites"].append({"suite": "TestNodeResourceTopologyMultiContainerContainerScope", "nodes": nrts, "cases": cases})

    with open(os.path.join(OUT, "nrt_filter.json"), "w") as f:
        json.dump(out, f, indent=1)
    for s in out["suites"]:
        print(s["suite"], "nodes", len(s["nodes"]), "cases", len(s["cases"]),
              "rejects", sum(1 for c in s["cases"] if c["want"]))


def fixture_func(text, name):
    k = text.index("func " + name + "(")
    i = text.index("{", text.index("NodeResourceTopology {", k) + len("NodeResourceTopology"))
    body = text[i:match_brace(text, i)]
    nrts = parse_nrts(body)
    for n in nrts:
        n.pop("_span")
        n["node_extra"] = {}
    return nrts


def main_scores():
    src = open(os.path.join(REF, "score_test.go")).read()
    CONST["gpu"] = "gpu"
    default_nodes = fixture_func(src, "defaultNUMANodes")
    four_nodes = fixture_func(src, "fourNUMANodes")
    out = {"source": "pkg/noderesourcetopology/score_test.go — TestNodeResourceScorePlugin :87-194 (fixtures "
                     "defaultNUMANodes :643-714 with policy SingleNUMANodeContainerLevel; only the arg-max node and its "
                     "score are asserted), TestNodeResourceScorePluginLeastNUMA :196-480 (fixtures defaultNUMANodes / "
                     "fourNUMANodes :716-947; every node's score asserted). Pods are makePodByResourceList(s): "
                     "requests == limits (Guaranteed).",
           "generated_by": "tests/golden/extract_nrt_golden.py", "suites": []}
    pod = {"init": [], "containers": [cont({"cpu": "2", "memory": str(20 * 1024 * 1024)})]}
    out["suites"].append({
        "suite": "TestNodeResourceScorePlugin", "nodes": default_nodes, "policy_override": "SingleNUMANodeContainerLevel",
        "cases": [{"name": "MostAllocated strategy", "strategy": "MostAllocated", "pod": pod, "want_max": {"Node2": 70}},
                  {"name": "BalancedAllocation strategy", "strategy": "BalancedAllocation", "pod": pod, "want_max": {"Node3": 100}},
                  {"name": "LeastAllocated strategy", "strategy": "LeastAllocated", "pod": pod, "want_max": {"Node1": 73}}]})
    body = func_body(src, "TestNodeResourceScorePluginLeastNUMA")
    cases = []
    for ct in split_cases(table_after(body, "testCases := []struct")):
        name = re.search(r'name:\s*"((?:[^"\\]|\\.)*)"', ct).group(1)
        k = ct.index("podRequests:")
        i = ct.index("{", k)
        conts = [cont(parse_resource_list(g)) for g in top_level_groups(ct[i + 1:match_brace(ct, i) - 1])]
        k = ct.index("wantedRes:")
        i = ct.index("{", k)
        want = {m.group(1): int(m.group(2)) for m in re.finditer(r'"(\w+)":\s*(\d+)', ct[i:match_brace(ct, i)])}
        nm = re.search(r"nodes:\s*(\w+)\((?:withPolicy\(topologyv1alpha2\.(\w+)\))?\)", ct)
        cases.append({"name": name, "strategy": "LeastNUMANodes", "pod": {"init": [], "containers": conts},
                      "fixture": nm.group(1), "policy_override": nm.group(2), "want": want})
    out["suites"].append({"suite": "TestNodeResourceScorePluginLeastNUMA",
                          "fixtures": {"defaultNUMANodes": default_nodes, "fourNUMANodes": four_nodes}, "cases": cases})
    with open(os.path.join(OUT, "nrt_score.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("score: default nodes", len(default_nodes), "four-NUMA nodes", len(four_nodes), "LeastNUMA cases", len(cases))


if __name__ == "__main__":
    main()
    main_scores()

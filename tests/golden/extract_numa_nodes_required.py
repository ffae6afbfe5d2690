"""Extracts TestNUMANodesRequired (pkg/noderesourcetopology/least_numa_test.go:35-704) into
tests/golden/numa_nodes_required.json.  Run in the build container:  python tests/golden/extract_numa_nodes_required.py"""
import json
import os
import re

SRC = "/root/reference/pkg/noderesourcetopology/least_numa_test.go"
NAMES = {"v1.ResourceCPU": "cpu", "v1.ResourceMemory": "memory", "gpuResource": "gpu"}


def block(text, start):
    depth, i, in_str = 0, start, False
    while i < len(text):
        ch = text[i]
        if in_str:
            if ch == '"':
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
    raise ValueError("unbalanced")


def resources(txt):
    out = {}
    for name, a, b in re.findall(r"([\w.]+):\s*(?:\*resource\.NewQuantity\((\d+),[^)]*\)|resource\.MustParse\(\"([^\"]+)\"\))", txt):
        out[NAMES[name]] = a or b
    return out


def main():
    text = open(SRC).read()
    m = re.search(r"func TestNUMANodesRequired", text)
    t0 = text.index("testCases := []struct", m.end())
    start = text.index("}{", t0) + 1
    table = text[start + 1:block(text, start) - 1]
    cases, i = [], 0
    while True:
        j = table.find("{", i)
        if j < 0:
            break
        end = block(table, j)
        entry = table[j:end]
        i = end
        desc = re.search(r"description:\s*\"([^\"]*)\"", entry).group(1)
        ns = entry.index("{", entry.index("numaNodes:"))
        body = entry[ns + 1:block(entry, ns) - 1]
        zones, k = [], 0
        while True:
            a = body.find("{", k)
            if a < 0:
                break
            b = block(body, a)
            z = body[a:b]
            k = b
            zid = int(re.search(r"NUMAID:\s*(\d+)", z).group(1))
            rs = z.index("{", z.index("Resources:"))
            res = resources(z[rs:block(z, rs)])
            costs = {}
            if "Costs:" in z:
                cs = z.index("{", z.index("Costs:"))
                costs = {int(x): int(y) for x, y in re.findall(r"(\d+):\s*(\d+)", z[cs:block(z, cs)])}
            zones.append(dict(id=zid, resources=res, costs=costs))
        ps = entry.index("{", entry.index("podResources:"))
        pod = resources(entry[ps:block(entry, ps)])
        bm = re.search(r"expectedBitmask:\s*NewTestBitmask\(([^)]*)\)", entry)
        bits = [int(x) for x in re.findall(r"\d+", bm.group(1))] if bm else None
        mind = re.search(r"expectedMinDistance:\s*(true|false)", entry)
        cases.append(dict(name=desc, zones=zones, pod=pod, expected_bits=bits,
                          expected_min_distance=(mind.group(1) == "true") if mind else False,
                          expect_error="expectedErr:" in entry and "expectedErr:         nil" not in entry and "expectedErr: nil" not in entry))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "numa_nodes_required.json")
    with open(out, "w") as f:
        json.dump({"source": "pkg/noderesourcetopology/least_numa_test.go:35-704 (TestNUMANodesRequired), qos Guaranteed; "
                             "zones in LIST order with their NUMA ids and cost maps", "cases": cases}, f, indent=1)
    print(len(cases), "cases ->", out)


if __name__ == "__main__":
    main()

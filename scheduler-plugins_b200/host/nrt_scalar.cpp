#include "nrt_scalar.hpp"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>

#include "../../include/b200sched.h"

namespace b200host {

namespace {

constexpr int kMaxNUMAId = 64;     // pluginhelpers.go:42
constexpr int kMaxDistance = 255;  // least_numa.go:32

long NameToID(const std::string& name) {  // numanode.NameToID: "node-<decimal>"
  if (name.rfind("node-", 0) != 0 || name.size() == 5 || name.size() > 14) return -1;
  long id = 0;
  for (size_t i = 5; i < name.size(); ++i) {
    if (name[i] < '0' || name[i] > '9') return -1;
    id = id * 10 + (name[i] - '0');
  }
  return id;
}

// isResourceSetSuitable: numaresources.go:137-142
bool Suitable(QOS qos, const std::string& res, int64_t qty, int64_t numa_qty) {
  if (qos != QOS::Guaranteed && IsNUMAAffineResource(res)) return true;
  return numa_qty >= qty;
}

// resourcesAvailableInAnyNUMANodes: filter.go:90-160.  Returns the chosen NUMA id, or -1 with ok = false.
int AvailableInAnyNUMANodes(const NUMANodeList& numa_nodes, const ResourceList& resources, QOS qos, const NodeInfo& node,
                            bool* ok) {
  uint64_t bitmask = ~0ull;  // bm.NewEmptyBitMask(); Fill()
  // util.ResourceList(nodeInfo.GetAllocatable()): cpu, memory, pods, ephemeral-storage + scalars (pkg/util/resource.go:28-44)
  auto node_has = [&](const std::string& r) {
    if (r == ResourceCPU || r == ResourceMemory || r == ResourcePods || r == ResourceEphemeralStorage) return true;
    const Node* nd = node.GetNode();
    return nd && nd->allocatable.count(r) && IsScalarResourceName(r);
  };
  for (const auto& [resource, quantity] : resources) {
    if (quantity == 0) continue;  // :101-105
    if (!node_has(resource)) {    // :107-113
      *ok = false;
      return -1;
    }
    uint64_t res_bits = 0;
    bool has_numa_affinity = false;
    for (const auto& numa : numa_nodes) {
      auto it = numa.resources.find(resource);
      if (it == numa.resources.end()) continue;  // :121-125
      has_numa_affinity = true;
      if (!Suitable(qos, resource, quantity, it->second)) continue;
      if (numa.numa_id >= 0 && numa.numa_id < 64) res_bits |= 1ull << numa.numa_id;
    }
    if (!has_numa_affinity && IsHostLevelResource(resource)) continue;  // :139-142
    bitmask &= res_bits;
    if (bitmask == 0) {  // :144-148
      *ok = false;
      return -1;
    }
  }
  *ok = true;
  return __builtin_ctzll(bitmask);  // lowest set bit (:154); an unconstrained request keeps bit 0
}

// subtractResourcesFromNUMANodeList: numaresources.go:145-182 (every list entry with that NUMA id)
bool SubtractFromNUMANodeList(NUMANodeList& nodes, int numa_id, QOS qos, const ResourceList& container_res) {
  for (auto& node : nodes) {
    if (node.numa_id != numa_id) continue;
    for (const auto& [res, qty] : container_res) {
      if (qos != QOS::Guaranteed && IsNUMAAffineResource(res)) continue;
      if (qty == 0) continue;
      auto it = node.resources.find(res);
      if (it == node.resources.end()) continue;
      if (it->second - qty < 0) return false;
      it->second -= qty;
    }
  }
  return true;
}

int64_t GoDiv(int64_t a, int64_t b) { return b == 0 ? 0 : a / b; }
int64_t F2I(double x) {  // amd64 int64(float64)
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
int64_t WeightOf(const std::map<std::string, int64_t>& w, const std::string& r) {  // score.go:49-60
  auto it = w.find(r);
  return (it == w.end() || it->second < 1) ? 1 : it->second;
}

// least_allocated.go:25-55 / most_allocated.go:25-54 / balanced_allocation.go:27-54 on one zone
int64_t StrategyScore(int strategy, const ResourceList& requested, const ResourceList& allocatable,
                      const std::map<std::string, int64_t>& weights) {
  if (strategy == B200S_NRT_BALANCED_ALLOCATION) {
    std::vector<double> fractions;
    for (const auto& [res, req] : requested) {
      auto it = allocatable.find(res);
      const int64_t cap = it == allocatable.end() ? 0 : QuantityValue(it->second);
      const double f = cap == 0 ? 1.0 : (double)QuantityValue(req) / (double)cap;
      if (f > 1) return 0;
      fractions.push_back(f);
    }
    // gonum stat.Variance (unbiased, two-pass with compensation)
    const double n = (double)fractions.size();
    double sum = 0;
    for (double f : fractions) sum += f;
    const double mean = sum / n;
    double ss = 0, comp = 0;
    for (double f : fractions) {
      const double d = f - mean;
      ss += d * d;
      comp += d;
    }
    const double variance = (ss - comp * comp / n) / (n - 1);
    return F2I((1 - variance) * 100.0);
  }
  int64_t node_score = 0, weight_sum = 0;
  for (const auto& [res, req] : requested) {
    auto it = allocatable.find(res);
    const int64_t cap = it == allocatable.end() ? 0 : it->second;
    int64_t s = 0;
    if (cap != 0 && req <= cap) {
      const int64_t cv = QuantityValue(cap), rv = QuantityValue(req);
      s = strategy == B200S_NRT_MOST_ALLOCATED ? GoDiv(rv * 100, cv) : GoDiv((cv - rv) * 100, cv);
    }
    const int64_t w = WeightOf(weights, res);
    node_score += s * w;
    weight_sum += w;
  }
  return GoDiv(node_score, weight_sum);
}

// scoreForEachNUMANode: score.go:110-124
int64_t ScoreForEachNUMANode(int strategy, const ResourceList& requested, const NUMANodeList& numa_list,
                             const std::map<std::string, int64_t>& weights) {
  std::vector<int64_t> numa_scores(numa_list.size(), 0);
  int64_t min_score = 0;
  for (const auto& numa : numa_list) {
    const int64_t s = StrategyScore(strategy, requested, numa.resources, weights);
    if (min_score == 0 || (s != 0 && s < min_score)) min_score = s;
    numa_scores.at((size_t)numa.numa_id) = s;  // :120 -- Go panics when NUMAID >= len(numaList)
  }
  return min_score;
}

// gonum combin.Combinations(n, k): lexicographic index tuples
std::vector<std::vector<int>> Combinations(int n, int k) {
  std::vector<std::vector<int>> out;
  std::vector<int> idx(k);
  for (int i = 0; i < k; ++i) idx[i] = i;
  while (true) {
    out.push_back(idx);
    int i = k - 1;
    while (i >= 0 && idx[i] == n - k + i) --i;
    if (i < 0) break;
    ++idx[i];
    for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
  }
  return out;
}

float NodesAvgDistance(const NUMANodeList& numa_nodes, const std::vector<int>& nodes) {  // least_numa.go:116-138
  if (nodes.empty()) return (float)kMaxDistance;
  int accu = 0;
  for (int n1 : nodes)
    for (int n2 : nodes) {
      auto it = numa_nodes[n1].costs.find(numa_nodes[n2].numa_id);
      accu += it == numa_nodes[n1].costs.end() ? kMaxDistance : it->second;
    }
  return (float)accu / (float)(nodes.size() * nodes.size());
}

}  // namespace

bool OnlyNonNUMAResources(const NUMANodeList& numa_nodes, const ResourceList& resources) {  // pluginhelpers.go:163-173
  for (const auto& [res, q] : resources) {
    (void)q;
    for (const auto& numa : numa_nodes)
      if (numa.resources.count(res)) return false;
  }
  return true;
}

namespace {

// subtractFromNUMAs: numaresources.go:184-215 -- `nodes` are NUMA ids used as LIST indices, exactly as the reference
void SubtractFromNUMAs(const ResourceList& resources, NUMANodeList& numa_nodes, const std::vector<int>& nodes) {
  for (const auto& [res, qty] : resources) {
    int64_t quantity = qty;
    for (int node : nodes) {
      if (quantity == 0) break;
      auto& n_res = numa_nodes.at((size_t)node).resources;  // Go panics beyond the list
      auto it = n_res.find(res);
      if (it == n_res.end()) continue;
      if (quantity >= it->second) {
        quantity -= it->second;
        it->second = 0;
      } else {
        it->second -= quantity;
        quantity = 0;
      }
    }
  }
}

int64_t NormalizeLeastNUMA(int count, bool is_min, int max_numa) {  // least_numa.go:91-100
  const int64_t unit = 100 / (int64_t)max_numa;
  const int64_t s = 100 - (int64_t)count * unit;
  return is_min ? s + unit / 2 : s;
}

std::vector<const Container*> AllContainers(const Pod& pod) {
  std::vector<const Container*> v;
  for (const auto& c : pod.init_containers) v.push_back(&c);
  for (const auto& c : pod.containers) v.push_back(&c);
  return v;
}

}  // namespace

NUMANodeList CreateNUMANodeList(const NodeResourceTopology& nrt) {
  NUMANodeList nodes;
  std::vector<int> zone_of_id(kMaxNUMAId, 0);
  for (size_t i = 0; i < nrt.zones.size(); ++i) {
    const Zone& z = nrt.zones[i];
    if (z.type != "Node") continue;
    const long id = NameToID(z.name);
    if (id < 0 || id > kMaxNUMAId) continue;  // `numaID > maxNUMAId` (:113) lets 64 through and then indexes out of range
    if (id == kMaxNUMAId) throw std::out_of_range("NUMA id 64 indexes numaIDToZoneIDx out of range (pluginhelpers.go:120)");
    zone_of_id[(size_t)id] = (int)i;
    NUMANode n;
    n.numa_id = (int)id;
    for (const auto& [name, r] : z.resources) n.resources[name] = r.available;  // extractResources: Available
    nodes.push_back(std::move(n));
  }
  for (auto& n : nodes)  // extractCosts of the LAST zone that carried this id (:131-133)
    for (const auto& [name, v] : nrt.zones[(size_t)zone_of_id[(size_t)n.numa_id]].costs) {
      const long id = NameToID(name);
      if (id >= 0) n.costs[(int)id] = (int)v;
    }
  return nodes;
}

Status ScalarFilter(const Pod& pod, const NodeInfo& node, const NodeResourceTopology& nrt) {
  const TopologyManager conf = TopologyManagerFromNodeResourceTopology(nrt);
  if (conf.policy != "single-numa-node") return {};  // filter.go:228-230
  NUMANodeList numa_nodes = CreateNUMANodeList(nrt);
  const QOS qos = GetPodQOS(pod);
  bool ok = false;
  if (conf.scope == "pod") {  // singleNUMAPodLevelHandler :162-173
    AvailableInAnyNUMANodes(numa_nodes, GetPodEffectiveRequest(pod), qos, node, &ok);
    return ok ? Status{} : Status{Code::Unschedulable, "cannot align pod"};
  }
  for (const auto& c : pod.init_containers) {  // :43-55
    AvailableInAnyNUMANodes(numa_nodes, c.requests, qos, node, &ok);
    if (!ok) return {Code::Unschedulable, c.restart_always ? "cannot align sidecar container" : "cannot align init container"};
  }
  for (const auto& c : pod.containers) {  // :57-76
    const int numa_id = AvailableInAnyNUMANodes(numa_nodes, c.requests, qos, node, &ok);
    if (!ok) return {Code::Unschedulable, "cannot align container"};
    if (!SubtractFromNUMANodeList(numa_nodes, numa_id, qos, c.requests)) return {Code::Error, "inconsistent resource accounting"};
  }
  return {};
}

std::vector<int> NumaNodesRequired(QOS qos, const NUMANodeList& numa_nodes, const ResourceList& resources, bool* is_min_distance) {
  const int n = (int)numa_nodes.size();
  for (int k = 1; k <= n; ++k) {
    const auto combos = Combinations(n, k);
    float min_avg = (float)kMaxDistance;  // minAvgDistanceInCombinations :102-114
    for (const auto& c : combos) min_avg = std::min(min_avg, NodesAvgDistance(numa_nodes, c));
    const std::vector<int>* best = nullptr;
    float min_distance = 256;
    bool is_min = false;
    for (const auto& c : combos) {  // findSuitableCombination :179-208
      bool valid = true;  // isValidCombineResources :224-233
      for (int idx : c)
        for (const auto& [res, q] : resources) {
          (void)q;
          if (!numa_nodes[idx].resources.count(res)) valid = false;
        }
      if (!valid) continue;
      ResourceList combined;  // combineResources :140-154
      for (int idx : c)
        for (const auto& [res, q] : numa_nodes[idx].resources) combined[res] += q;
      bool fit = true;  // checkResourcesFit :210-222
      for (const auto& [res, q] : resources) {
        if (q == 0) continue;
        auto it = combined.find(res);
        if (!Suitable(qos, res, q, it == combined.end() ? 0 : it->second)) fit = false;
      }
      if (!fit) continue;
      const float d = NodesAvgDistance(numa_nodes, c);
      if (d == min_avg) {
        best = &c;
        is_min = true;
        break;
      }
      if (d < min_distance) {
        min_distance = d;
        best = &c;
      }
    }
    if (best) {
      std::vector<int> ids;
      for (int idx : *best) ids.push_back(numa_nodes[idx].numa_id);  // bm.Add(numaNodes[nodeIdx].NUMAID)
      std::sort(ids.begin(), ids.end());                            // bm.GetBits(): ascending
      *is_min_distance = is_min;
      return ids;
    }
  }
  *is_min_distance = false;
  return {};
}

int64_t ScalarScore(const Pod& pod, const NodeResourceTopology& nrt, int strategy, const std::map<std::string, int64_t>& weights) {
  const TopologyManager conf = TopologyManagerFromNodeResourceTopology(nrt);
  NUMANodeList numa_nodes = CreateNUMANodeList(nrt);
  const QOS qos = GetPodQOS(pod);
  const bool scope_pod = conf.scope == "pod";
  if (strategy == B200S_NRT_LEAST_NUMA_NODES) {  // score.go:168-176: no policy check
    if (scope_pod) {                             // leastNUMAPodScopeScore :73-89
      const ResourceList req = GetPodEffectiveRequest(pod);
      if (OnlyNonNUMAResources(numa_nodes, req)) return 100;
      bool is_min = false;
      const auto ids = NumaNodesRequired(qos, numa_nodes, req, &is_min);
      return ids.empty() ? 0 : NormalizeLeastNUMA((int)ids.size(), is_min, conf.max_numa_nodes);
    }
    int max_count = 0;  // leastNUMAContainerScopeScore :35-71
    bool all_min = true;
    for (const Container* c : AllContainers(pod)) {
      if (OnlyNonNUMAResources(numa_nodes, c->requests)) continue;
      bool is_min = false;
      const auto ids = NumaNodesRequired(qos, numa_nodes, c->requests, &is_min);
      if (ids.empty()) return 0;
      if (!is_min) all_min = false;
      max_count = std::max(max_count, (int)ids.size());
      SubtractFromNUMAs(c->requests, numa_nodes, ids);
    }
    return max_count == 0 ? 100 : NormalizeLeastNUMA(max_count, all_min, conf.max_numa_nodes);
  }
  if (conf.policy != "single-numa-node") return 0;  // score.go:178-191
  if (scope_pod) return ScoreForEachNUMANode(strategy, GetPodEffectiveRequest(pod), numa_nodes, weights);  // :142-150
  const auto conts = AllContainers(pod);  // containerScopeScore :152-165: stat.Mean over init + app containers, truncated
  double sum = 0;
  for (const Container* c : conts) sum += (double)ScoreForEachNUMANode(strategy, c->requests, numa_nodes, weights);
  return F2I(sum / (double)conts.size());
}

}  // namespace b200host

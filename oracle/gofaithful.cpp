// Oracle / CPU baseline: "Go-faithful" restatement of the reference's CPU path for
// NodeResourcesAllocatable — TEST INFRASTRUCTURE, see oracle.h.
//
// Where oracle/alloc.c restates the ARITHMETIC on flat columns, this file restates the reference's
// per-call STRUCTURE, because that is what the Go scheduler actually executes per (pod, node):
//   Allocatable.Score (allocatable.go:63) -> resourceAllocationScorer.score (resource_allocation.go:49-76):
//     two make(map) per call (:60-61), for each configured resource calculateResourceAllocatableRequest
//     (:79-100: string switch on the resource name) incl. calculatePodResourceRequest (:105-131: loops
//     over the pod's containers with a map lookup each), then the scorer closure ranging over the weight
//     map (allocatable.go:117-128); finally NormalizeScore over the NodeScoreList (allocatable.go:143-168).
// One scheduling cycle (= one pod over all feasible nodes) runs on one thread; cycles of different pods run
// on `threads` threads — more generous to the CPU than upstream, which schedules one pod at a time with a
// 16-goroutine fan-out over nodes (targetloadpacking_test.go:386-405 mirrors that Parallelizer).
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "oracle.h"

namespace {

using ResourceMap = std::unordered_map<std::string, int64_t>;

struct Container {
  ResourceMap requests;
};
struct Pod {
  std::vector<Container> containers, init_containers;
  ResourceMap overhead;
  bool has_overhead = false;
};
struct Resource {  // framework.Resource
  int64_t milli_cpu = 0, memory = 0, ephemeral_storage = 0;
  ResourceMap scalar;
};
struct NodeInfo {
  std::string name;
  bool has_node = true;
  Resource allocatable, requested, non_zero_requested;
};
struct NodeScore {
  std::string name;
  int64_t score;
};

int64_t calculatePodResourceRequest(const Pod& pod, const std::string& resource) {  // resource_allocation.go:105-131
  int64_t pod_request = 0;
  for (const auto& c : pod.containers) {
    auto it = c.requests.find(resource);
    if (it != c.requests.end()) pod_request += it->second;
  }
  for (const auto& c : pod.init_containers) {
    auto it = c.requests.find(resource);
    if (it != c.requests.end() && pod_request < it->second) pod_request = it->second;
  }
  if (pod.has_overhead) {
    auto it = pod.overhead.find(resource);
    if (it != pod.overhead.end()) pod_request += it->second;
  }
  return pod_request;
}

void calculateResourceAllocatableRequest(const NodeInfo& ni, const Pod& pod, const std::string& resource,
                                         int64_t* alloc, int64_t* req) {  // :79-100
  int64_t pod_request = calculatePodResourceRequest(pod, resource);
  if (resource == "cpu") {
    *alloc = ni.allocatable.milli_cpu;
    *req = ni.non_zero_requested.milli_cpu + pod_request;
  } else if (resource == "memory") {
    *alloc = ni.allocatable.memory;
    *req = ni.non_zero_requested.memory + pod_request;
  } else if (resource == "ephemeral-storage") {
    *alloc = ni.allocatable.ephemeral_storage;
    *req = ni.requested.ephemeral_storage + pod_request;
  } else {
    auto it = ni.allocatable.scalar.find(resource);
    *alloc = it == ni.allocatable.scalar.end() ? 0 : it->second;
    auto jt = ni.requested.scalar.find(resource);
    *req = (jt == ni.requested.scalar.end() ? 0 : jt->second) + pod_request;
  }
}

struct Plugin {
  ResourceMap resource_to_weight;
  int mode;
  // resourceScorer closure: allocatable.go:117-128 (+ score :130-140)
  int64_t scorer(const ResourceMap& /*requested*/, const ResourceMap& allocable) const {
    int64_t node_score = 0, weight_sum = 0;
    for (const auto& kv : resource_to_weight) {
      auto it = allocable.find(kv.first);
      int64_t capacity = it == allocable.end() ? 0 : it->second;
      int64_t rs = mode == 0 ? orc_wrap_mul(-1, capacity) : (mode == 1 ? capacity : 0);
      node_score = orc_wrap_add(node_score, orc_wrap_mul(rs, kv.second));
      weight_sum = orc_wrap_add(weight_sum, kv.second);
    }
    return orc_go_div(node_score, weight_sum);
  }
  // resourceAllocationScorer.score: resource_allocation.go:49-76
  bool score(const Pod& pod, const NodeInfo& ni, int64_t* out) const {
    if (!ni.has_node) return false;
    ResourceMap requested(resource_to_weight.size()), allocatable(resource_to_weight.size());  // :60-61
    for (const auto& kv : resource_to_weight)
      calculateResourceAllocatableRequest(ni, pod, kv.first, &allocatable[kv.first], &requested[kv.first]);
    *out = scorer(requested, allocatable);
    return true;
  }
};

void normalize(std::vector<NodeScore>& scores) {  // allocatable.go:143-168
  int64_t highest = -INT64_MAX, lowest = INT64_MAX;
  for (const auto& s : scores) {
    if (s.score > highest) highest = s.score;
    if (s.score < lowest) lowest = s.score;
  }
  int64_t old_range = orc_wrap_sub(highest, lowest);
  for (auto& s : scores)
    s.score = old_range == 0 ? 0 : orc_go_div(orc_wrap_mul(orc_wrap_sub(s.score, lowest), 100), old_range);
}

}  // namespace

extern "C" void orc_gofaithful_alloc_batch(const int64_t* const* cols, const char* const* res_names, int R, int N,
                                           const int64_t* w, int mode, int P, const int64_t* pod_cpu_milli,
                                           const int64_t* pod_mem_bytes, const uint64_t* feasible, int words,
                                           int64_t* out, int pitch, int threads, double* compute_seconds) {
  Plugin pl;
  pl.mode = mode;
  for (int r = 0; r < R; ++r) pl.resource_to_weight[res_names[r]] = w[r];
  std::vector<NodeInfo> nodes(N);
  for (int n = 0; n < N; ++n) {
    nodes[n].name = "node-" + std::to_string(n);
    for (int r = 0; r < R; ++r) {
      std::string nm = res_names[r];
      if (nm == "cpu") nodes[n].allocatable.milli_cpu = cols[r][n];
      else if (nm == "memory") nodes[n].allocatable.memory = cols[r][n];
      else if (nm == "ephemeral-storage") nodes[n].allocatable.ephemeral_storage = cols[r][n];
      else nodes[n].allocatable.scalar[nm] = cols[r][n];
    }
  }
  // the NodeInfo list above is the scheduler's pre-existing snapshot: only the cycles below are timed
  const auto t_start = std::chrono::steady_clock::now();
  std::atomic<int> next(0);
  auto worker = [&]() {
    std::vector<NodeScore> list;
    std::vector<int> idx;
    for (;;) {
      int p = next.fetch_add(1);
      if (p >= P) return;
      Pod pod;  // two app containers splitting the pod's request (what calculatePodResourceRequest walks)
      pod.containers.resize(2);
      pod.containers[0].requests = {{"cpu", pod_cpu_milli[p] / 2}, {"memory", pod_mem_bytes[p] / 2}};
      pod.containers[1].requests = {{"cpu", pod_cpu_milli[p] - pod_cpu_milli[p] / 2},
                                    {"memory", pod_mem_bytes[p] - pod_mem_bytes[p] / 2}};
      list.clear();
      idx.clear();
      for (int n = 0; n < N; ++n) {
        if (feasible && !((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull)) continue;
        int64_t s;
        if (!pl.score(pod, nodes[n], &s)) continue;
        list.push_back(NodeScore{nodes[n].name, s});  // fwk.NodeScore{Name, Score}
        idx.push_back(n);
      }
      normalize(list);
      int64_t* row = out + (size_t)p * pitch;
      memset(row, 0, sizeof(int64_t) * (size_t)pitch);
      for (size_t i = 0; i < list.size(); ++i) row[idx[i]] = list[i].score;
    }
  };
  if (threads <= 1) {
    worker();
  } else {
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) ts.emplace_back(worker);
    for (auto& t : ts) t.join();
  }
  if (compute_seconds)
    *compute_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
}

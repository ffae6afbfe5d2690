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
// Two fan-out shapes:
//   orc_gofaithful_alloc_batch     one scheduling cycle (= one pod over all feasible nodes) per thread, cycles of
//                                  different pods on `threads` pinned threads -- the throughput shape, more generous
//                                  to the CPU than upstream (which schedules one pod at a time);
//   orc_gofaithful_alloc_cycles16  upstream's shape: pods one after the other, each cycle's Score calls fanned out
//                                  over a 16-worker Parallelizer in chunks of chunkSizeFor(n, 16) nodes
//                                  (k8s.io/kubernetes pkg/scheduler/framework/parallelize, not in the reference tree;
//                                  the reference's own harness mirrors it, targetloadpacking_test.go:386-405), then
//                                  NormalizeScore on the scheduling goroutine.
// Go's per-call maps hold <= 8 entries: one bucket, allocated from the goroutine's P-local cache without a global
// lock.  The restatement therefore keeps them as small inline maps (string keys, linear probe within the one
// bucket) instead of std::unordered_map -- whose node allocations serialise on the C allocator and made round 1's
// baseline swing 5x between boxes.
#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "oracle.h"

namespace {

// map[string]int64 with at most 8 entries (one Go map bucket)
struct ResourceMap {
  struct Entry {
    const std::string* key;
    int64_t value;
  };
  Entry e[8];
  int n = 0;
  ResourceMap() = default;
  explicit ResourceMap(size_t /*hint*/) {}
  ResourceMap(std::initializer_list<std::pair<const char*, int64_t>> init) {
    for (auto& kv : init) (*this)[intern(kv.first)] = kv.second;
  }
  static const std::string& intern(const char* s) {  // resource names live for the process (Go string constants)
    static const std::string cpu = "cpu", memory = "memory", eph = "ephemeral-storage";
    static thread_local std::vector<std::string*> others;
    if (cpu == s) return cpu;
    if (memory == s) return memory;
    if (eph == s) return eph;
    for (auto* o : others)
      if (*o == s) return *o;
    others.push_back(new std::string(s));
    return *others.back();
  }
  const Entry* find(const std::string& k) const {
    for (int i = 0; i < n; ++i)
      if (*e[i].key == k) return &e[i];  // string compare, as the Go map does after the hash matched
    return nullptr;
  }
  const Entry* end() const { return nullptr; }
  int64_t& operator[](const std::string& k) {
    for (int i = 0; i < n; ++i)
      if (*e[i].key == k) return e[i].value;
    e[n].key = &k;
    e[n].value = 0;
    return e[n++].value;
  }
  size_t size() const { return (size_t)n; }
  const Entry* begin() const { return e; }
  const Entry* endp() const { return e + n; }
};

void pin_to_cpu(int t) {
  const int ncpu = (int)std::thread::hardware_concurrency();
  if (ncpu <= 0) return;
  cpu_set_t set;
  CPU_ZERO(&set);
  CPU_SET(t % ncpu, &set);
  pthread_setaffinity_np(pthread_self(), sizeof(set), &set);  // best effort (a cgroup may forbid it)
}

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
    if (it) pod_request += it->value;
  }
  for (const auto& c : pod.init_containers) {
    auto it = c.requests.find(resource);
    if (it && pod_request < it->value) pod_request = it->value;
  }
  if (pod.has_overhead) {
    auto it = pod.overhead.find(resource);
    if (it) pod_request += it->value;
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
    *alloc = it ? it->value : 0;
    auto jt = ni.requested.scalar.find(resource);
    *req = (jt ? jt->value : 0) + pod_request;
  }
}

struct Plugin {
  ResourceMap resource_to_weight;
  int mode;
  // resourceScorer closure: allocatable.go:117-128 (+ score :130-140)
  int64_t scorer(const ResourceMap& /*requested*/, const ResourceMap& allocable) const {
    int64_t node_score = 0, weight_sum = 0;
    for (const auto* kv = resource_to_weight.begin(); kv != resource_to_weight.endp(); ++kv) {
      auto it = allocable.find(*kv->key);
      int64_t capacity = it ? it->value : 0;
      int64_t rs = mode == 0 ? orc_wrap_mul(-1, capacity) : (mode == 1 ? capacity : 0);
      node_score = orc_wrap_add(node_score, orc_wrap_mul(rs, kv->value));
      weight_sum = orc_wrap_add(weight_sum, kv->value);
    }
    return orc_go_div(node_score, weight_sum);
  }
  // resourceAllocationScorer.score: resource_allocation.go:49-76
  bool score(const Pod& pod, const NodeInfo& ni, int64_t* out) const {
    if (!ni.has_node) return false;
    ResourceMap requested(resource_to_weight.size()), allocatable(resource_to_weight.size());  // :60-61
    for (const auto* kv = resource_to_weight.begin(); kv != resource_to_weight.endp(); ++kv)
      calculateResourceAllocatableRequest(ni, pod, *kv->key, &allocatable[*kv->key], &requested[*kv->key]);
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

void build_snapshot(Plugin& pl, std::vector<NodeInfo>& nodes, const int64_t* const* cols, const char* const* res_names,
                    int R, int N, const int64_t* w) {
  for (int r = 0; r < R; ++r) pl.resource_to_weight[ResourceMap::intern(res_names[r])] = w[r];
  nodes.resize(N);
  for (int n = 0; n < N; ++n) {
    nodes[n].name = "node-" + std::to_string(n);
    for (int r = 0; r < R; ++r) {
      std::string nm = res_names[r];
      if (nm == "cpu") nodes[n].allocatable.milli_cpu = cols[r][n];
      else if (nm == "memory") nodes[n].allocatable.memory = cols[r][n];
      else if (nm == "ephemeral-storage") nodes[n].allocatable.ephemeral_storage = cols[r][n];
      else nodes[n].allocatable.scalar[ResourceMap::intern(res_names[r])] = cols[r][n];
    }
  }
}

Pod make_pod(int64_t cpu_milli, int64_t mem_bytes) {  // two app containers splitting the pod's request
  Pod pod;
  pod.containers.resize(2);
  pod.containers[0].requests = {{"cpu", cpu_milli / 2}, {"memory", mem_bytes / 2}};
  pod.containers[1].requests = {{"cpu", cpu_milli - cpu_milli / 2}, {"memory", mem_bytes - mem_bytes / 2}};
  return pod;
}

// upstream parallelize.chunkSizeFor(n, parallelism): sqrt(n), capped so that every worker gets a piece
int chunk_size_for(int n, int parallelism) {
  int s = (int)std::sqrt((double)n);
  const int r = n / parallelism + 1;
  if (s > r) s = r;
  return s < 1 ? 1 : s;
}

// A persistent pool standing in for goroutines (which cost ~1 us to start; an OS thread costs ~50 us): workers spin
// on a generation counter, take chunks from an atomic cursor, and the caller joins on a done counter.
struct Pool {
  int workers;
  std::vector<std::thread> threads;
  std::atomic<uint64_t> gen{0};
  std::atomic<int> cursor{0}, done{0};
  std::atomic<bool> stop{false};
  int pieces = 0, chunk = 1;
  const std::function<void(int, int)>* fn = nullptr;
  explicit Pool(int w) : workers(w) {
    for (int t = 0; t < w; ++t)
      threads.emplace_back([this, t] {
        pin_to_cpu(t);
        uint64_t seen = 0;
        for (;;) {
          while (gen.load(std::memory_order_acquire) == seen && !stop.load(std::memory_order_relaxed)) {
          }
          if (stop.load()) return;
          seen = gen.load();
          for (;;) {
            const int c0 = cursor.fetch_add(chunk);
            if (c0 >= pieces) break;
            (*fn)(c0, std::min(pieces, c0 + chunk));
          }
          done.fetch_add(1, std::memory_order_release);
        }
      });
  }
  void until(int n, const std::function<void(int, int)>& f) {  // Parallelizer.Until(ctx, n, doWorkPiece)
    pieces = n;
    chunk = chunk_size_for(n, workers);
    fn = &f;
    cursor.store(0);
    done.store(0);
    gen.fetch_add(1, std::memory_order_release);
    while (done.load(std::memory_order_acquire) < workers) {
    }
  }
  ~Pool() {
    stop.store(true);
    for (auto& t : threads) t.join();
  }
};

}  // namespace

// Upstream shape: pods scheduled one at a time; per cycle the Score calls of the feasible nodes are fanned out over
// `workers` (16 upstream), then NormalizeScore runs serially.  cycle_seconds[p] = wall time of pod p's cycle.
extern "C" void orc_gofaithful_alloc_cycles16(const int64_t* const* cols, const char* const* res_names, int R, int N,
                                              const int64_t* w, int mode, int P, const int64_t* pod_cpu_milli,
                                              const int64_t* pod_mem_bytes, const uint64_t* feasible, int words,
                                              int64_t* out, int pitch, int workers, double* cycle_seconds) {
  Plugin pl;
  pl.mode = mode;
  std::vector<NodeInfo> nodes;
  build_snapshot(pl, nodes, cols, res_names, R, N, w);
  Pool pool(workers < 1 ? 1 : workers);
  std::vector<int> feas_idx;
  std::vector<NodeScore> list;
  for (int p = 0; p < P; ++p) {
    const auto t0 = std::chrono::steady_clock::now();
    const Pod pod = make_pod(pod_cpu_milli[p], pod_mem_bytes[p]);
    feas_idx.clear();
    for (int n = 0; n < N; ++n)
      if (!feasible || ((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull)) feas_idx.push_back(n);
    list.assign(feas_idx.size(), NodeScore{std::string(), 0});
    const std::function<void(int, int)> piece = [&](int a, int b) {
      for (int i = a; i < b; ++i) {
        const NodeInfo& ni = nodes[feas_idx[i]];
        int64_t s = 0;
        pl.score(pod, ni, &s);
        list[i] = NodeScore{ni.name, s};  // pluginToNodeScores[pl.Name()][index]
      }
    };
    pool.until((int)feas_idx.size(), piece);
    normalize(list);
    int64_t* row = out + (size_t)p * pitch;
    memset(row, 0, sizeof(int64_t) * (size_t)pitch);
    for (size_t i = 0; i < list.size(); ++i) row[feas_idx[i]] = list[i].score;
    if (cycle_seconds) cycle_seconds[p] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
}

extern "C" void orc_gofaithful_alloc_batch(const int64_t* const* cols, const char* const* res_names, int R, int N,
                                           const int64_t* w, int mode, int P, const int64_t* pod_cpu_milli,
                                           const int64_t* pod_mem_bytes, const uint64_t* feasible, int words,
                                           int64_t* out, int pitch, int threads, double* compute_seconds) {
  Plugin pl;
  pl.mode = mode;
  std::vector<NodeInfo> nodes;
  build_snapshot(pl, nodes, cols, res_names, R, N, w);
  // the NodeInfo list above is the scheduler's pre-existing snapshot: only the cycles below are timed
  const auto t_start = std::chrono::steady_clock::now();
  std::atomic<int> next(0);
  auto worker = [&](int t) {
    if (threads > 1) pin_to_cpu(t);
    std::vector<NodeScore> list;
    std::vector<int> idx;
    list.reserve((size_t)N);
    idx.reserve((size_t)N);
    for (;;) {
      int p = next.fetch_add(1);
      if (p >= P) return;
      const Pod pod = make_pod(pod_cpu_milli[p], pod_mem_bytes[p]);  // what calculatePodResourceRequest walks
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
    worker(0);
  } else {
    std::vector<std::thread> ts;
    for (int t = 0; t < threads; ++t) ts.emplace_back(worker, t);
    for (auto& t : ts) t.join();
  }
  if (compute_seconds)
    *compute_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
}

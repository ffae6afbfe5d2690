// C++ mirror of the reference's plugin interface for the hot path (stand-in for the Go shim of
// INTEGRATION.md §3).  Same plugin names, method names, argument meaning, status codes and
// messages as the Go plugins; every score / filter verdict comes from libb200sched.so through the
// C-ABI (include/b200sched.h) — there is no arithmetic fallback here: if the engine fails, the
// methods return fwk.Error with the engine's message.
//
//   Allocatable                  pkg/noderesources/allocatable.go:45,63,79,143
//   TargetLoadPacking            pkg/trimaran/targetloadpacking/targetloadpacking.go:44,65,107,193
//   LoadVariationRiskBalancing   pkg/trimaran/loadvariationriskbalancing/loadvariationriskbalancing.go:41,56,84,135
//   TopologyMatch                pkg/noderesourcetopology/plugin.go:42,91  filter.go:176  score.go:62
//   NetworkOverhead              pkg/networkaware/networkoverhead/networkoverhead.go:49,119,174,326,362,389
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/b200sched.h"
#include "objects.hpp"

namespace b200host {

// RAII over one engine context (one GPU, one shard).
class Engine {
 public:
  explicit Engine(int device);
  ~Engine();
  Engine(const Engine&) = delete;
  b200s_ctx* ctx() const { return ctx_; }
  std::string LastError() const;
  void Check(int rc, const char* what) const;  // throws std::runtime_error with the engine message

 private:
  b200s_ctx* ctx_ = nullptr;
};

struct ScheduledPodInfo {  // trimaran.podInfo, handler.go:41-45
  int64_t timestamp_unix = 0;
  std::shared_ptr<Pod> pod;
};

// fwk.Handle as far as the five plugins use it: the cycle's snapshot, the pod lister, and the
// CR / metrics sources the reference reaches through informers and controller-runtime clients.
struct Handle {
  int device = 0;
  uint64_t generation = 1;                      // bump when node_infos / CRs / metrics change
  std::vector<NodeInfo> node_infos;             // SnapshotSharedLister().NodeInfos().List()
  std::vector<std::shared_ptr<Pod>> pods;       // pod lister
  std::shared_ptr<WatcherMetrics> metrics;      // Collector's last fetch (collector.go:102-107); null = never fetched
  std::map<std::string, std::vector<ScheduledPodInfo>> scheduled_pods_cache;  // PodAssignEventHandler (handler.go:47)
  std::map<std::string, std::shared_ptr<NodeResourceTopology>> nrts;          // nrtcache.Passthrough by node name
  std::map<std::string, bool> nrt_not_fresh;                                   // CachedNRTInfo.Fresh == false
  std::map<std::string, std::shared_ptr<AppGroup>> app_groups;
  std::map<std::string, std::shared_ptr<NetworkTopology>> network_topologies;
  // Upstream's cache refreshes its snapshot node by node (NodeInfo.Generation); TouchNode records which node an
  // event changed so that the plugins rewrite that row of the resident columns instead of re-flattening all of
  // them.  Touch() = anything else changed (node list, metrics fetch, CRs): everything is re-flattened.
  void Touch() {
    ++generation;
    node_log_.clear();
    log_base_ = generation;
  }
  void TouchNode(int32_t i) {
    ++generation;
    if (node_log_.size() >= (1u << 16)) {  // nobody will replay a log this long: start over (= Touch)
      node_log_.clear();
      log_base_ = generation;
      return;
    }
    node_log_.emplace_back(generation, i);
  }
  // True when every change after generation g is a logged single-node change; *out = those nodes (deduplicated).
  bool NodesChangedSince(uint64_t g, std::vector<int32_t>* out) const;

 private:
  std::vector<std::pair<uint64_t, int32_t>> node_log_;
  uint64_t log_base_ = 1;
};

struct CycleState {
  std::map<std::string, std::shared_ptr<void>> data;
};

// Per-cycle result of one engine call for one pod.
struct CycleResult {
  std::map<std::string, int32_t> index;  // node name -> column
  std::vector<uint8_t> scores;           // u8 transport
  std::vector<uint64_t> feasible;
  std::vector<uint8_t> reasons;
  std::string engine_error;              // non-empty: the engine call failed
  std::set<int32_t> unsupported_nodes;   // NodeResourceTopologyMatch: columns whose NRT is outside the dense encoding
  bool score_equally = false;            // NetworkOverhead
  std::vector<int64_t> satisfied, violated;  // diagnostics for NetworkOverhead's message (host-side recount)
};

// ---------------------------------------------------------------- NodeResourcesAllocatable
struct ResourceSpec {
  std::string name;
  int64_t weight = 0;
};
struct NodeResourcesAllocatableArgs {
  std::string mode;  // "Least" | "Most" | "" (defaults to Least, allocatable.go:90-92)
  std::vector<ResourceSpec> resources;
};

class Allocatable {
 public:
  static constexpr const char* AllocatableName = "NodeResourcesAllocatable";
  // NewAllocatable: nil args -> defaults {cpu: 1<<20, memory: 1}, Least.  Throws std::invalid_argument with the
  // reference's validation message for non-positive weights / unknown mode.
  static std::unique_ptr<Allocatable> New(const NodeResourcesAllocatableArgs* args, std::shared_ptr<Handle> h);
  std::string Name() const { return AllocatableName; }
  // upstream calls PreScore with the feasible nodes; NormalizeScore then sees exactly that list
  Status PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  Status NormalizeScore(CycleState& state, const Pod& pod, std::vector<NodeScore>& scores);

 private:
  Allocatable() = default;
  void EnsureSnapshot();
  // rewrite only the rows of the nodes touched since snap_gen_; false = re-flatten everything
  bool PatchSnapshot();
  int64_t patched_rows_ = 0;

 public:
  int64_t PatchedRows() const { return patched_rows_; }  // rows rewritten through b200s_snapshot_patch_* so far
  // test hook (fault injection): leaves a patch open on the engine, as a patch that failed half-way would
  int DebugLeavePatchOpen() { return b200s_snapshot_patch_begin(eng_->ctx(), 0); }

 private:
  std::shared_ptr<CycleResult> Run(const Pod& pod, const std::vector<NodeInfo>* feasible);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  std::vector<ResourceSpec> res_;
  int mode_ = 0;
  uint64_t snap_gen_ = 0;
  std::map<std::string, int32_t> index_;
  int32_t n_ = 0, npad_ = 0;
};

// ---------------------------------------------------------------- Trimaran
struct TargetLoadPackingArgs {
  int64_t target_utilization = 40;                 // defaults.go:48-52
  int64_t default_requests_cpu_milli = 1000;
  std::string default_requests_multiplier = "1.5";
};

class TargetLoadPacking {
 public:
  static constexpr const char* Name_ = "TargetLoadPacking";
  static std::unique_ptr<TargetLoadPacking> New(const TargetLoadPackingArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  Status NormalizeScore(CycleState&, const Pod&, std::vector<NodeScore>&) { return {}; }  // no-op, :193

 private:
  TargetLoadPacking() = default;
  void EnsureSnapshot();
  bool PatchSnapshot();
  int64_t patched_rows_ = 0;

 public:
  int64_t PatchedRows() const { return patched_rows_; }  // rows rewritten through b200s_snapshot_patch_* so far

 private:
  struct Row {
    double util = 0;
    int64_t cap = 0, missing = 0;
    uint8_t flags = 0;
  };
  Row FlattenRow(const NodeInfo& ni) const;
  std::shared_ptr<CycleResult> Run(const Pod& pod);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  TargetLoadPackingArgs args_;
  double multiplier_ = 1.5;
  uint64_t snap_gen_ = 0;
  std::map<std::string, int32_t> index_;
  int32_t n_ = 0, npad_ = 0;
};

struct LoadVariationRiskBalancingArgs {
  double safe_variance_margin = 1.0, safe_variance_sensitivity = 1.0;  // defaults.go:62-64
};

class LoadVariationRiskBalancing {
 public:
  static constexpr const char* Name_ = "LoadVariationRiskBalancing";
  static std::unique_ptr<LoadVariationRiskBalancing> New(const LoadVariationRiskBalancingArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  Status NormalizeScore(CycleState&, const Pod&, std::vector<NodeScore>&) { return {}; }  // no-op, :135

 private:
  LoadVariationRiskBalancing() = default;
  void EnsureSnapshot();
  bool PatchSnapshot();
  int64_t patched_rows_ = 0;

 public:
  int64_t PatchedRows() const { return patched_rows_; }  // rows rewritten through b200s_snapshot_patch_* so far

 private:
  struct Row {
    double ca = 0, cs = 0, ma = 0, ms = 0;
    int64_t acpu = 0, amem = 0;
    uint8_t flags = 0;
  };
  Row FlattenRow(const NodeInfo& ni) const;
  std::shared_ptr<CycleResult> Run(const Pod& pod);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  LoadVariationRiskBalancingArgs args_;
  uint64_t snap_gen_ = 0;
  std::map<std::string, int32_t> index_;
  int32_t n_ = 0, npad_ = 0;
};

// Trimaran Peaks (pkg/trimaran/peaks/peaks.go).  As with NodeResourcesAllocatable, Score returns the value already
// normalised over the list PreScore saw and NormalizeScore is a no-op: the pipeline Score -> NormalizeScore yields
// exactly what the reference's pipeline yields (its raw Score values are only an intermediate).
struct PowerModel {
  double k0 = 0, k1 = 0, k2 = 0;  // power = k0 + k1 * e^(k2 * utilisation), apis/config/types.go:301-307
};
struct PeaksArgs {
  std::map<std::string, PowerModel> node_power_model;
};
class Peaks {
 public:
  static constexpr const char* Name_ = "Peaks";
  static std::unique_ptr<Peaks> New(const PeaksArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  Status NormalizeScore(CycleState&, const Pod&, std::vector<NodeScore>&) { return {}; }

 private:
  Peaks() = default;
  void EnsureSnapshot();
  std::shared_ptr<CycleResult> Run(const Pod& pod, const std::vector<NodeInfo>* feasible);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  PeaksArgs args_;
  uint64_t snap_gen_ = 0;
  std::map<std::string, int32_t> index_;
  int32_t n_ = 0, npad_ = 0;
};

// Trimaran LowRiskOverCommitment (pkg/trimaran/lowriskovercommitment/lowriskovercommitment.go)
struct LowRiskOverCommitmentArgs {
  int64_t smoothing_window_size = 5;                      // defaults.go:70
  double risk_limit_weight_cpu = 0.5, risk_limit_weight_memory = 0.5;  // defaults.go:72-77
};
class LowRiskOverCommitment {
 public:
  static constexpr const char* Name_ = "LowRiskOverCommitment";
  static std::unique_ptr<LowRiskOverCommitment> New(const LowRiskOverCommitmentArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  Status NormalizeScore(CycleState&, const Pod&, std::vector<NodeScore>&) { return {}; }  // :150-152

 private:
  LowRiskOverCommitment() = default;
  void EnsureSnapshot();
  std::shared_ptr<CycleResult> Run(const Pod& pod);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  LowRiskOverCommitmentArgs args_;
  uint64_t snap_gen_ = 0;
  std::map<std::string, int32_t> index_;
  int32_t n_ = 0, npad_ = 0;
};

// ---------------------------------------------------------------- NodeResourceTopologyMatch
struct NodeResourceTopologyMatchArgs {
  std::string scoring_strategy = "LeastAllocated";  // defaults.go:84-87
  std::vector<ResourceSpec> resources = {{"cpu", 1}, {"memory", 1}};
};

class TopologyMatch {
 public:
  static constexpr const char* Name_ = "NodeResourceTopologyMatch";
  static std::unique_ptr<TopologyMatch> New(const NodeResourceTopologyMatchArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status Filter(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);

 private:
  TopologyMatch() = default;
  std::shared_ptr<CycleResult> Run(CycleState& state, const Pod& pod);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  NodeResourceTopologyMatchArgs args_;
  int strategy_ = B200S_NRT_LEAST_ALLOCATED;
};

// ---------------------------------------------------------------- NetworkOverhead
struct NetworkOverheadArgs {
  std::vector<std::string> namespaces = {"default"};  // defaults.go:97-99
  std::string weights_name = "UserDefined";
  std::string network_topology_name = "nt-default";
};

class NetworkOverhead {
 public:
  static constexpr const char* Name_ = "NetworkOverhead";
  static std::unique_ptr<NetworkOverhead> New(const NetworkOverheadArgs& args, std::shared_ptr<Handle> h);
  std::string Name() const { return Name_; }
  Status PreFilter(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes);
  Status Filter(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  std::pair<int64_t, Status> Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo);
  // The engine normalises over the feasible list it is given here (upstream passes the nodes that survived
  // every filter), and rewrites `scores` — the one place where the list itself is an input.
  Status NormalizeScore(CycleState& state, const Pod& pod, std::vector<NodeScore>& scores);

 private:
  NetworkOverhead() = default;
  std::shared_ptr<CycleResult> Run(CycleState& state, const Pod& pod, const std::vector<std::string>* feasible_names,
                                   bool raw_scores);
  std::shared_ptr<Handle> h_;
  std::unique_ptr<Engine> eng_;
  NetworkOverheadArgs args_;
};

}  // namespace b200host

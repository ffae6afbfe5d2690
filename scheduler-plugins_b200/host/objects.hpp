// Host-side object model + the reference's host-side rules (C++ stand-in for the Go shim; the
// image has no Go toolchain).  Names follow the Kubernetes API / the reference so that tests read
// like the reference's own.  No scoring arithmetic lives here — only flattening rules:
//   quantities (apimachinery resource.Quantity), QoS (v1qos.GetPodQOS), resource-name predicates
//   (v1helper), GetPodEffectiveRequest (pkg/util/resource.go:51-85), PredictUtilisation
//   (targetloadpacking.go:198-205), GetResourceRequested (resourcestats.go:110-146),
//   GetResourceData (resourcestats.go:89-107), createNUMANodeList / TopologyManager
//   (pluginhelpers.go:105-161, nodeconfig/topologymanager.go:78-161).
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200host {

// ---- resource.Quantity as exact milli-units --------------------------------------------------
// "500m" = 500, "2" = 2000, "1Gi" = 1073741824000.  Finer than 1m is outside the dense encoding.
int64_t ParseQuantity(const std::string& s);  // throws std::invalid_argument
inline int64_t QuantityValue(int64_t milli) {  // Quantity.Value(): rounded up
  return milli >= 0 ? (milli + 999) / 1000 : -((-milli) / 1000);
}
using ResourceList = std::map<std::string, int64_t>;  // name -> milli

constexpr const char* ResourceCPU = "cpu";
constexpr const char* ResourceMemory = "memory";
constexpr const char* ResourcePods = "pods";
constexpr const char* ResourceEphemeralStorage = "ephemeral-storage";
constexpr const char* LabelTopologyRegion = "topology.kubernetes.io/region";
constexpr const char* LabelTopologyZone = "topology.kubernetes.io/zone";
constexpr const char* AppGroupLabel = "appgroup.diktyo.x-k8s.io";
constexpr const char* AppGroupSelectorLabel = "appgroup.diktyo.x-k8s.io.workload";

struct Container {
  std::string name;
  ResourceList requests, limits;
  bool restart_always = false;  // init container with RestartPolicy Always = sidecar (pkg/util/sidecar.go:25)
};

struct Pod {
  std::string name, uid, node_name;
  std::map<std::string, std::string> labels;
  std::vector<Container> init_containers, containers;
  bool has_overhead = false;
  ResourceList overhead;
};

struct Node {
  std::string name;
  std::map<std::string, std::string> labels;
  ResourceList capacity, allocatable;
};

struct NodeInfo {
  std::shared_ptr<Node> node;  // Node() == nil is representable
  std::vector<std::shared_ptr<Pod>> pods;  // GetPods(): the pods already on the node
  const Node* GetNode() const { return node.get(); }
};

enum class Code { Success = 0, Error = 1, Unschedulable = 2 };
struct Status {
  Code code = Code::Success;
  std::string message;
  bool IsSuccess() const { return code == Code::Success; }
};
struct NodeScore {
  std::string name;
  int64_t score = 0;
};

// ---- load-watcher metrics (paypal/load-watcher types; constants only) -------------------------
struct Metric {
  std::string type;      // "CPU" | "Memory"
  std::string op;        // "AVG" | "STD" | "Latest" | ""
  double value = 0;
};
struct NodeMetrics {
  std::vector<Metric> metrics;
};
struct WatcherMetrics {
  int64_t window_end = 0;
  bool has_map = false;  // Data.NodeMetricsMap != nil
  std::map<std::string, NodeMetrics> node_metrics;
};

// ---- NodeResourceTopology CR -------------------------------------------------------------------
struct ZoneResource {
  int64_t capacity = 0, available = 0;
};
struct Zone {
  std::string name, type = "Node";
  std::map<std::string, ZoneResource> resources;
  std::map<std::string, int64_t> costs;
};
struct NodeResourceTopology {
  std::string name;
  std::vector<std::string> topology_policies;
  std::map<std::string, std::string> attributes;
  std::vector<Zone> zones;
};

// ---- AppGroup / NetworkTopology CRs (diktyo-io APIs; fields the plugin reads) -----------------
struct DependencyInfo {
  std::string selector;
  int64_t max_network_cost = 0;
};
struct AppGroupWorkload {
  std::string selector;
  std::vector<DependencyInfo> dependencies;
};
struct AppGroup {
  std::string name;
  std::vector<AppGroupWorkload> workloads;
};
struct CostInfo {
  std::string destination;
  int64_t network_cost = 0;
};
struct OriginInfo {
  std::string origin;
  std::vector<CostInfo> cost_list;
};
struct TopologyInfo {
  std::string topology_key;
  std::vector<OriginInfo> origin_list;
};
struct WeightInfo {
  std::string name;
  std::vector<TopologyInfo> topology_list;
};
struct NetworkTopology {
  std::string name;
  std::vector<WeightInfo> weights;
};

// ---- rules ------------------------------------------------------------------------------------
bool IsNativeResource(const std::string& n);
bool IsHugePageResourceName(const std::string& n);
bool IsNUMAAffineResource(const std::string& n);   // numaresources.go:120-135
bool IsHostLevelResource(const std::string& n);    // numaresources.go:105-118
bool IsScalarResourceName(const std::string& n);   // schedutil.IsScalarResourceName [upstream]

enum class QOS { Guaranteed = 0, Burstable = 1, BestEffort = 2 };
QOS GetPodQOS(const Pod& p);
bool IncludeNonNative(const Pod& p);                       // resourcerequests/exclusive.go:26-41
ResourceList GetPodEffectiveRequest(const Pod& p);         // pkg/util/resource.go:51-85
int64_t PredictUtilisation(const Container& c, int64_t default_milli, double multiplier);
int64_t PodPredictedCPU(const Pod& p, int64_t default_milli, double multiplier);  // targetloadpacking.go:122-129
void GetResourceRequested(const Pod& p, int64_t* cpu_milli, int64_t* mem_bytes);  // resourcestats.go:110-146
void GetResourceLimits(const Pod& p, int64_t* cpu_milli, int64_t* mem_bytes);     // resourcestats.go:117-122
// resource.GetResourceRequestQuantity(pod, cpu).MilliValue() [k8s.io/kubernetes/pkg/api/v1/resource, upstream]:
// sum of the app containers' requests, raised to the largest init container request, plus the pod overhead when
// the total is non-zero (peaks.go:113-114)
int64_t GetResourceRequestQuantityCPU(const Pod& p);
// GetNodeRequestsAndLimits (resourcestats.go:160-228) WITHOUT the pending pod and before the capacity cap: requests
// and limits summed over the pods already on the node, each pod's limits raised to its requests (SetMaxLimits).
// out = {request cpu milli, request memory bytes, limit cpu milli, limit memory bytes}
void NodeRequestsAndLimitsOfRunningPods(const NodeInfo& ni, int64_t out[4]);
// GetResourceData: resourcestats.go:89-107
void GetResourceData(const std::vector<Metric>& m, const std::string& type, double* avg, double* std_, bool* valid);

struct TopologyManager {  // nodeconfig/topologymanager.go:64-86
  std::string scope = "container", policy = "none";
  int max_numa_nodes = 8;
};
TopologyManager TopologyManagerFromNodeResourceTopology(const NodeResourceTopology& nrt);

}  // namespace b200host

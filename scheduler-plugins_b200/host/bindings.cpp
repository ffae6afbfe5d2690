// pybind11 bindings of the C++ host mirror — test glue only (the Go shim is the real consumer of
// this interface; pytest drives it so the parity tests read like the reference's table tests).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "plugins.hpp"
#include "nrt_scalar.hpp"

namespace py = pybind11;
using namespace b200host;

static ResourceList ToRL(const std::map<std::string, std::string>& m) {
  ResourceList r;
  for (auto& [k, v] : m) r[k] = ParseQuantity(v);
  return r;
}

PYBIND11_MODULE(_b200host, m) {
  m.doc() = "C++ mirror of the scheduler-plugins hot-path plugins over libb200sched.so";
  m.def("parse_quantity", &ParseQuantity);
  m.def("resource_list", &ToRL, "Kubernetes quantity strings -> milli-unit ResourceList");
  m.def("is_host_level_resource", &IsHostLevelResource);
  m.def("is_numa_affine_resource", &IsNUMAAffineResource);
  m.def("get_resource_requested", [](const Pod& p) {
    int64_t c, mm;
    GetResourceRequested(p, &c, &mm);
    return std::make_pair(c, mm);
  });
  m.def("get_resource_limits", [](const Pod& p) {
    int64_t c, mm;
    GetResourceLimits(p, &c, &mm);
    return std::make_pair(c, mm);
  });
  m.def("get_resource_request_quantity_cpu", &GetResourceRequestQuantityCPU);
  m.def("node_requests_and_limits_of_running_pods", [](const NodeInfo& ni) {
    int64_t o[4];
    NodeRequestsAndLimitsOfRunningPods(ni, o);
    return std::vector<int64_t>(o, o + 4);
  });
  m.def("pod_qos", [](const Pod& p) { return (int)GetPodQOS(p); });
  m.def("pod_effective_request", &GetPodEffectiveRequest);
  m.def("pod_predicted_cpu", &PodPredictedCPU);

  py::enum_<Code>(m, "Code").value("Success", Code::Success).value("Error", Code::Error).value("Unschedulable", Code::Unschedulable);
  py::class_<Status>(m, "Status")
      .def(py::init<>())
      .def_readwrite("code", &Status::code)
      .def_readwrite("message", &Status::message)
      .def("is_success", &Status::IsSuccess);
  py::class_<NodeScore>(m, "NodeScore")
      .def(py::init([](std::string n, int64_t s) { return NodeScore{std::move(n), s}; }), py::arg("name"), py::arg("score") = 0)
      .def_readwrite("name", &NodeScore::name)
      .def_readwrite("score", &NodeScore::score);
  py::class_<Container>(m, "Container")
      .def(py::init<>())
      .def_readwrite("name", &Container::name)
      .def_readwrite("requests", &Container::requests)
      .def_readwrite("limits", &Container::limits)
      .def_readwrite("restart_always", &Container::restart_always);
  py::class_<Pod, std::shared_ptr<Pod>>(m, "Pod")
      .def(py::init<>())
      .def_readwrite("name", &Pod::name)
      .def_readwrite("uid", &Pod::uid)
      .def_readwrite("node_name", &Pod::node_name)
      .def_readwrite("labels", &Pod::labels)
      .def_readwrite("init_containers", &Pod::init_containers)
      .def_readwrite("containers", &Pod::containers)
      .def_readwrite("has_overhead", &Pod::has_overhead)
      .def_readwrite("overhead", &Pod::overhead);
  py::class_<Node, std::shared_ptr<Node>>(m, "Node")
      .def(py::init<>())
      .def_readwrite("name", &Node::name)
      .def_readwrite("labels", &Node::labels)
      .def_readwrite("capacity", &Node::capacity)
      .def_readwrite("allocatable", &Node::allocatable);
  py::class_<NodeInfo>(m, "NodeInfo")
      .def(py::init<>())
      .def(py::init([](std::shared_ptr<Node> n) { return NodeInfo{std::move(n)}; }))
      .def_readwrite("node", &NodeInfo::node)
      .def_readwrite("pods", &NodeInfo::pods);
  py::class_<Metric>(m, "Metric")
      .def(py::init([](std::string t, std::string o, double v) { return Metric{std::move(t), std::move(o), v}; }))
      .def_readwrite("type", &Metric::type)
      .def_readwrite("op", &Metric::op)
      .def_readwrite("value", &Metric::value);
  py::class_<NodeMetrics>(m, "NodeMetrics").def(py::init<>()).def_readwrite("metrics", &NodeMetrics::metrics);
  py::class_<WatcherMetrics, std::shared_ptr<WatcherMetrics>>(m, "WatcherMetrics")
      .def(py::init<>())
      .def_readwrite("window_end", &WatcherMetrics::window_end)
      .def_readwrite("has_map", &WatcherMetrics::has_map)
      .def_readwrite("node_metrics", &WatcherMetrics::node_metrics);
  py::class_<ZoneResource>(m, "ZoneResource")
      .def(py::init([](int64_t c, int64_t a) { return ZoneResource{c, a}; }))
      .def_readwrite("capacity", &ZoneResource::capacity)
      .def_readwrite("available", &ZoneResource::available);
  py::class_<Zone>(m, "Zone")
      .def(py::init<>())
      .def_readwrite("name", &Zone::name)
      .def_readwrite("type", &Zone::type)
      .def_readwrite("resources", &Zone::resources)
      .def_readwrite("costs", &Zone::costs);
  py::class_<NodeResourceTopology, std::shared_ptr<NodeResourceTopology>>(m, "NodeResourceTopology")
      .def(py::init<>())
      .def_readwrite("name", &NodeResourceTopology::name)
      .def_readwrite("topology_policies", &NodeResourceTopology::topology_policies)
      .def_readwrite("attributes", &NodeResourceTopology::attributes)
      .def_readwrite("zones", &NodeResourceTopology::zones);
  py::class_<DependencyInfo>(m, "DependencyInfo")
      .def(py::init([](std::string s, int64_t c) { return DependencyInfo{std::move(s), c}; }), py::arg("selector"),
           py::arg("max_network_cost") = 0);
  py::class_<AppGroupWorkload>(m, "AppGroupWorkload")
      .def(py::init([](std::string s, std::vector<DependencyInfo> d) { return AppGroupWorkload{std::move(s), std::move(d)}; }));
  py::class_<AppGroup, std::shared_ptr<AppGroup>>(m, "AppGroup")
      .def(py::init([](std::string n, std::vector<AppGroupWorkload> w) {
        auto a = std::make_shared<AppGroup>();
        a->name = std::move(n);
        a->workloads = std::move(w);
        return a;
      }));
  py::class_<CostInfo>(m, "CostInfo").def(py::init([](std::string d, int64_t c) { return CostInfo{std::move(d), c}; }));
  py::class_<OriginInfo>(m, "OriginInfo")
      .def(py::init([](std::string o, std::vector<CostInfo> c) { return OriginInfo{std::move(o), std::move(c)}; }));
  py::class_<TopologyInfo>(m, "TopologyInfo")
      .def(py::init([](std::string k, std::vector<OriginInfo> o) { return TopologyInfo{std::move(k), std::move(o)}; }));
  py::class_<WeightInfo>(m, "WeightInfo")
      .def(py::init([](std::string n, std::vector<TopologyInfo> t) { return WeightInfo{std::move(n), std::move(t)}; }));
  py::class_<NetworkTopology, std::shared_ptr<NetworkTopology>>(m, "NetworkTopology")
      .def(py::init([](std::string n, std::vector<WeightInfo> w) {
        auto t = std::make_shared<NetworkTopology>();
        t->name = std::move(n);
        t->weights = std::move(w);
        return t;
      }));
  py::class_<ScheduledPodInfo>(m, "ScheduledPodInfo")
      .def(py::init([](int64_t ts, std::shared_ptr<Pod> p) { return ScheduledPodInfo{ts, std::move(p)}; }));
  py::class_<Handle, std::shared_ptr<Handle>>(m, "Handle")
      .def(py::init([]() { return std::make_shared<Handle>(); }))
      .def_readwrite("device", &Handle::device)
      .def_readwrite("generation", &Handle::generation)
      .def_readwrite("node_infos", &Handle::node_infos)
      .def_readwrite("pods", &Handle::pods)
      .def_readwrite("metrics", &Handle::metrics)
      .def_readwrite("scheduled_pods_cache", &Handle::scheduled_pods_cache)
      .def_readwrite("nrts", &Handle::nrts)
      .def_readwrite("nrt_not_fresh", &Handle::nrt_not_fresh)
      .def_readwrite("app_groups", &Handle::app_groups)
      .def_readwrite("network_topologies", &Handle::network_topologies)
      .def("touch", &Handle::Touch)
      .def("touch_node", &Handle::TouchNode)
      .def("nodes_changed_since", [](const Handle& h, uint64_t g) -> py::object {
        std::vector<int32_t> idx;
        if (!h.NodesChangedSince(g, &idx)) return py::none();
        return py::cast(idx);
      });
  py::class_<CycleState>(m, "CycleState").def(py::init<>());

  py::class_<ResourceSpec>(m, "ResourceSpec")
      .def(py::init([](std::string n, int64_t w) { return ResourceSpec{std::move(n), w}; }));
  py::class_<NodeResourcesAllocatableArgs>(m, "NodeResourcesAllocatableArgs")
      .def(py::init<>())
      .def_readwrite("mode", &NodeResourcesAllocatableArgs::mode)
      .def_readwrite("resources", &NodeResourcesAllocatableArgs::resources);
  py::class_<Allocatable>(m, "Allocatable")
      .def_static("new", [](const NodeResourcesAllocatableArgs* a, std::shared_ptr<Handle> h) { return Allocatable::New(a, std::move(h)); },
                  py::arg("args").none(true), py::arg("handle"))
      .def("name", &Allocatable::Name)
      .def("pre_score", &Allocatable::PreScore)
      .def("score", &Allocatable::Score)
      .def("patched_rows", &Allocatable::PatchedRows)
      .def("debug_leave_patch_open", &Allocatable::DebugLeavePatchOpen)
      .def("normalize_score", [](Allocatable& a, CycleState& s, const Pod& p, std::vector<NodeScore> l) {
        Status st = a.NormalizeScore(s, p, l);
        return std::make_pair(st, l);
      });
  py::class_<TargetLoadPackingArgs>(m, "TargetLoadPackingArgs")
      .def(py::init<>())
      .def_readwrite("target_utilization", &TargetLoadPackingArgs::target_utilization)
      .def_readwrite("default_requests_cpu_milli", &TargetLoadPackingArgs::default_requests_cpu_milli)
      .def_readwrite("default_requests_multiplier", &TargetLoadPackingArgs::default_requests_multiplier);
  py::class_<TargetLoadPacking>(m, "TargetLoadPacking")
      .def_static("new", &TargetLoadPacking::New)
      .def("name", &TargetLoadPacking::Name)
      .def("pre_score", &TargetLoadPacking::PreScore)
      .def("score", &TargetLoadPacking::Score)
      .def("patched_rows", &TargetLoadPacking::PatchedRows);
  py::class_<LoadVariationRiskBalancingArgs>(m, "LoadVariationRiskBalancingArgs")
      .def(py::init<>())
      .def_readwrite("safe_variance_margin", &LoadVariationRiskBalancingArgs::safe_variance_margin)
      .def_readwrite("safe_variance_sensitivity", &LoadVariationRiskBalancingArgs::safe_variance_sensitivity);
  py::class_<LoadVariationRiskBalancing>(m, "LoadVariationRiskBalancing")
      .def_static("new", &LoadVariationRiskBalancing::New)
      .def("name", &LoadVariationRiskBalancing::Name)
      .def("pre_score", &LoadVariationRiskBalancing::PreScore)
      .def("score", &LoadVariationRiskBalancing::Score)
      .def("patched_rows", &LoadVariationRiskBalancing::PatchedRows);
  py::class_<PowerModel>(m, "PowerModel")
      .def(py::init([](double k0, double k1, double k2) { return PowerModel{k0, k1, k2}; }));
  py::class_<PeaksArgs>(m, "PeaksArgs").def(py::init<>()).def_readwrite("node_power_model", &PeaksArgs::node_power_model);
  py::class_<Peaks>(m, "Peaks")
      .def_static("new", &Peaks::New)
      .def("name", &Peaks::Name)
      .def("pre_score", &Peaks::PreScore)
      .def("score", &Peaks::Score);
  py::class_<LowRiskOverCommitmentArgs>(m, "LowRiskOverCommitmentArgs")
      .def(py::init<>())
      .def_readwrite("smoothing_window_size", &LowRiskOverCommitmentArgs::smoothing_window_size)
      .def_readwrite("risk_limit_weight_cpu", &LowRiskOverCommitmentArgs::risk_limit_weight_cpu)
      .def_readwrite("risk_limit_weight_memory", &LowRiskOverCommitmentArgs::risk_limit_weight_memory);
  py::class_<LowRiskOverCommitment>(m, "LowRiskOverCommitment")
      .def_static("new", &LowRiskOverCommitment::New)
      .def("name", &LowRiskOverCommitment::Name)
      .def("pre_score", &LowRiskOverCommitment::PreScore)
      .def("score", &LowRiskOverCommitment::Score);
  py::class_<NodeResourceTopologyMatchArgs>(m, "NodeResourceTopologyMatchArgs")
      .def(py::init<>())
      .def_readwrite("scoring_strategy", &NodeResourceTopologyMatchArgs::scoring_strategy)
      .def_readwrite("resources", &NodeResourceTopologyMatchArgs::resources);
  // the scalar path for shapes outside the dense encoding (host/nrt_scalar.cpp): exposed so that the reference's
  // TestNUMANodesRequired vectors with unsorted / sparse NUMA ids are COMPUTED, not skipped
  m.def("numa_nodes_required", [](int qos, const NodeResourceTopology& nrt, const ResourceList& resources) {
    bool is_min = false;
    auto ids = NumaNodesRequired((QOS)qos, CreateNUMANodeList(nrt), resources, &is_min);
    return py::make_tuple(ids, is_min);
  });
  m.def("only_non_numa_resources", [](const NodeResourceTopology& nrt, const ResourceList& resources) {
    return OnlyNonNUMAResources(CreateNUMANodeList(nrt), resources);
  });
  m.def("scalar_filter", [](const Pod& pod, const NodeInfo& ni, const NodeResourceTopology& nrt) { return ScalarFilter(pod, ni, nrt); });
  m.def("scalar_score", [](const Pod& pod, const NodeResourceTopology& nrt, int strategy, const std::map<std::string, int64_t>& w) {
    return ScalarScore(pod, nrt, strategy, w);
  });
  py::class_<TopologyMatch>(m, "TopologyMatch")
      .def_static("new", &TopologyMatch::New)
      .def("name", &TopologyMatch::Name)
      .def("filter", &TopologyMatch::Filter)
      .def("score", &TopologyMatch::Score);
  py::class_<NetworkOverheadArgs>(m, "NetworkOverheadArgs")
      .def(py::init<>())
      .def_readwrite("namespaces", &NetworkOverheadArgs::namespaces)
      .def_readwrite("weights_name", &NetworkOverheadArgs::weights_name)
      .def_readwrite("network_topology_name", &NetworkOverheadArgs::network_topology_name);
  py::class_<NetworkOverhead>(m, "NetworkOverhead")
      .def_static("new", &NetworkOverhead::New)
      .def("name", &NetworkOverhead::Name)
      .def("pre_filter", &NetworkOverhead::PreFilter)
      .def("filter", &NetworkOverhead::Filter)
      .def("score", &NetworkOverhead::Score)
      .def("normalize_score", [](NetworkOverhead& a, CycleState& s, const Pod& p, std::vector<NodeScore> l) {
        Status st = a.NormalizeScore(s, p, l);
        return std::make_pair(st, l);
      });
}

// C++ mirror of the five plugins: flatten -> C-ABI -> per-node lookup.  See plugins.hpp.
#include "plugins.hpp"
#include "nrt_scalar.hpp"

#include <algorithm>
#include <cstring>
#include <set>
#include <stdexcept>

namespace b200host {

// ---------------------------------------------------------------- Engine
Engine::Engine(int device) {
  int rc = b200s_init(device, &ctx_);
  if (rc != B200S_OK) throw std::runtime_error(std::string("b200s_init failed: ") + b200s_last_error(nullptr));
}
Engine::~Engine() {
  if (ctx_) b200s_shutdown(ctx_);
}
std::string Engine::LastError() const { return b200s_last_error(ctx_); }
void Engine::Check(int rc, const char* what) const {
  if (rc != B200S_OK) throw std::runtime_error(std::string(what) + ": " + LastError());
}

namespace {

int32_t NPad(int32_t n) { return ((std::max(n, 1) + B200S_NODE_ALIGN - 1) / B200S_NODE_ALIGN) * B200S_NODE_ALIGN; }

Status ErrorStatus(const std::string& msg) { return Status{Code::Error, msg}; }

// node name -> column index of the handle's snapshot; nodes without a Node object are skipped
std::map<std::string, int32_t> IndexOf(const std::vector<NodeInfo>& nodes) {
  std::map<std::string, int32_t> idx;
  for (size_t i = 0; i < nodes.size(); ++i)
    if (nodes[i].GetNode()) idx[nodes[i].GetNode()->name] = (int32_t)i;
  return idx;
}

std::vector<uint64_t> FeasibleWords(const std::map<std::string, int32_t>& index, int32_t npad,
                                    const std::vector<std::string>& names) {
  std::vector<uint64_t> w(npad / 64, 0);
  for (auto& nm : names) {
    auto it = index.find(nm);
    if (it != index.end()) w[it->second >> 6] |= 1ull << (it->second & 63);
  }
  return w;
}

std::vector<std::string> NamesOf(const std::vector<NodeInfo>& nodes) {
  std::vector<std::string> out;
  for (auto& n : nodes)
    if (n.GetNode()) out.push_back(n.GetNode()->name);
  return out;
}

int64_t Get(const ResourceList& r, const std::string& k) {
  auto it = r.find(k);
  return it == r.end() ? 0 : it->second;
}

template <class T>
std::shared_ptr<T> Read(CycleState& s, const std::string& key) {
  auto it = s.data.find(key);
  if (it == s.data.end()) return nullptr;
  return std::static_pointer_cast<T>(it->second);
}

std::pair<int64_t, Status> Lookup(const std::shared_ptr<CycleResult>& c, const NodeInfo& ni) {
  if (!c->engine_error.empty()) return {0, ErrorStatus(c->engine_error)};
  auto it = c->index.find(ni.GetNode()->name);
  if (it == c->index.end()) return {0, ErrorStatus("node not in the cycle's snapshot: " + ni.GetNode()->name)};
  return {(int64_t)c->scores[it->second], Status{}};
}

}  // namespace

// =============================================================== NodeResourcesAllocatable
std::unique_ptr<Allocatable> Allocatable::New(const NodeResourcesAllocatableArgs* args, std::shared_ptr<Handle> h) {
  std::unique_ptr<Allocatable> a(new Allocatable());
  a->h_ = std::move(h);
  a->res_ = {{ResourceMemory, 1}, {ResourceCPU, 1 << 20}};  // defaultResourcesToWeightMap, resource_allocation.go:36
  std::string mode = "Least";
  if (args) {
    if (!args->mode.empty()) mode = args->mode;
    // ValidateNodeResourcesAllocatableArgs: validation_pluginargs.go:65-95
    for (size_t i = 0; i < args->resources.size(); ++i)
      if (args->resources[i].weight <= 0)
        throw std::invalid_argument("resources[" + std::to_string(i) + "].weight: Invalid value: " +
                                    std::to_string(args->resources[i].weight) + ": resource weight of " +
                                    args->resources[i].name + " should be a positive value, got :" +
                                    std::to_string(args->resources[i].weight));
    if (mode != "Least" && mode != "Most") throw std::invalid_argument("mode: Invalid value: \"" + mode + "\": invalid support ModeType");
    if (!args->resources.empty()) a->res_ = args->resources;
  }
  a->mode_ = mode == "Most" ? B200S_ALLOC_MOST : B200S_ALLOC_LEAST;
  a->eng_.reset(new Engine(a->h_->device));
  std::vector<int64_t> w;
  for (auto& r : a->res_) w.push_back(r.weight);
  a->eng_->Check(b200s_config_allocatable(a->eng_->ctx(), a->mode_, (int32_t)w.size(), w.data()), "config_allocatable");
  return a;
}

// calculateResourceAllocatableRequest's allocatable side: resource_allocation.go:79-100
static int64_t AllocatableColumn(const Node& n, const std::string& res) {
  if (res == ResourceCPU) return Get(n.allocatable, ResourceCPU);                      // GetMilliCPU
  if (res == ResourceMemory) return QuantityValue(Get(n.allocatable, ResourceMemory));  // bytes
  if (res == ResourceEphemeralStorage) return QuantityValue(Get(n.allocatable, ResourceEphemeralStorage));
  if (IsScalarResourceName(res)) return QuantityValue(Get(n.allocatable, res));
  return 0;
}

bool Handle::NodesChangedSince(uint64_t g, std::vector<int32_t>* out) const {
  if (g < log_base_ || g > generation) return false;
  std::vector<int32_t> idx;
  uint64_t logged = 0;
  for (auto it = node_log_.rbegin(); it != node_log_.rend() && it->first > g; ++it, ++logged) idx.push_back(it->second);
  if (logged != generation - g) return false;  // generation moved without a log entry (set directly)
  std::sort(idx.begin(), idx.end());
  idx.erase(std::unique(idx.begin(), idx.end()), idx.end());
  for (int32_t i : idx)
    if (i < 0 || i >= (int32_t)node_infos.size()) return false;
  *out = std::move(idx);
  return true;
}

// A patch pays off while few rows changed; past a quarter of the nodes one bulk upload is cheaper.
static bool WorthPatching(size_t changed, int32_t n) { return changed * 4 <= (size_t)std::max(n, 1); }

bool Allocatable::PatchSnapshot() {
  std::vector<int32_t> idx;
  if (snap_gen_ == 0 || n_ != (int32_t)h_->node_infos.size() || !h_->NodesChangedSince(snap_gen_, &idx) ||
      !WorthPatching(idx.size(), n_))
    return false;
  const auto& nodes = h_->node_infos;
  for (int32_t i : idx) {  // the name -> column map must still hold
    const Node* nd = nodes[i].GetNode();
    auto it = nd ? index_.find(nd->name) : index_.end();
    if (nd && (it == index_.end() || it->second != i)) return false;
  }
  std::vector<std::vector<int64_t>> cols(res_.size(), std::vector<int64_t>(std::max<size_t>(idx.size(), 1), 0));
  for (size_t j = 0; j < idx.size(); ++j)
    if (const Node* nd = nodes[idx[j]].GetNode())
      for (size_t r = 0; r < res_.size(); ++r) cols[r][j] = AllocatableColumn(*nd, res_[r].name);
  std::vector<const int64_t*> ptrs;
  for (auto& c : cols) ptrs.push_back(c.data());
  // A failed patch (allocation, CUDA error, size check) leaves the engine's snapshot open and half rewritten:
  // report "not patched" so that EnsureSnapshot does the full b200s_snapshot_begin upload, which resets it.
  try {
    eng_->Check(b200s_snapshot_patch_begin(eng_->ctx(), h_->generation), "snapshot_patch_begin");
    eng_->Check(b200s_snapshot_patch_allocatable(eng_->ctx(), (int32_t)idx.size(), idx.data(), (int32_t)ptrs.size(), ptrs.data()),
                "snapshot_patch_allocatable");
    eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  } catch (const std::exception&) {
    snap_gen_ = 0;
    return false;
  }
  snap_gen_ = h_->generation;
  patched_rows_ += (int64_t)idx.size();
  return true;
}

void Allocatable::EnsureSnapshot() {
  if (snap_gen_ == h_->generation && n_ == (int32_t)h_->node_infos.size()) return;
  if (PatchSnapshot()) return;
  const auto& nodes = h_->node_infos;
  n_ = (int32_t)nodes.size();
  npad_ = NPad(n_);
  index_ = IndexOf(nodes);
  std::vector<std::vector<int64_t>> cols(res_.size(), std::vector<int64_t>(std::max(n_, 1), 0));
  for (int32_t i = 0; i < n_; ++i)
    if (nodes[i].GetNode())
      for (size_t r = 0; r < res_.size(); ++r) cols[r][i] = AllocatableColumn(*nodes[i].GetNode(), res_[r].name);
  std::vector<const int64_t*> ptrs;
  for (auto& c : cols) ptrs.push_back(c.data());
  eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n_, 0, n_), "snapshot_begin");
  eng_->Check(b200s_snapshot_allocatable(eng_->ctx(), (int32_t)ptrs.size(), ptrs.data()), "snapshot_allocatable");
  eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  snap_gen_ = h_->generation;
}

std::shared_ptr<CycleResult> Allocatable::Run(const Pod&, const std::vector<NodeInfo>* feasible) {
  auto c = std::make_shared<CycleResult>();
  try {
    EnsureSnapshot();
    c->index = index_;
    c->scores.assign(npad_, 0);
    std::vector<uint64_t> words;
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    if (feasible) {
      words = FeasibleWords(index_, npad_, NamesOf(*feasible));
      b.feasible = words.data();
    }
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_ALLOCATABLE, &b, B200S_OUT_U8, c->scores.data(), nullptr, nullptr),
                "score_batch(NodeResourcesAllocatable)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status Allocatable::PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes) {
  auto c = Run(pod, &nodes);
  state.data[std::string("PreScore") + AllocatableName] = c;
  return c->engine_error.empty() ? Status{} : ErrorStatus(c->engine_error);
}

std::pair<int64_t, Status> Allocatable::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};  // resource_allocation.go:53-56
  auto c = Read<CycleResult>(state, std::string("PreScore") + AllocatableName);
  if (!c) {  // PreScore not called (unit-test style): every node of the snapshot is in the list
    c = Run(pod, nullptr);
    state.data[std::string("PreScore") + AllocatableName] = c;
  }
  return Lookup(c, nodeInfo);
}

Status Allocatable::NormalizeScore(CycleState&, const Pod&, std::vector<NodeScore>&) {
  return {};  // Score already returned the value normalised over the cycle's feasible list
}

// =============================================================== TargetLoadPacking
std::unique_ptr<TargetLoadPacking> TargetLoadPacking::New(const TargetLoadPackingArgs& args, std::shared_ptr<Handle> h) {
  std::unique_ptr<TargetLoadPacking> p(new TargetLoadPacking());
  p->h_ = std::move(h);
  p->args_ = args;
  try {  // strconv.ParseFloat, targetloadpacking.go:80-83
    size_t pos = 0;
    p->multiplier_ = std::stod(args.default_requests_multiplier, &pos);
    if (pos != args.default_requests_multiplier.size()) throw std::invalid_argument("trailing characters");
  } catch (const std::exception& e) {
    throw std::invalid_argument(std::string("unable to parse DefaultRequestsMultiplier: ") + e.what());
  }
  p->eng_.reset(new Engine(p->h_->device));
  p->eng_->Check(b200s_config_tlp(p->eng_->ctx(), args.target_utilization), "config_tlp");
  return p;
}

TargetLoadPacking::Row TargetLoadPacking::FlattenRow(const NodeInfo& ni) const {
  Row row;
  const Node* nd = ni.GetNode();
  if (!nd) return row;
  row.cap = Get(nd->capacity, ResourceCPU);  // Status.Capacity, targetloadpacking.go:146
  const WatcherMetrics* wm = h_->metrics.get();
  // GetNodeMetrics: collector.go:110-123
  if (!wm || !wm->has_map) return row;
  auto it = wm->node_metrics.find(nd->name);
  if (it == wm->node_metrics.end()) return row;
  row.flags |= B200S_TLP_HAS_METRICS;
  for (auto& mt : it->second.metrics)  // LAST matching entry wins, :131-140
    if (mt.type == "CPU" && (mt.op == "AVG" || mt.op == "Latest")) {
      row.util = mt.value;
      row.flags |= B200S_TLP_CPU_FOUND;
    }
  // missing utilisation of recently bound pods, :151-167
  auto sc = h_->scheduled_pods_cache.find(nd->name);
  if (sc != h_->scheduled_pods_cache.end())
    for (auto& info : sc->second) {
      const int64_t ts = info.timestamp_unix, end = wm->window_end;
      if (ts > end || (ts <= end && (end - ts) < 60)) {  // metricsAgentReportingIntervalSeconds = 60, :46
        for (auto& cont : info.pod->containers)
          row.missing += PredictUtilisation(cont, args_.default_requests_cpu_milli, multiplier_);
        row.missing += Get(info.pod->overhead, ResourceCPU);
      }
    }
  return row;
}

// A bind adds one entry to ScheduledPodsCache[node] (handler.go:131-167): one row changes, the rest stays.
bool TargetLoadPacking::PatchSnapshot() {
  std::vector<int32_t> idx;
  if (snap_gen_ == 0 || n_ != (int32_t)h_->node_infos.size() || !h_->NodesChangedSince(snap_gen_, &idx) ||
      !WorthPatching(idx.size(), n_))
    return false;
  const size_t m = std::max<size_t>(idx.size(), 1);
  std::vector<double> util(m, 0);
  std::vector<int64_t> cap(m, 0), missing(m, 0);
  std::vector<uint8_t> flags(m, 0);
  for (size_t j = 0; j < idx.size(); ++j) {
    const NodeInfo& ni = h_->node_infos[idx[j]];
    auto it = ni.GetNode() ? index_.find(ni.GetNode()->name) : index_.end();
    if (ni.GetNode() && (it == index_.end() || it->second != idx[j])) return false;
    const Row row = FlattenRow(ni);
    util[j] = row.util, cap[j] = row.cap, missing[j] = row.missing, flags[j] = row.flags;
  }
  // A failed patch (allocation, CUDA error, size check) leaves the engine's snapshot open and half rewritten:
  // report "not patched" so that EnsureSnapshot does the full b200s_snapshot_begin upload, which resets it.
  try {
    eng_->Check(b200s_snapshot_patch_begin(eng_->ctx(), h_->generation), "snapshot_patch_begin");
    eng_->Check(b200s_snapshot_patch_tlp(eng_->ctx(), (int32_t)idx.size(), idx.data(), util.data(), cap.data(), missing.data(),
                                         flags.data()),
                "snapshot_patch_tlp");
    eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  } catch (const std::exception&) {
    snap_gen_ = 0;
    return false;
  }
  snap_gen_ = h_->generation;
  patched_rows_ += (int64_t)idx.size();
  return true;
}

void TargetLoadPacking::EnsureSnapshot() {
  if (snap_gen_ == h_->generation && n_ == (int32_t)h_->node_infos.size()) return;
  if (PatchSnapshot()) return;
  const auto& nodes = h_->node_infos;
  n_ = (int32_t)nodes.size();
  npad_ = NPad(n_);
  index_ = IndexOf(nodes);
  const int m = std::max(n_, 1);
  std::vector<double> util(m, 0);
  std::vector<int64_t> cap(m, 0), missing(m, 0);
  std::vector<uint8_t> flags(m, 0);
  for (int32_t i = 0; i < n_; ++i) {
    const Row row = FlattenRow(nodes[i]);
    util[i] = row.util, cap[i] = row.cap, missing[i] = row.missing, flags[i] = row.flags;
  }
  eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n_, 0, n_), "snapshot_begin");
  eng_->Check(b200s_snapshot_tlp(eng_->ctx(), util.data(), cap.data(), missing.data(), flags.data()), "snapshot_tlp");
  eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  snap_gen_ = h_->generation;
}

std::shared_ptr<CycleResult> TargetLoadPacking::Run(const Pod& pod) {
  auto c = std::make_shared<CycleResult>();
  try {
    EnsureSnapshot();
    c->index = index_;
    c->scores.assign(npad_, 0);
    int64_t pod_cpu = PodPredictedCPU(pod, args_.default_requests_cpu_milli, multiplier_);
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.tlp_pod_cpu_milli = &pod_cpu;
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_TLP, &b, B200S_OUT_U8, c->scores.data(), nullptr, nullptr),
                "score_batch(TargetLoadPacking)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status TargetLoadPacking::PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>&) {
  auto c = Run(pod);
  state.data[std::string("PreScore") + Name_] = c;
  return c->engine_error.empty() ? Status{} : ErrorStatus(c->engine_error);
}

std::pair<int64_t, Status> TargetLoadPacking::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};
  auto c = Read<CycleResult>(state, std::string("PreScore") + Name_);
  if (!c) {
    c = Run(pod);
    state.data[std::string("PreScore") + Name_] = c;
  }
  return Lookup(c, nodeInfo);
}

// =============================================================== LoadVariationRiskBalancing
std::unique_ptr<LoadVariationRiskBalancing> LoadVariationRiskBalancing::New(const LoadVariationRiskBalancingArgs& args,
                                                                           std::shared_ptr<Handle> h) {
  std::unique_ptr<LoadVariationRiskBalancing> p(new LoadVariationRiskBalancing());
  p->h_ = std::move(h);
  p->args_ = args;
  p->eng_.reset(new Engine(p->h_->device));
  p->eng_->Check(b200s_config_lvrb(p->eng_->ctx(), args.safe_variance_margin, args.safe_variance_sensitivity), "config_lvrb");
  return p;
}

LoadVariationRiskBalancing::Row LoadVariationRiskBalancing::FlattenRow(const NodeInfo& ni) const {
  Row row;
  const Node* nd = ni.GetNode();
  if (!nd) return row;
  row.acpu = Get(nd->allocatable, ResourceCPU);                    // resourcestats.go:55-59: Allocatable
  row.amem = QuantityValue(Get(nd->allocatable, ResourceMemory));  // :62 am.Value()
  const WatcherMetrics* wm = h_->metrics.get();
  if (!wm || !wm->has_map) return row;
  auto it = wm->node_metrics.find(nd->name);
  if (it == wm->node_metrics.end()) return row;
  row.flags |= B200S_LVRB_HAS_METRICS;
  bool ok;
  GetResourceData(it->second.metrics, "CPU", &row.ca, &row.cs, &ok);
  if (ok) row.flags |= B200S_LVRB_CPU_OK;
  GetResourceData(it->second.metrics, "Memory", &row.ma, &row.ms, &ok);
  if (ok) row.flags |= B200S_LVRB_MEM_OK;
  return row;
}

bool LoadVariationRiskBalancing::PatchSnapshot() {
  std::vector<int32_t> idx;
  if (snap_gen_ == 0 || n_ != (int32_t)h_->node_infos.size() || !h_->NodesChangedSince(snap_gen_, &idx) ||
      !WorthPatching(idx.size(), n_))
    return false;
  const size_t m = std::max<size_t>(idx.size(), 1);
  std::vector<double> ca(m, 0), cs(m, 0), ma(m, 0), ms(m, 0);
  std::vector<int64_t> acpu(m, 0), amem(m, 0);
  std::vector<uint8_t> flags(m, 0);
  for (size_t j = 0; j < idx.size(); ++j) {
    const NodeInfo& ni = h_->node_infos[idx[j]];
    auto it = ni.GetNode() ? index_.find(ni.GetNode()->name) : index_.end();
    if (ni.GetNode() && (it == index_.end() || it->second != idx[j])) return false;
    const Row row = FlattenRow(ni);
    ca[j] = row.ca, cs[j] = row.cs, ma[j] = row.ma, ms[j] = row.ms, acpu[j] = row.acpu, amem[j] = row.amem, flags[j] = row.flags;
  }
  // A failed patch (allocation, CUDA error, size check) leaves the engine's snapshot open and half rewritten:
  // report "not patched" so that EnsureSnapshot does the full b200s_snapshot_begin upload, which resets it.
  try {
    eng_->Check(b200s_snapshot_patch_begin(eng_->ctx(), h_->generation), "snapshot_patch_begin");
    eng_->Check(b200s_snapshot_patch_lvrb(eng_->ctx(), (int32_t)idx.size(), idx.data(), ca.data(), cs.data(), ma.data(), ms.data(),
                                          acpu.data(), amem.data(), flags.data()),
                "snapshot_patch_lvrb");
    eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  } catch (const std::exception&) {
    snap_gen_ = 0;
    return false;
  }
  snap_gen_ = h_->generation;
  patched_rows_ += (int64_t)idx.size();
  return true;
}

void LoadVariationRiskBalancing::EnsureSnapshot() {
  if (snap_gen_ == h_->generation && n_ == (int32_t)h_->node_infos.size()) return;
  if (PatchSnapshot()) return;
  const auto& nodes = h_->node_infos;
  n_ = (int32_t)nodes.size();
  npad_ = NPad(n_);
  index_ = IndexOf(nodes);
  const int m = std::max(n_, 1);
  std::vector<double> ca(m, 0), cs(m, 0), ma(m, 0), ms(m, 0);
  std::vector<int64_t> acpu(m, 0), amem(m, 0);
  std::vector<uint8_t> flags(m, 0);
  for (int32_t i = 0; i < n_; ++i) {
    const Row row = FlattenRow(nodes[i]);
    ca[i] = row.ca, cs[i] = row.cs, ma[i] = row.ma, ms[i] = row.ms, acpu[i] = row.acpu, amem[i] = row.amem, flags[i] = row.flags;
  }
  eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n_, 0, n_), "snapshot_begin");
  eng_->Check(b200s_snapshot_lvrb(eng_->ctx(), ca.data(), cs.data(), ma.data(), ms.data(), acpu.data(), amem.data(),
                                  flags.data()),
              "snapshot_lvrb");
  eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  snap_gen_ = h_->generation;
}

std::shared_ptr<CycleResult> LoadVariationRiskBalancing::Run(const Pod& pod) {
  auto c = std::make_shared<CycleResult>();
  try {
    EnsureSnapshot();
    c->index = index_;
    c->scores.assign(npad_, 0);
    int64_t cpu, mem;
    GetResourceRequested(pod, &cpu, &mem);
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.lvrb_req_cpu_milli = &cpu;
    b.lvrb_req_mem_bytes = &mem;
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_LVRB, &b, B200S_OUT_U8, c->scores.data(), nullptr, nullptr),
                "score_batch(LoadVariationRiskBalancing)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status LoadVariationRiskBalancing::PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>&) {
  auto c = Run(pod);
  state.data[std::string("PreScore") + Name_] = c;
  return c->engine_error.empty() ? Status{} : ErrorStatus(c->engine_error);
}

std::pair<int64_t, Status> LoadVariationRiskBalancing::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};
  auto c = Read<CycleResult>(state, std::string("PreScore") + Name_);
  if (!c) {
    c = Run(pod);
    state.data[std::string("PreScore") + Name_] = c;
  }
  return Lookup(c, nodeInfo);
}

// =============================================================== Peaks
std::unique_ptr<Peaks> Peaks::New(const PeaksArgs& args, std::shared_ptr<Handle> h) {
  std::unique_ptr<Peaks> p(new Peaks());
  p->h_ = std::move(h);
  p->args_ = args;
  p->eng_ = std::make_unique<Engine>(p->h_->device);
  return p;
}

void Peaks::EnsureSnapshot() {
  if (snap_gen_ == h_->generation && n_ == (int32_t)h_->node_infos.size()) return;
  const auto& nodes = h_->node_infos;
  n_ = (int32_t)nodes.size();
  npad_ = NPad(n_);
  index_ = IndexOf(nodes);
  const int m = std::max(n_, 1);
  std::vector<double> util(m, 0), k1(m, 0), k2(m, 0);
  std::vector<int64_t> cap(m, 0);
  std::vector<uint8_t> flags(m, 0);
  const WatcherMetrics* wm = h_->metrics.get();
  for (int32_t i = 0; i < n_; ++i) {
    const Node* nd = nodes[i].GetNode();
    if (!nd) continue;
    cap[i] = Get(nd->capacity, ResourceCPU);  // Status.Capacity, peaks.go:131
    auto pm = args_.node_power_model.find(nd->name);  // getPowerModel :193-199: missing -> {0, 0, 0}
    if (pm != args_.node_power_model.end()) k1[i] = pm->second.k1, k2[i] = pm->second.k2;
    if (!wm || !wm->has_map) continue;
    auto it = wm->node_metrics.find(nd->name);
    if (it == wm->node_metrics.end()) continue;
    flags[i] |= B200S_TLP_HAS_METRICS;
    for (auto& mt : it->second.metrics)  // the FIRST matching entry wins, :117-126
      if (mt.type == "CPU" && (mt.op == "AVG" || mt.op == "Latest")) {
        util[i] = mt.value;
        flags[i] |= B200S_TLP_CPU_FOUND;
        break;
      }
  }
  eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n_, 0, n_), "snapshot_begin");
  eng_->Check(b200s_snapshot_peaks(eng_->ctx(), util.data(), cap.data(), flags.data(), k1.data(), k2.data()), "snapshot_peaks");
  eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  snap_gen_ = h_->generation;
}

std::shared_ptr<CycleResult> Peaks::Run(const Pod& pod, const std::vector<NodeInfo>* feasible) {
  auto c = std::make_shared<CycleResult>();
  try {
    EnsureSnapshot();
    c->index = index_;
    c->scores.assign(npad_, 0);
    const int64_t cpu = GetResourceRequestQuantityCPU(pod);  // :113-114
    std::vector<uint64_t> words;
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.peaks_pod_cpu_milli = &cpu;
    if (feasible) {
      words = FeasibleWords(index_, npad_, NamesOf(*feasible));
      b.feasible = words.data();
    }
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_PEAKS, &b, B200S_OUT_U8, c->scores.data(), nullptr, nullptr),
                "score_batch(Peaks)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status Peaks::PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>& nodes) {
  auto c = Run(pod, &nodes);
  state.data[std::string("PreScore") + Name_] = c;
  return c->engine_error.empty() ? Status{} : ErrorStatus(c->engine_error);
}

std::pair<int64_t, Status> Peaks::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};
  auto c = Read<CycleResult>(state, std::string("PreScore") + Name_);
  if (!c) {  // PreScore not called: every node of the snapshot is in the list
    c = Run(pod, nullptr);
    state.data[std::string("PreScore") + Name_] = c;
  }
  return Lookup(c, nodeInfo);
}

// =============================================================== LowRiskOverCommitment
std::unique_ptr<LowRiskOverCommitment> LowRiskOverCommitment::New(const LowRiskOverCommitmentArgs& args,
                                                                  std::shared_ptr<Handle> h) {
  std::unique_ptr<LowRiskOverCommitment> p(new LowRiskOverCommitment());
  p->h_ = std::move(h);
  p->args_ = args;
  // SetDefaults_LowRiskOverCommitmentArgs, apis/config/v1/defaults.go:170-183
  if (p->args_.smoothing_window_size <= 0) p->args_.smoothing_window_size = 5;
  auto fix = [](double& w) {
    if (!(w >= 0 && w <= 1)) w = 0.5;
  };
  fix(p->args_.risk_limit_weight_cpu);
  fix(p->args_.risk_limit_weight_memory);
  p->eng_ = std::make_unique<Engine>(p->h_->device);
  p->eng_->Check(b200s_config_low_risk(p->eng_->ctx(), p->args_.smoothing_window_size, p->args_.risk_limit_weight_cpu,
                                       p->args_.risk_limit_weight_memory),
                 "config_low_risk");
  return p;
}

void LowRiskOverCommitment::EnsureSnapshot() {
  if (snap_gen_ == h_->generation && n_ == (int32_t)h_->node_infos.size()) return;
  const auto& nodes = h_->node_infos;
  n_ = (int32_t)nodes.size();
  npad_ = NPad(n_);
  index_ = IndexOf(nodes);
  const int m = std::max(n_, 1);
  std::vector<double> ca(m, 0), cs(m, 0), ma(m, 0), ms(m, 0);
  std::vector<int64_t> acpu(m, 0), amem(m, 0), rc(m, 0), rm(m, 0), lc(m, 0), lm(m, 0);
  std::vector<uint8_t> flags(m, 0);
  const WatcherMetrics* wm = h_->metrics.get();
  for (int32_t i = 0; i < n_; ++i) {
    const Node* nd = nodes[i].GetNode();
    if (!nd) continue;
    acpu[i] = Get(nd->allocatable, ResourceCPU);                    // resourcestats.go:170-175
    amem[i] = QuantityValue(Get(nd->allocatable, ResourceMemory));
    {  // GetNodeRequestsAndLimits :181-206 without the pending pod
      int64_t sums[4];
      NodeRequestsAndLimitsOfRunningPods(nodes[i], sums);
      rc[i] = sums[0], rm[i] = sums[1], lc[i] = sums[2], lm[i] = sums[3];
    }
    if (!wm || !wm->has_map) continue;
    auto it = wm->node_metrics.find(nd->name);
    if (it == wm->node_metrics.end()) continue;
    flags[i] |= B200S_LVRB_HAS_METRICS;
    bool ok;
    GetResourceData(it->second.metrics, "CPU", &ca[i], &cs[i], &ok);
    if (ok) flags[i] |= B200S_LVRB_CPU_OK;
    GetResourceData(it->second.metrics, "Memory", &ma[i], &ms[i], &ok);
    if (ok) flags[i] |= B200S_LVRB_MEM_OK;
  }
  eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n_, 0, n_), "snapshot_begin");
  eng_->Check(b200s_snapshot_low_risk(eng_->ctx(), ca.data(), cs.data(), ma.data(), ms.data(), acpu.data(), amem.data(),
                                      flags.data(), rc.data(), rm.data(), lc.data(), lm.data()),
              "snapshot_low_risk");
  eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
  snap_gen_ = h_->generation;
}

std::shared_ptr<CycleResult> LowRiskOverCommitment::Run(const Pod& pod) {
  auto c = std::make_shared<CycleResult>();
  try {
    EnsureSnapshot();
    c->index = index_;
    c->scores.assign(npad_, 0);
    int64_t v[4];  // CreatePodResourcesStateData :257-267
    GetResourceRequested(pod, &v[0], &v[1]);
    GetResourceLimits(pod, &v[2], &v[3]);
    v[2] = std::max(v[2], v[0]), v[3] = std::max(v[3], v[1]);
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.low_risk_pod = v;
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_LOW_RISK, &b, B200S_OUT_U8, c->scores.data(), nullptr, nullptr),
                "score_batch(LowRiskOverCommitment)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status LowRiskOverCommitment::PreScore(CycleState& state, const Pod& pod, const std::vector<NodeInfo>&) {
  auto c = Run(pod);
  state.data[std::string("PreScore") + Name_] = c;
  return c->engine_error.empty() ? Status{} : ErrorStatus(c->engine_error);
}

std::pair<int64_t, Status> LowRiskOverCommitment::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};
  auto c = Read<CycleResult>(state, std::string("PreScore") + Name_);
  if (!c) {  // recalculating, :113-118
    c = Run(pod);
    state.data[std::string("PreScore") + Name_] = c;
  }
  return Lookup(c, nodeInfo);
}

// =============================================================== NodeResourceTopologyMatch
std::unique_ptr<TopologyMatch> TopologyMatch::New(const NodeResourceTopologyMatchArgs& args, std::shared_ptr<Handle> h) {
  std::unique_ptr<TopologyMatch> p(new TopologyMatch());
  p->h_ = std::move(h);
  p->args_ = args;
  static const std::map<std::string, int> strat = {{"MostAllocated", B200S_NRT_MOST_ALLOCATED},
                                                   {"BalancedAllocation", B200S_NRT_BALANCED_ALLOCATION},
                                                   {"LeastAllocated", B200S_NRT_LEAST_ALLOCATED},
                                                   {"LeastNUMANodes", B200S_NRT_LEAST_NUMA_NODES}};
  auto it = strat.find(args.scoring_strategy);
  if (it == strat.end()) throw std::invalid_argument("illegal scoring strategy found");  // score.go:138
  p->strategy_ = it->second;
  p->eng_.reset(new Engine(p->h_->device));
  return p;
}

namespace {

constexpr int Z_MAX = B200S_NRT_MAX_ZONES, R_MAX = B200S_NRT_MAX_RES, C_MAX = B200S_NRT_MAX_CONT;

// createNUMANodeList (pluginhelpers.go:105-134): zones of type Node named node-<id>, id <= 64.
// The dense encoding needs ids 0..k-1 in list order (see include/b200sched.h); else UNSUPPORTED.
// numanode.NameToID: "node-<decimal id>" -> id, -1 if the name has another shape
long NumaNameToID(const std::string& name) {
  if (name.rfind("node-", 0) != 0 || name.size() == 5 || name.size() > 5 + 9) return -1;
  long id = 0;
  for (size_t i = 5; i < name.size(); ++i) {
    if (name[i] < '0' || name[i] > '9') return -1;
    id = id * 10 + (name[i] - '0');
  }
  return id;
}

bool NumaZones(const NodeResourceTopology& nrt, std::vector<const Zone*>* out) {
  std::vector<int> ids;
  for (auto& z : nrt.zones) {
    if (z.type != "Node") continue;
    long id = NumaNameToID(z.name);
    if (id < 0 || id > 64) continue;
    out->push_back(&z);
    ids.push_back((int)id);
  }
  if ((int)ids.size() > Z_MAX) return false;
  for (size_t i = 0; i < ids.size(); ++i)
    if (ids[i] != (int)i) return false;
  return true;
}

}  // namespace

std::shared_ptr<CycleResult> TopologyMatch::Run(CycleState& state, const Pod& pod) {
  const std::string key = std::string("PreFilter") + Name_;
  if (auto c = Read<CycleResult>(state, key)) return c;
  auto c = std::make_shared<CycleResult>();
  state.data[key] = c;
  try {
    const auto& nodes = h_->node_infos;
    const int32_t n = (int32_t)nodes.size(), npad = NPad(n), m = std::max(n, 1);
    c->index = IndexOf(nodes);
    // ---- resource-slot dictionary of this pod: every requested resource name (cpu, memory first)
    std::vector<std::string> names = {ResourceCPU, ResourceMemory};
    auto add = [&](const std::string& r) {
      if (std::find(names.begin(), names.end(), r) == names.end()) names.push_back(r);
    };
    for (auto& ct : pod.init_containers)
      for (auto& kv : ct.requests) add(kv.first);
    for (auto& ct : pod.containers)
      for (auto& kv : ct.requests) add(kv.first);
    if (pod.has_overhead)
      for (auto& kv : pod.overhead) add(kv.first);
    const int R = std::min((int)names.size(), R_MAX);
    std::map<std::string, int> slot;
    for (size_t r = 0; r < names.size(); ++r) slot[names[r]] = (int)r;
    // ---- nodes
    std::vector<std::vector<const Zone*>> zl(m);
    std::vector<bool> ok(m, true);
    int Z = 1;
    for (int32_t i = 0; i < n; ++i) {
      const Node* nd = nodes[i].GetNode();
      if (!nd) continue;
      auto it = h_->nrts.find(nd->name);
      if (it == h_->nrts.end() || !it->second) continue;
      ok[i] = NumaZones(*it->second, &zl[i]);
      if (ok[i]) Z = std::max(Z, (int)zl[i].size());
    }
    std::vector<uint8_t> res_flags(R), node_flags(m, 0), nz(m, 0), node_res_mask(m, 0), zmask((size_t)Z * m, 0);
    std::vector<uint16_t> max_numa(m, 8);
    std::vector<int64_t> avail((size_t)Z * R * m, 0);
    std::vector<int32_t> cost((size_t)Z * Z * m, -1);
    for (int r = 0; r < R; ++r)
      res_flags[r] = (IsNUMAAffineResource(names[r]) ? B200S_NRT_RES_AFFINE : 0) |
                     (IsHostLevelResource(names[r]) ? B200S_NRT_RES_HOST_LEVEL : 0);
    for (int32_t i = 0; i < n; ++i) {
      const Node* nd = nodes[i].GetNode();
      if (!nd) continue;
      uint8_t fl = 0;
      auto nf = h_->nrt_not_fresh.find(nd->name);
      if (nf == h_->nrt_not_fresh.end() || !nf->second) fl |= B200S_NRT_NODE_FRESH;
      // util.ResourceList(GetAllocatable()): pkg/util/resource.go:28-44
      for (int r = 0; r < R; ++r) {
        const std::string& nm = names[r];
        bool present = nm == ResourceCPU || nm == ResourceMemory || nm == ResourcePods || nm == ResourceEphemeralStorage ||
                       (nd->allocatable.count(nm) && IsScalarResourceName(nm));
        if (present) node_res_mask[i] |= (uint8_t)(1u << r);
      }
      auto it = h_->nrts.find(nd->name);
      if (it != h_->nrts.end() && it->second) {
        fl |= B200S_NRT_NODE_HAS_NRT;
        TopologyManager tm = TopologyManagerFromNodeResourceTopology(*it->second);
        if (tm.policy == "single-numa-node") fl |= B200S_NRT_NODE_SINGLE_NUMA;
        if (tm.scope == "pod") fl |= B200S_NRT_NODE_SCOPE_POD;
        max_numa[i] = (uint16_t)tm.max_numa_nodes;
        if (!ok[i]) {
          fl |= B200S_NRT_NODE_UNSUPPORTED;
          c->unsupported_nodes.insert(i);
        } else {
          nz[i] = (uint8_t)zl[i].size();
          for (size_t z = 0; z < zl[i].size(); ++z) {
            for (int r = 0; r < R; ++r) {
              auto zr = zl[i][z]->resources.find(names[r]);
              if (zr == zl[i][z]->resources.end()) continue;
              zmask[z * m + i] |= (uint8_t)(1u << r);
              avail[(z * R + r) * m + i] = zr->second.available;  // extractResources: Available
            }
            for (auto& [cname, cval] : zl[i][z]->costs) {  // extractCosts: pluginhelpers.go:136-153
              long id = NumaNameToID(cname);
              if (id >= 0 && id < (long)zl[i].size()) cost[(z * Z + id) * m + i] = (int32_t)cval;
            }
          }
        }
      }
      node_flags[i] = fl;
    }
    // ---- the pod
    uint8_t qos = (uint8_t)GetPodQOS(pod), pflags = 0;
    if (GetPodQOS(pod) == QOS::BestEffort && !IncludeNonNative(pod)) pflags |= B200S_NRT_POD_FILTER_BYPASS;
    std::vector<const Container*> conts;
    for (auto& ct : pod.init_containers) conts.push_back(&ct);
    for (auto& ct : pod.containers) conts.push_back(&ct);
    uint8_t n_init = (uint8_t)pod.init_containers.size(), n_app = (uint8_t)pod.containers.size();
    if (conts.size() > (size_t)C_MAX) {
      pflags |= B200S_NRT_POD_UNSUPPORTED;
      conts.clear();
      n_init = n_app = 0;
    }
    std::vector<uint8_t> kind(C_MAX, 0), req_mask(C_MAX + 1, 0);
    std::vector<int64_t> req((size_t)(C_MAX + 1) * R, 0);
    auto put = [&](int cidx, const ResourceList& rl) {
      for (auto& [nm, q] : rl) {
        int r = slot[nm];
        if (r >= R) {
          pflags |= B200S_NRT_POD_UNSUPPORTED;
          continue;
        }
        req_mask[cidx] |= (uint8_t)(1u << r);
        req[(size_t)cidx * R + r] = q;
      }
    };
    for (size_t ci = 0; ci < conts.size(); ++ci) {
      kind[ci] = ci < n_init ? (conts[ci]->restart_always ? B200S_CONT_SIDECAR : B200S_CONT_INIT) : B200S_CONT_APP;
      put((int)ci, conts[ci]->requests);
    }
    put(C_MAX, GetPodEffectiveRequest(pod));
    // ---- engine
    b200s_nrt_nodes nn;
    nn.n_zones = Z;
    nn.n_res = R;
    nn.res_flags = res_flags.data();
    nn.node_flags = node_flags.data();
    nn.max_numa = max_numa.data();
    nn.n_zones_node = nz.data();
    nn.node_res_mask = node_res_mask.data();
    nn.zone_res_mask = zmask.data();
    nn.avail = avail.data();
    nn.cost = cost.data();
    std::vector<int64_t> w(R, 1);
    for (auto& rs : args_.resources) {
      auto s = slot.find(rs.name);
      if (s != slot.end() && s->second < R) w[s->second] = rs.weight;
    }
    eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n, 0, n), "snapshot_begin");
    // columns above are [..][m] with m == max(n,1): identical to [..][n] for n >= 1
    eng_->Check(b200s_snapshot_nrt(eng_->ctx(), &nn), "snapshot_nrt");
    eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
    eng_->Check(b200s_config_nrt(eng_->ctx(), strategy_, R, w.data()), "config_nrt");
    b200s_nrt_pods np;
    np.qos = &qos;
    np.flags = &pflags;
    np.n_init = &n_init;
    np.n_app = &n_app;
    np.cont_kind = kind.data();
    np.req_mask = req_mask.data();
    np.req = req.data();
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.nrt = &np;
    c->scores.assign(npad, 0);
    c->feasible.assign(npad / 64, 0);
    c->reasons.assign(npad, 0);
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_NRT, &b, B200S_OUT_U8, c->scores.data(), c->feasible.data(),
                                  c->reasons.data()),
                "score_batch(NodeResourceTopologyMatch)");
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  return c;
}

Status TopologyMatch::Filter(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return ErrorStatus("node not found");  // filter.go:177-179
  auto c = Run(state, pod);
  if (!c->engine_error.empty()) return ErrorStatus(c->engine_error);
  auto it = c->index.find(nodeInfo.GetNode()->name);
  if (it == c->index.end()) return ErrorStatus("node not in the cycle's snapshot");
  switch (c->reasons[it->second]) {
    case B200S_REASON_OK: return {};
    case B200S_REASON_NRT_INVALID_TOPOLOGY: return {Code::Unschedulable, "invalid node topology data"};
    case B200S_REASON_NRT_ALIGN_POD: return {Code::Unschedulable, "cannot align pod"};
    case B200S_REASON_NRT_ALIGN_CONTAINER: return {Code::Unschedulable, "cannot align container"};
    case B200S_REASON_NRT_ALIGN_INIT: return {Code::Unschedulable, "cannot align init container"};
    case B200S_REASON_NRT_ALIGN_SIDECAR: return {Code::Unschedulable, "cannot align sidecar container"};
    case B200S_REASON_NRT_ACCOUNTING: return {Code::Error, "inconsistent resource accounting"};
    case B200S_REASON_UNSUPPORTED:
      // shape outside the dense encoding (NUMA ids not 0..k-1 in order, > 8 zones / resources / containers): this
      // pair is answered by the scalar path on the host objects -- what the embedded Go plugin is in the Go shim.
      // The reference's gates up to here (fresh, NRT present, single-numa-node) already passed in the kernel.
      try {
        return ScalarFilter(pod, nodeInfo, *h_->nrts.at(nodeInfo.GetNode()->name));
      } catch (const std::exception& e) {
        return ErrorStatus(e.what());
      }
    default: return ErrorStatus("unexpected reason code");
  }
}

std::pair<int64_t, Status> TopologyMatch::Score(CycleState& state, const Pod& pod, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return {0, ErrorStatus("node not found")};
  auto c = Run(state, pod);
  if (!c->engine_error.empty()) return {0, ErrorStatus(c->engine_error)};
  auto it = c->index.find(nodeInfo.GetNode()->name);
  if (it == c->index.end()) return {0, ErrorStatus("node not in the cycle's snapshot")};
  if (c->reasons[it->second] == B200S_REASON_UNSUPPORTED ||
      (c->unsupported_nodes.count(it->second) && GetPodQOS(pod) == QOS::Guaranteed)) {
    // outside the dense encoding: the scalar path (score.go:88-101 after the QoS / freshness / nil-NRT gates)
    try {
      if (GetPodQOS(pod) != QOS::Guaranteed) return {100, Status{}};  // score.go:72-75 comes before everything
      std::map<std::string, int64_t> w;
      for (auto& rs : args_.resources) w[rs.name] = rs.weight;
      auto nf = h_->nrt_not_fresh.find(nodeInfo.GetNode()->name);
      if (nf != h_->nrt_not_fresh.end() && nf->second) return {0, Status{}};
      return {ScalarScore(pod, *h_->nrts.at(nodeInfo.GetNode()->name), strategy_, w), Status{}};
    } catch (const std::exception& e) {
      return {0, ErrorStatus(e.what())};
    }
  }
  // upstream only scores nodes that passed every filter; a node this plugin rejected has no score
  return {(int64_t)c->scores[it->second], Status{}};
}

// =============================================================== NetworkOverhead
std::unique_ptr<NetworkOverhead> NetworkOverhead::New(const NetworkOverheadArgs& args, std::shared_ptr<Handle> h) {
  std::unique_ptr<NetworkOverhead> p(new NetworkOverhead());
  p->h_ = std::move(h);
  p->args_ = args;
  p->eng_.reset(new Engine(p->h_->device));
  p->eng_->Check(b200s_config_network_overhead(p->eng_->ctx(), 1, 1), "config_network_overhead");
  return p;
}

namespace {

struct NetohState : CycleResult {
  std::vector<int64_t> raw;          // finalCostMap
  std::vector<uint32_t> counts;      // satisfied | violated << 16
  std::string status_message;
  std::vector<b200s_netoh_dep> deps;
  int32_t npad = 0;
};

// FindTopologyKey / FindOriginCosts: pkg/networkaware/util/util.go:156-191 (binary searches)
const std::vector<OriginInfo>* FindTopologyKey(const std::vector<TopologyInfo>& list, const std::string& key) {
  int low = 0, high = (int)list.size() - 1;
  while (low <= high) {
    int mid = (low + high) / 2;
    if (list[mid].topology_key == key) return &list[mid].origin_list;
    if (list[mid].topology_key < key) low = mid + 1; else high = mid - 1;
  }
  return nullptr;
}
const std::vector<CostInfo>* FindOriginCosts(const std::vector<OriginInfo>& list, const std::string& origin) {
  int low = 0, high = (int)list.size() - 1;
  while (low <= high) {
    int mid = (low + high) / 2;
    if (list[mid].origin == origin) return &list[mid].cost_list;
    if (list[mid].origin < origin) low = mid + 1; else high = mid - 1;
  }
  return nullptr;
}

}  // namespace

std::shared_ptr<CycleResult> NetworkOverhead::Run(CycleState& state, const Pod& pod,
                                                  const std::vector<std::string>* feasible_names,
                                                  bool list_is_feasible_set) {
  auto c = std::make_shared<NetohState>();
  c->score_equally = true;  // PreFilterState{scoreEqually: true}, :176-178
  try {
    const auto& nodes = h_->node_infos;
    const int32_t n = (int32_t)nodes.size(), npad = NPad(n), m = std::max(n, 1);
    c->npad = npad;
    c->index = IndexOf(nodes);
    c->scores.assign(npad, 0);
    c->feasible.assign(npad / 64, 0);
    c->reasons.assign(npad, 0);
    c->raw.assign(npad, 0);
    c->counts.assign(npad, 0);
    auto lbl = [](const std::map<std::string, std::string>& l, const char* k) {
      auto it = l.find(k);
      return it == l.end() ? std::string() : it->second;
    };
    const std::string ag_name = lbl(pod.labels, AppGroupLabel);
    if (ag_name.empty()) {
      c->status_message = "Pod does not belong to an AppGroup, return";  // :185-188
      return c;
    }
    auto agi = h_->app_groups.find(ag_name);  // findAppGroupNetworkOverhead :654-673
    auto nti = h_->network_topologies.find(args_.network_topology_name);
    const AppGroup* ag = agi == h_->app_groups.end() ? nullptr : agi->second.get();
    NetworkTopology nt = nti == h_->network_topologies.end() || !nti->second ? NetworkTopology{} : *nti->second;
    const bool netperf = args_.weights_name == "NetperfCosts";
    if (!netperf)  // sortNetworkTopologyCosts :438-445
      for (auto& w : nt.weights)
        std::sort(w.topology_list.begin(), w.topology_list.end(),
                  [](const TopologyInfo& a, const TopologyInfo& b) { return a.topology_key < b.topology_key; });
    // GetDependencyList: util.go:194-212
    std::vector<DependencyInfo> dependency_list;
    const std::string selector = lbl(pod.labels, AppGroupSelectorLabel);
    if (ag)
      for (auto& w : ag->workloads)
        if (w.selector == selector)
          for (auto& d : w.dependencies) dependency_list.push_back(d);
    if (dependency_list.empty()) {
      c->status_message = "Pod has no dependencies, return";  // :203-205
      return c;
    }
    // pods of the AppGroup from the lister, then GetScheduledList: util.go:215-231
    std::vector<const Pod*> scheduled;
    size_t ag_pods = 0;
    for (auto& p : h_->pods)
      if (lbl(p->labels, AppGroupLabel) == ag_name) {
        ++ag_pods;
        if (!p->node_name.empty()) scheduled.push_back(p.get());
      }
    if (ag_pods == 0) {
      c->status_message = "No pods yet allocated, return";  // :214-217
      return c;
    }
    if (scheduled.empty()) {
      c->status_message = "Scheduled list is empty, return";  // :222-225
      return c;
    }
    // ---- name dictionary shared by region and zone label values (they share the costMap namespace)
    std::map<std::string, uint16_t> dict = {{"", 0}};
    auto id_of = [&](const std::string& s) {
      auto it = dict.find(s);
      if (it != dict.end()) return it->second;
      uint16_t v = (uint16_t)dict.size();
      dict[s] = v;
      return v;
    };
    std::vector<uint16_t> rid(m, 0), zid(m, 0);
    for (int32_t i = 0; i < n; ++i) {
      const Node* nd = nodes[i].GetNode();
      if (!nd) continue;
      rid[i] = id_of(lbl(nd->labels, LabelTopologyRegion));
      zid[i] = id_of(lbl(nd->labels, LabelTopologyZone));
    }
    for (auto& w : nt.weights)
      for (auto& t : w.topology_list)
        for (auto& o : t.origin_list) {
          id_of(o.origin);
          for (auto& ci : o.cost_list) id_of(ci.destination);
        }
    const int K = (int)dict.size();
    std::vector<int64_t> zc((size_t)K * K, B200S_NETOH_MISSING), rc((size_t)K * K, B200S_NETOH_MISSING);
    // populateCostMap (:448-497) for every distinct origin label value
    for (auto& w : nt.weights) {
      if (w.name != args_.weights_name) continue;
      for (int pass = 0; pass < 2; ++pass) {
        const char* key = pass == 0 ? LabelTopologyRegion : LabelTopologyZone;
        const std::vector<OriginInfo>* ol = FindTopologyKey(w.topology_list, key);
        std::vector<OriginInfo> sorted = ol ? *ol : std::vector<OriginInfo>{};
        if (!netperf)
          std::sort(sorted.begin(), sorted.end(), [](const OriginInfo& a, const OriginInfo& b) { return a.origin < b.origin; });
        std::vector<int64_t>& mat = pass == 0 ? rc : zc;
        std::set<uint16_t> origins(pass == 0 ? rid.begin() : zid.begin(), pass == 0 ? rid.end() : zid.end());
        for (auto& [name, oid] : dict) {
          if (name.empty() || !origins.count(oid)) continue;
          const std::vector<CostInfo>* costs = FindOriginCosts(sorted, name);
          if (!costs) continue;
          for (auto& ci : *costs) mat[(size_t)oid * K + dict[ci.destination]] = ci.network_cost;
        }
      }
    }
    // ---- (placed pod, dependency) pairs in scheduledList x dependencyList order
    for (const Pod* sp : scheduled) {
      const std::string sel = lbl(sp->labels, AppGroupSelectorLabel);
      for (auto& d : dependency_list) {
        if (sel != d.selector) continue;
        auto hi = c->index.find(sp->node_name);
        if (hi == c->index.end())
          throw std::runtime_error("getting pod hostname from Snapshot: nodeinfo not found for node name \"" + sp->node_name + "\"");
        b200s_netoh_dep e;
        e.host_node = hi->second;
        e.host_region = rid[hi->second];
        e.host_zone = zid[hi->second];
        e.max_network_cost = d.max_network_cost;
        c->deps.push_back(e);
      }
    }
    c->score_equally = false;
    // ---- engine
    eng_->Check(b200s_snapshot_begin(eng_->ctx(), h_->generation, n, 0, n), "snapshot_begin");
    eng_->Check(b200s_snapshot_network_overhead(eng_->ctx(), rid.data(), zid.data(), K, zc.data(), rc.data()),
                "snapshot_network_overhead");
    eng_->Check(b200s_snapshot_commit(eng_->ctx()), "snapshot_commit");
    uint8_t eq = 0;
    int32_t off[2] = {0, (int32_t)c->deps.size()};
    b200s_netoh_pods np;
    np.score_equally = &eq;
    np.dep_offset = off;
    np.deps = c->deps.empty() ? nullptr : c->deps.data();
    b200s_pod_batch b;
    memset(&b, 0, sizeof(b));
    b.n_pods = 1;
    b.netoh = &np;
    std::vector<uint64_t> words;
    if (feasible_names) {
      words = FeasibleWords(c->index, npad, *feasible_names);
      b.feasible = words.data();
    }
    // PreFilter/Filter: the upstream cycle (own verdict ANDed in).  NormalizeScore(list): exactly the list.
    eng_->Check(b200s_config_network_overhead(eng_->ctx(), 1, list_is_feasible_set ? 0 : 1), "config_network_overhead");
    eng_->Check(b200s_score_batch(eng_->ctx(), B200S_PLUGIN_NETWORK_OVERHEAD, &b, B200S_OUT_U8, c->scores.data(),
                                  c->feasible.data(), c->reasons.data()),
                "score_batch(NetworkOverhead)");
    eng_->Check(b200s_fetch_network_overhead_raw(eng_->ctx(), c->raw.data(), c->raw.size() * 8), "fetch raw");
    eng_->Check(b200s_fetch_network_overhead_counts(eng_->ctx(), c->counts.data(), c->counts.size() * 4), "fetch counts");
    c->status_message = "PreFilter State updated";
  } catch (const std::exception& e) {
    c->engine_error = e.what();
  }
  (void)state;
  return c;
}

static const char* kNetohKey = "PreFilterNetworkOverhead";  // preFilterStateKey, networkoverhead.go:61

Status NetworkOverhead::PreFilter(CycleState& state, const Pod& pod, const std::vector<NodeInfo>&) {
  auto c = Run(state, pod, nullptr, false);
  state.data[kNetohKey] = c;
  if (!c->engine_error.empty()) return ErrorStatus(c->engine_error);
  return Status{Code::Success, std::static_pointer_cast<NetohState>(c)->status_message};
}

Status NetworkOverhead::Filter(CycleState& state, const Pod&, const NodeInfo& nodeInfo) {
  if (!nodeInfo.GetNode()) return ErrorStatus("node not found");  // :330-332
  auto c = Read<NetohState>(state, kNetohKey);
  if (!c) return ErrorStatus("not eligible due to failed to read from cycleState");  // :336-340
  if (!c->engine_error.empty()) return ErrorStatus(c->engine_error);
  if (c->score_equally) return {};  // :343-346
  auto it = c->index.find(nodeInfo.GetNode()->name);
  if (it == c->index.end()) return ErrorStatus("node not in the cycle's snapshot");
  if (c->reasons[it->second] == B200S_REASON_NETOH_VIOLATED) {
    const uint32_t cv = c->counts[it->second];
    return {Code::Unschedulable, "Node " + nodeInfo.GetNode()->name +
                                     " does not meet several network requirements from Workload dependencies: Satisfied: " +
                                     std::to_string(cv & 0xffff) + " Violated: " + std::to_string(cv >> 16)};  // :353-357
  }
  return {};
}

std::pair<int64_t, Status> NetworkOverhead::Score(CycleState& state, const Pod&, const NodeInfo& nodeInfo) {
  auto c = Read<NetohState>(state, kNetohKey);
  if (!c) return {0, ErrorStatus("not eligible due to failed to read from cycleState, return min score")};  // :370-374
  if (!c->engine_error.empty()) return {0, ErrorStatus(c->engine_error)};
  if (c->score_equally) return {0, Status{Code::Success, "scoreEqually enabled: minimum score"}};  // :377-379
  auto it = c->index.find(nodeInfo.GetNode()->name);
  if (it == c->index.end()) return {0, ErrorStatus("node not in the cycle's snapshot")};
  // finalCostMap[nodeName]: the accumulated cost; NormalizeScore inverts it (:382-385)
  return {c->raw[it->second], Status{Code::Success, "Accumulated cost added as score, normalization ensures lower costs are favored"}};
}

Status NetworkOverhead::NormalizeScore(CycleState& state, const Pod& pod, std::vector<NodeScore>& scores) {
  auto c = Read<NetohState>(state, kNetohKey);
  if (!c) return ErrorStatus("not eligible due to failed to read from cycleState");
  if (c->score_equally) return {};  // every score is 0: getMinMaxScores -> (0,0) -> early return (:400-402)
  // the list IS the feasible set the normalisation runs over (:397-415): evaluate with exactly these nodes
  std::vector<std::string> names;
  for (auto& s : scores) names.push_back(s.name);
  auto r = Run(state, pod, &names, true);
  if (!r->engine_error.empty()) return ErrorStatus(r->engine_error);
  for (auto& s : scores) {
    auto it = r->index.find(s.name);
    if (it == r->index.end()) return ErrorStatus("node not in the cycle's snapshot: " + s.name);
    // nodes the plugin's own Filter rejects are not in upstream's list; if a caller passes one anyway the
    // engine leaves it unscored
    s.score = (int64_t)r->scores[it->second];
  }
  return {};
}

}  // namespace b200host

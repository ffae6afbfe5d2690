// Engine context of libb200sched (internal; the public surface is include/b200sched.h).
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace b200s {

// Grow-only device buffer: a scheduling cycle re-uses the previous cycle's allocation.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool owned = true;  // false: a view into another allocation (the pod arena); never freed here
  void view(void* ptr) {
    if (owned && p) cudaFree(p);
    p = ptr;
    cap = 0;
    owned = false;
  }
  cudaError_t ensure(size_t bytes) {
    if (!owned) {
      p = nullptr;
      cap = 0;
      owned = true;
    }
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;  // slack so a slightly larger next cycle does not realloc
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

// Double-buffered pinned staging for host->device uploads that must not synchronise the stream: acquire() hands out the
// other buffer (waiting only until the copies queued from it two uploads ago have run), commit() marks the copies queued.
struct PinStage {
  void* p[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  int cur = 0;
  cudaError_t acquire(size_t bytes, void** out) {
    cur ^= 1;
    cudaError_t e = cudaSuccess;
    if (used[cur]) e = cudaEventSynchronize(ev[cur]);
    if (e != cudaSuccess) return e;
    if (!ev[cur]) e = cudaEventCreateWithFlags(&ev[cur], cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    if (bytes > cap[cur]) {
      if (p[cur]) cudaFreeHost(p[cur]);
      p[cur] = nullptr;
      cap[cur] = 0;
      e = cudaHostAlloc(&p[cur], bytes + bytes / 4 + 4096, cudaHostAllocDefault);
      if (e != cudaSuccess) return e;
      cap[cur] = bytes + bytes / 4 + 4096;
    }
    *out = p[cur];
    return cudaSuccess;
  }
  cudaError_t commit(cudaStream_t st) {
    used[cur] = true;
    return cudaEventRecord(ev[cur], st);
  }
  void release() {
    for (int i = 0; i < 2; ++i) {
      if (p[i]) cudaFreeHost(p[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      p[i] = nullptr;
      ev[i] = nullptr;
      cap[i] = 0;
      used[i] = false;
    }
  }
};

struct PluginOut {
  DevBuf scores;   // [P][Npad] int64 or u8
  DevBuf feas;     // [P][Npad/64] u64 (own filter AND upstream mask)
  DevBuf reasons;  // [P][Npad] u8 (filter plugins)
  int dtype = 0;
  int P = 0;
  bool valid = false;
  bool has_feas = false;
  bool has_reasons = false;
};

// Per-pod normalisation parameters of NormalizeScore (allocatable.go:143, networkoverhead.go:389)
struct alignas(16) NormParam {
  int64_t lo;      // min over the feasible set
  int64_t range;   // hi - lo (wrapping, as Go computes it)
  uint32_t magic;  // fast path: floor(2^(32+shift)/range) (or 2^32-1 for powers of two)
  uint32_t shift;  // fast path: floor(log2(range))
  uint32_t mode;   // 0: all zero (range == 0 or empty set); 1: 32-bit fast path; 2: generic int64 path
  uint32_t pad;
};

struct Comm;  // NCCL communicator wrapper (comm.cu)
struct Nrt2;  // state of the batched NodeResourceTopologyMatch path (nrt2.cu)

}  // namespace b200s

struct b200s_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  // b200s_score_batch pipelines a large batch in pod chunks: the D2H of chunk i runs here while the main stream
  // already copies chunk i+1's inputs in (PCIe is full duplex) -- created on first use
  cudaStream_t d2h_stream = nullptr;
  cudaEvent_t ev_kernels = nullptr, ev_d2h = nullptr, ev_h2d = nullptr;
  // sharded NormalizeScore: all-reduces run here while the main stream computes (created on first use)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_chunk[4] = {nullptr, nullptr, nullptr, nullptr}, ev_reduced[4] = {nullptr, nullptr, nullptr, nullptr};
  // cross-eval overlap of the sharded Allocatable pre-pass (min/max, exchange, parameters) with the previous eval's
  // P x N kernel: inputs-ready marker on the main stream, per-parity parameter buffers and their hand-over events
  cudaEvent_t ev_inputs = nullptr, ev_params[2] = {nullptr, nullptr}, ev_norm_done[2] = {nullptr, nullptr};
  bool ev_norm_valid[2] = {false, false};
  uint64_t alloc_eval_seq = 0;
  b200s::DevBuf pod_lo_alt, norm_params_alt;  // parity-1 twins of pod_lo / norm_params
  std::mutex mu;
  std::string err;
  uint64_t launches = 0;
  b200s::Comm* comm = nullptr;

  // ---- snapshot ----
  bool snap_open = false, snap_valid = false;
  bool snap_patching = false;  // open through b200s_snapshot_patch_begin: same node list, rows rewritten in place
  uint64_t gen = 0;
  // patch rows travel like the small pod columns: packed into pinned memory, one copy, one scatter launch
  void* patch_stage = nullptr;
  size_t patch_stage_cap = 0;
  b200s::DevBuf patch_dev;
  int N = 0, Npad = 0, node_off = 0, Nglobal = 0;

  // NodeResourcesAllocatable
  bool has_alloc = false;
  int alloc_R = 0;
  b200s::DevBuf alloc_cols;  // [R][Npad] int64
  bool alloc_cfg = false;
  int alloc_mode = 0;
  int alloc_cfg_R = 0;
  int64_t alloc_w[16] = {0};
  uint64_t alloc_cfg_gen = 0;        // bumped on every config change
  uint64_t alloc_prepared_key = ~0ull;  // (snapshot gen, cfg gen) the raw/sorted arrays belong to
  b200s::DevBuf alloc_raw;         // [Npad] int64   raw score (pod independent, allocatable.go:117-127)
  b200s::DevBuf alloc_sorted_raw;  // [N] int64      raw sorted ascending
  b200s::DevBuf alloc_order;       // [N] int32      node index of sorted position
  b200s::DevBuf alloc_iota;        // [N] int32
  b200s::DevBuf sort_tmp;
  uint64_t snap_serial = 0;        // bumped on every commit

  // TargetLoadPacking
  bool has_tlp = false;
  b200s::DevBuf tlp_util, tlp_cap, tlp_missing, tlp_flags;
  bool tlp_cfg = false;
  int64_t tlp_target = 40;
  bool tlp_sane = false;  // every row finite and non-negative: the scores are in 0..100 (a byte table can hold them)

  // LoadVariationRiskBalancing
  bool has_lvrb = false;
  b200s::DevBuf lvrb_f64;  // [4][Npad] cpuAvg cpuStd memAvg memStd
  b200s::DevBuf lvrb_i64;  // [2][Npad] allocCpuMilli allocMemBytes
  b200s::DevBuf lvrb_flags;
  bool lvrb_sane = false;  // every metric finite, allocatable non-negative: no NaN can reach a score
  bool lvrb_cfg = false;
  double lvrb_margin = 1.0, lvrb_sens = 1.0;

  // Peaks
  bool has_peaks = false;
  b200s::DevBuf peaks_util, peaks_cap, peaks_flags, peaks_k;  // peaks_k [2][Npad] f64: k1, k2

  // LowRiskOverCommitment
  bool has_lowrisk = false;
  b200s::DevBuf lowrisk_f64;   // [4][Npad] cpuAvg cpuStd memAvg memStd
  b200s::DevBuf lowrisk_i64;   // [6][Npad] allocCpu allocMem nodeReqCpu nodeReqMem nodeLimCpu nodeLimMem
  b200s::DevBuf lowrisk_flags;
  b200s::DevBuf lowrisk_load;  // [2][Npad] f64 riskLoad (cpu, memory): depends on the node only, derived per snapshot
  uint64_t lowrisk_prepared_key = ~0ull;
  uint64_t lowrisk_cfg_gen = 0;
  bool lowrisk_cfg = false;
  int64_t lowrisk_window = 5;
  double lowrisk_w_cpu = 0.5, lowrisk_w_mem = 0.5;

  // NodeResourceTopologyMatch
  bool has_nrt = false;
  int nrt_Z = 0, nrt_R = 0;
  bool nrt_has_cost = false;
  uint8_t nrt_res_flags[B200S_NRT_MAX_RES] = {0};
  b200s::DevBuf nrt_node_flags, nrt_max_numa, nrt_nz, nrt_node_res_mask, nrt_zone_res_mask, nrt_avail,
      nrt_cost, nrt_perm;  // nrt_perm [Npad] int32: thread slot -> node, nodes grouped by control-flow class
  // [N] host mirror of the sort key behind nrt_perm: control-flow class (flags << 8 | zones) in bits 40..55, then
  // the largest zone's availability of resource slots 0 and 1 (20 bits each, quantised)
  std::vector<uint64_t> nrt_key_h;
  bool nrt_perm_dirty = false;
  bool nrt_cfg = false;
  uint64_t nrt_cfg_gen = 0;  // bumped on every config change
  b200s::Nrt2* nrt2 = nullptr;
  int nrt_strategy = B200S_NRT_LEAST_ALLOCATED;
  int64_t nrt_w[B200S_NRT_MAX_RES] = {1, 1, 1, 1, 1, 1, 1, 1};

  // NetworkOverhead
  bool has_netoh = false;
  int netoh_K = 0;
  b200s::DevBuf netoh_region, netoh_zone, netoh_zone_cost, netoh_region_cost;
  // distinct (region, zone) label pairs of the snapshot: the cost/filter result of a node depends on the node
  // only through its pair (plus the few dependencies hosted on the node itself)
  int netoh_NQ = 0;
  b200s::DevBuf netoh_pair_id, netoh_pair_r, netoh_pair_z;  // [Npad] int32, [NQ] u16, [NQ] u16
  // host mirror of the dictionary, so that a patched node finds (or appends) its pair without a rebuild
  std::vector<uint16_t> netoh_pair_r_h, netoh_pair_z_h;
  std::unordered_map<uint32_t, int32_t> netoh_dict;
  bool netoh_pairs_dirty = false;
  b200s::DevBuf netoh_pair_cost, netoh_pair_sv;              // [P][NQ] int64 cost, u32 satisfied | violated << 16

  // ---- pods ----
  bool pods_valid = false;
  int P = 0;
  // small pod columns travel as ONE host->device copy: packed into a pinned staging buffer, copied into a
  // device arena, and the column buffers become views into it (a P=1 cycle has ~13 tiny columns)
  b200s::DevBuf pods_arena;
  void* pods_stage = nullptr;
  size_t pods_stage_cap = 0;
  bool netoh_attr_set = false;  // netoh_fast4_kernel's dynamic shared-memory limit raised on this device
  bool async_upload = false;    // b200s_config_async_upload: b200s_pods_upload queues and returns
  b200s::PinStage pods_stage2;  // ... through this double-buffered staging block
  // ... and the upstream mask (the one large column) travels on its own stream into the buffer the batch before the
  // previous one used, so that the copy of chunk i + 1 overlaps the kernels of chunk i
  cudaStream_t h2d_stream = nullptr;
  cudaEvent_t ev_mark[2] = {nullptr, nullptr}, ev_copied = nullptr;
  uint64_t upload_seq = 0;
  b200s::DevBuf feasible_alt;
  bool has_feasible = false;
  b200s::DevBuf feasible_in;  // [P][Npad/64]
  // eval_combined chains the filters: each plugin's "upstream" set is what the previous filters left
  bool defer_sync = false;  // inside b200s_score_batch: queue everything, synchronise once at the end
  const uint64_t* mask_override = nullptr;
  const uint64_t* upstream_mask() const {
    if (mask_override) return mask_override;
    return has_feasible ? feasible_in.as<uint64_t>() : nullptr;
  }
  bool has_tlp_pods = false, has_lvrb_pods = false, has_nrt_pods = false, has_netoh_pods = false;
  bool has_peaks_pods = false, has_lowrisk_pods = false;
  b200s::DevBuf tlp_pod_cpu, lvrb_req_cpu, lvrb_req_mem, peaks_pod_cpu, lowrisk_pod;
  b200s::DevBuf nrt_pod_qos, nrt_pod_flags, nrt_pod_ninit, nrt_pod_napp, nrt_pod_kind, nrt_pod_req_mask,
      nrt_pod_req;
  b200s::DevBuf netoh_equal, netoh_dep_off, netoh_deps;
  // distinct pod keys of the batch (score-table path of the Trimaran plugins: pending pods of one Deployment / Job
  // carry identical request columns, so a plugin's row depends on the pod only through a key few pods are alone with)
  int tlp_U = 0, lvrb_U = 0;  // 0 = no dictionary (small batch, or a negative / absurd request)
  std::vector<int64_t> tlp_uniq_h, lvrb_uniq_cpu_h, lvrb_uniq_mem_h;
  std::vector<int32_t> tlp_row_h, lvrb_row_h;
  b200s::DevBuf tlp_uniq, tlp_row, lvrb_uniq_cpu, lvrb_uniq_mem, lvrb_row, score_table;
  int netoh_total_deps = 0, netoh_max_deps = 0;

  // ---- per-pod scratch ----
  b200s::DevBuf pod_lo, pod_hi;  // [P] int64
  b200s::DevBuf norm_params;     // [P] NormParam
  b200s::DevBuf raw_scores;      // [P][Npad] int64 scratch (NetworkOverhead raw cost)
  b200s::DevBuf netoh_counts;    // [P][Npad] u32 satisfied | violated << 16 (opt-in)
  bool netoh_want_counts = false;
  bool netoh_apply_filter = true;
  int netoh_raw_P = -1;          // P of the last NetworkOverhead eval (raw/counts validity)

  // ---- outputs ----
  b200s::PluginOut out[B200S_PLUGIN_COUNT];
  b200s::DevBuf total;       // [P][Npad] int64
  b200s::DevBuf total_feas;  // [P][Npad/64]
  b200s::DevBuf topk_local;  // [P][k] entries of this shard
  b200s::DevBuf topk_all;    // [world][P][k]
  b200s::DevBuf topk_final;  // [P][k]
  b200s::DevBuf topk_slices; // [slices][P][k] per-node-slice winners of a small batch
  void* small_bounce = nullptr;  // one pinned page for small device-to-host results (a cycle's winners)
  void* cycle_cells_base = nullptr;
  bool cycle_cells_zero = false; // the fused cycle's min/max cells are zero (the folding CTA resets them)
  int fused_cycle = 1;          // b200s_config_fused_cycle: 0 off, 1 small batches go through cycle.cu (b200s_schedule_batch as
                                //   one graph launch), 2 the same without the graph
  void* cycle_graph = nullptr;  // cycle.cu: the instantiated graph of b200s_schedule_batch (copy + two kernels)
  bool hold_upload = false;     // b200s_schedule_batch: stage the pod columns, let the graph's copy node move them
  void* held_dst = nullptr;
  const void* held_src = nullptr;
  size_t held_bytes = 0;
  b200s::DevBuf cycle_scratch;  // fused cycle kernel: min/max cells, per-node partial scores, per-block winners
  int topk_k = 0;
  bool total_valid = false, topk_valid = false, feas_valid = false;  // feas_valid: total_feas without a resident top-k

  // ---- harness profiling: event pairs around the dominant kernel of each eval ----
  bool profiling = false;
  // slots 0..B200S_PLUGIN_COUNT-1: plugins; B200S_PLUGIN_COUNT + phase: the B200S_PHASE_* timers
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending[B200S_PLUGIN_COUNT + B200S_PHASE_COUNT];
  std::vector<cudaEvent_t> prof_pool;
  cudaEvent_t prof_get() {
    cudaEvent_t e = nullptr;
    if (!prof_pool.empty()) {
      e = prof_pool.back();
      prof_pool.pop_back();
    } else {
      cudaEventCreate(&e);
    }
    return e;
  }

  int set_err(int code, const std::string& msg) {
    err = msg;
    return code;
  }
};

namespace b200s {

#define B200S_CUDA_TRY(ctx, expr)                                                               \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return (ctx)->set_err(B200S_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define B200S_TRY(expr)     \
  do {                      \
    int _rc = (expr);       \
    if (_rc != B200S_OK) return _rc; \
  } while (0)

// per-plugin launchers (each in its own .cu)
int alloc_eval(b200s_ctx* c, int dtype);
int tlp_eval(b200s_ctx* c, int dtype);
int lvrb_eval(b200s_ctx* c, int dtype);
int nrt_eval(b200s_ctx* c, int dtype);
int netoh_eval(b200s_ctx* c, int dtype);
int peaks_eval(b200s_ctx* c, int dtype);
int lowrisk_eval(b200s_ctx* c, int dtype);
int combined_eval(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, int write_total);
int alloc_prepare(b200s_ctx* c);  // raw scores + sorted order of the snapshot (cached per snapshot / args)
// cycle.cu: the whole cycle in two launches -- one graph launch from b200s_schedule_batch (small P, single GPU)
bool cycle_applies(b200s_ctx* c, uint32_t mask, int k, int write_total, bool any_p = false);
int cycle_eval(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k);
int cycle_graph_run(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, b200s_topk_entry* host_out);
void cycle_graph_free(b200s_ctx* c);
int cycle_sequence(b200s_ctx* c, uint32_t mask, const int64_t* weights, b200s_topk_entry* winners);

// nrt2.cu: batched NodeResourceTopologyMatch path (score tables per distinct request vector + coalesced expansion)
void nrt2_destroy(b200s_ctx* c);
void nrt2_on_snapshot_full(b200s_ctx* c, const b200s_nrt_nodes* nn);
void nrt2_on_patch_rows(b200s_ctx* c, int count, const b200s_nrt_nodes* rows);
void nrt2_on_deduct(b200s_ctx* c, int count, const int64_t* deduct);
void nrt2_on_class_change(b200s_ctx* c);
void nrt2_on_pods(b200s_ctx* c, const b200s_nrt_pods* q, int P);
int nrt2_prepare(b200s_ctx* c);  // 1 = applicable and prepared, 0 = keep the direct kernel, < 0 = error
int nrt2_eval(b200s_ctx* c, int dtype);
void nrt2_set_force(b200s_ctx* c, int path);
int nrt2_last_path(b200s_ctx* c);
const char* nrt2_note(b200s_ctx* c);
void nrt2_note_direct(b200s_ctx* c);
int debug_div_check(b200s_ctx* c, const double* x, const double* d, int n, uint64_t* mismatches);

// comm.cu
int comm_allreduce_minmax(b200s_ctx* c, int64_t* lo, int64_t* hi, int count);  // in place, device, on c->stream
int comm_allreduce_minmax_on(b200s_ctx* c, cudaStream_t stream, int64_t* lo, int64_t* hi, int count);
int comm_ensure_streams(b200s_ctx* c);  // comm_stream + the chunk events
bool comm_has_peers(b200s_ctx* c);    // peer-memory exchange available (b200s_comm_peer_import done)
int comm_allgather(b200s_ctx* c, const void* send, void* recv, size_t bytes_per_rank);
int comm_rank(b200s_ctx* c);
int comm_world(b200s_ctx* c);
void comm_destroy(b200s_ctx* c);

// shared small kernels (norm.cu)
int build_norm_params(b200s_ctx* c, int P);  // pod_lo/pod_hi -> norm_params

int ensure_out(b200s_ctx* c, int plugin, int dtype, bool feas, bool reasons);

// Brackets the dominant kernel of an eval (or a phase: slot B200S_PLUGIN_COUNT + B200S_PHASE_*) with events when
// profiling is on.
struct KernelTimer {
  b200s_ctx* c;
  int plugin;
  cudaStream_t stream;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(b200s_ctx* ctx, int pl, cudaStream_t st = nullptr) : c(ctx), plugin(pl), stream(st ? st : ctx->stream) {
    if (c->profiling) {
      e0 = c->prof_get();
      e1 = c->prof_get();
      cudaEventRecord(e0, stream);
    }
  }
  ~KernelTimer() {
    if (e0) {
      cudaEventRecord(e1, stream);
      c->prof_pending[plugin].push_back({e0, e1});
    }
  }
};

}  // namespace b200s

/*
 * b200sched.h — C-ABI of the B200-native batched Filter/Score engine.
 *
 * This is the drop-in boundary: exactly what a cgo binding in
 * the Go plugins of kubernetes-sigs/scheduler-plugins would call once per
 * scheduling cycle.  Plain pointers and sizes only; no C++ types, no torch
 * types, no callbacks.  Every function returns B200S_OK (0) or a negative
 * error code; b200s_last_error() gives the message.  No exception or abort
 * crosses this boundary.
 *
 * Reference interfaces each entry point stands behind (paths relative to the
 * reference repo):
 *   NodeResourcesAllocatable  Score           pkg/noderesources/allocatable.go:63
 *                             NormalizeScore  pkg/noderesources/allocatable.go:143
 *   TargetLoadPacking         Score           pkg/trimaran/targetloadpacking/targetloadpacking.go:107
 *   LoadVariationRiskBalancing Score          pkg/trimaran/loadvariationriskbalancing/loadvariationriskbalancing.go:84
 *   NodeResourceTopologyMatch Filter          pkg/noderesourcetopology/filter.go:176
 *                             Score           pkg/noderesourcetopology/score.go:62
 *   NetworkOverhead           PreFilter       pkg/networkaware/networkoverhead/networkoverhead.go:174
 *                             Filter          pkg/networkaware/networkoverhead/networkoverhead.go:326
 *                             Score           pkg/networkaware/networkoverhead/networkoverhead.go:362
 *                             NormalizeScore  pkg/networkaware/networkoverhead/networkoverhead.go:389
 *   Peaks                     Score           pkg/trimaran/peaks/peaks.go:103
 *                             NormalizeScore  pkg/trimaran/peaks/peaks.go:152
 *   LowRiskOverCommitment     Score           pkg/trimaran/lowriskovercommitment/lowriskovercommitment.go:105
 *   upstream RunScorePlugins weight/sum + selectHost (restated; not in tree)
 *
 * Data model
 *   One ctx = one GPU = one contiguous shard of the node axis.  The host
 *   flattens the cycle's NodeInfo snapshot into struct-of-arrays columns
 *   (b200s_snapshot_*), uploads a batch of pending pods (b200s_pods_*), and
 *   asks for one plugin (b200s_eval) or the weighted combination with a
 *   per-pod top-k (b200s_eval_combined).  All input pointers are HOST memory,
 *   are read during the call and never retained (cgo rule: C must not keep Go
 *   pointers).  Outputs are engine-owned device matrices that the host fetches
 *   (b200s_fetch_*) into caller-owned buffers.
 *
 * Layout
 *   N  = nodes in this shard, Npad = N rounded up to B200S_NODE_ALIGN.
 *   score matrix  [P][Npad]      int64 (B200S_OUT_I64) or uint8 (B200S_OUT_U8)
 *   feasibility   [P][Npad/64]   uint64 words, bit j of word w = node 64*w+j,
 *                                1 = feasible
 *   reason codes  [P][Npad]      uint8 (B200S_REASON_*), filter plugins only
 *   Scores of infeasible or padding nodes are written as 0.
 */
#ifndef B200SCHED_H
#define B200SCHED_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200S_VERSION 100          /* major*10000 + minor*100 + patch */
#define B200S_NODE_ALIGN 128

/* ---- error codes ------------------------------------------------------- */
#define B200S_OK 0
#define B200S_ERR_INVALID (-1)     /* bad argument / shape */
#define B200S_ERR_CUDA (-2)        /* CUDA runtime error (message has detail) */
#define B200S_ERR_STATE (-3)       /* call out of order (no snapshot, no pods, ...) */
#define B200S_ERR_UNSUPPORTED (-4) /* shape outside the dense encoding: caller falls back */
#define B200S_ERR_NCCL (-5)
#define B200S_ERR_NOMEM (-6)

/* ---- plugins ------------------------------------------------------------ */
typedef enum {
  B200S_PLUGIN_ALLOCATABLE = 0,      /* NodeResourcesAllocatable */
  B200S_PLUGIN_TLP = 1,              /* TargetLoadPacking */
  B200S_PLUGIN_LVRB = 2,             /* LoadVariationRiskBalancing */
  B200S_PLUGIN_NRT = 3,              /* NodeResourceTopologyMatch */
  B200S_PLUGIN_NETWORK_OVERHEAD = 4, /* NetworkOverhead */
  B200S_PLUGIN_PEAKS = 5,            /* Trimaran Peaks */
  B200S_PLUGIN_LOW_RISK = 6,         /* Trimaran LowRiskOverCommitment */
  B200S_PLUGIN_COUNT = 7
} b200s_plugin;

typedef enum { B200S_OUT_I64 = 0, B200S_OUT_U8 = 1 } b200s_out_dtype;

/* NodeResourcesAllocatable modes (apis/config/types.go:45-52) */
#define B200S_ALLOC_LEAST 0
#define B200S_ALLOC_MOST 1

/* NodeResourceTopologyMatch scoring strategies (apis/config/types.go:145-156) */
#define B200S_NRT_MOST_ALLOCATED 0
#define B200S_NRT_BALANCED_ALLOCATION 1
#define B200S_NRT_LEAST_ALLOCATED 2
#define B200S_NRT_LEAST_NUMA_NODES 3

/* Filter reason codes (status code + message of the reference) */
#define B200S_REASON_OK 0
#define B200S_REASON_NRT_INVALID_TOPOLOGY 1 /* Unschedulable "invalid node topology data"      filter.go:196 */
#define B200S_REASON_NRT_ALIGN_POD 2        /* Unschedulable "cannot align pod"                filter.go:169 */
#define B200S_REASON_NRT_ALIGN_CONTAINER 3  /* Unschedulable "cannot align container"          filter.go:65  */
#define B200S_REASON_NRT_ALIGN_INIT 4       /* Unschedulable "cannot align init container"     filter.go:50  */
#define B200S_REASON_NRT_ALIGN_SIDECAR 5    /* Unschedulable "cannot align sidecar container"  filter.go:50  */
#define B200S_REASON_NRT_ACCOUNTING 6       /* Error "inconsistent resource accounting"        filter.go:73  */
#define B200S_REASON_NETOH_VIOLATED 7       /* Unschedulable "Node %v does not meet several network requirements ..." networkoverhead.go:355 */
#define B200S_REASON_UPSTREAM 8             /* infeasible in the caller-supplied mask */
#define B200S_REASON_UNSUPPORTED 9          /* shape outside the dense encoding: host falls back to the Go path */

typedef struct b200s_ctx b200s_ctx;

/* ---- lifecycle ---------------------------------------------------------- */
int b200s_version(void);
/* Creates an engine on CUDA device `device`.  Fails with B200S_ERR_CUDA when no
 * usable GPU exists — there is no CPU fallback. */
int b200s_init(int device, b200s_ctx** out);
void b200s_shutdown(b200s_ctx* ctx);
/* Thread-safe copy of the last error message of this ctx ("" if none). */
const char* b200s_last_error(b200s_ctx* ctx);
/* cudaStream_t the engine launches on (for event timing by the harness). */
void* b200s_stream(b200s_ctx* ctx);
/* Blocks until all queued engine work finished. */
int b200s_sync(b200s_ctx* ctx);
/* Number of engine kernels launched since init (harness bookkeeping). */
uint64_t b200s_launch_count(b200s_ctx* ctx);

/* ---- multi-GPU: node axis sharded, one ctx per rank --------- */
#define B200S_UNIQUE_ID_BYTES 128
int b200s_comm_unique_id(void* out_id /* B200S_UNIQUE_ID_BYTES */);
int b200s_comm_init(b200s_ctx* ctx, const void* id, int rank, int world);
/* Optional, one process per GPU on one NVLink / NVSwitch node: after b200s_comm_init every rank exports the CUDA IPC
 * handle of its symmetric exchange buffer, the host side all-gathers the handles (however it distributed the unique
 * id) and every rank imports the table handles[world][B200S_PEER_HANDLE_BYTES].  From then on the two per-pod
 * exchanges (min/max of the normalising plugins, top-k winners) are remote stores into peer memory + a flag instead
 * of NCCL collectives; payloads above 4 MiB per rank keep NCCL.  Collective: all ranks or none. */
#define B200S_PEER_HANDLE_BYTES 64
int b200s_comm_peer_export(b200s_ctx* ctx, void* out_handle /* B200S_PEER_HANDLE_BYTES */);
int b200s_comm_peer_import(b200s_ctx* ctx, const void* handles /* [world][B200S_PEER_HANDLE_BYTES] */);
int b200s_comm_rank(b200s_ctx* ctx);
int b200s_comm_world(b200s_ctx* ctx);

/* ---- snapshot: per-cycle node columns ------------------------------------ */
/* n_nodes: nodes in this shard; node_offset: global index of local node 0;
 * n_nodes_global: nodes over all shards. */
int b200s_snapshot_begin(b200s_ctx* ctx, uint64_t generation, int32_t n_nodes,
                         int32_t node_offset, int32_t n_nodes_global);

/* NodeResourcesAllocatable: alloc[r] is a column of N int64 in the units of
 * calculateResourceAllocatableRequest (resource_allocation.go:79-100): cpu in
 * milli-cores, memory / ephemeral-storage in bytes, scalars in units. */
int b200s_snapshot_allocatable(b200s_ctx* ctx, int32_t n_res, const int64_t* const* alloc);

/* TargetLoadPacking (targetloadpacking.go:131-167).  flags bit0: node has
 * metrics; bit1: a CPU metric with operator Average|Latest was found (value =
 * the LAST such entry).  cap_milli = Node.Status.Capacity cpu (NOT allocatable).
 * missing_milli = predicted CPU of recently bound pods not yet in the metrics
 * window (flattened by the host from PodAssignEventHandler, handler.go:131). */
#define B200S_TLP_HAS_METRICS 1u
#define B200S_TLP_CPU_FOUND 2u
int b200s_snapshot_tlp(b200s_ctx* ctx, const double* cpu_util_pct, const int64_t* cap_milli,
                       const int64_t* missing_milli, const uint8_t* flags);

/* LoadVariationRiskBalancing (resourcestats.go:45-107).  flags bit0: node has
 * metrics; bit1: CPU data valid; bit2: memory data valid (GetResourceData). */
#define B200S_LVRB_HAS_METRICS 1u
#define B200S_LVRB_CPU_OK 2u
#define B200S_LVRB_MEM_OK 4u
int b200s_snapshot_lvrb(b200s_ctx* ctx, const double* cpu_avg, const double* cpu_std,
                        const double* mem_avg, const double* mem_std,
                        const int64_t* alloc_cpu_milli, const int64_t* alloc_mem_bytes,
                        const uint8_t* flags);

/* Peaks (pkg/trimaran/peaks/peaks.go:103-146).  flags as for TLP (B200S_TLP_HAS_METRICS, B200S_TLP_CPU_FOUND) but
 * cpu_util_pct is the FIRST CPU metric with operator Average|Latest (:117-126 breaks at the first match; TLP keeps
 * the last).  cap_milli = Node.Status.Capacity cpu (:131).  k1, k2 = the node's entry in PeaksArgs.NodePowerModel,
 * 0 when it has none (getPowerModel :193-199; k0 does not enter the score). */
int b200s_snapshot_peaks(b200s_ctx* ctx, const double* cpu_util_pct, const int64_t* cap_milli, const uint8_t* flags,
                         const double* k1, const double* k2);

/* LowRiskOverCommitment (pkg/trimaran/lowriskovercommitment/lowriskovercommitment.go:105-254).  The first seven
 * columns are those of LVRB (GetResourceData averages / deviations, Allocatable cpu milli and memory bytes,
 * B200S_LVRB_* flags).  node_req_* / node_lim_* = requests and limits summed over the pods already on the node,
 * each pod's limits raised to its requests first (GetNodeRequestsAndLimits, resourcestats.go:160-228) -- i.e.
 * NodeRequestMinusPod / NodeLimitMinusPod before the capacity cap. */
int b200s_snapshot_low_risk(b200s_ctx* ctx, const double* cpu_avg, const double* cpu_std, const double* mem_avg,
                            const double* mem_std, const int64_t* alloc_cpu_milli, const int64_t* alloc_mem_bytes,
                            const uint8_t* flags, const int64_t* node_req_cpu_milli, const int64_t* node_req_mem_bytes,
                            const int64_t* node_lim_cpu_milli, const int64_t* node_lim_mem_bytes);

/* NodeResourceTopologyMatch.  Dense padded encoding of the per-node NRT object
 * as createNUMANodeList / TopologyManagerFromNodeResourceTopology see it
 * (pluginhelpers.go:105-161, nodeconfig/topologymanager.go:78-161).  Quantities
 * are exact milli-units (cpu "500m" = 500, memory "1Gi" = 1073741824000). */
#define B200S_NRT_MAX_ZONES 8
#define B200S_NRT_MAX_RES 8
#define B200S_NRT_MAX_CONT 8
#define B200S_NRT_NODE_HAS_NRT 1u      /* an NRT object exists for the node */
#define B200S_NRT_NODE_FRESH 2u        /* CachedNRTInfo.Fresh */
#define B200S_NRT_NODE_SINGLE_NUMA 4u  /* policy == single-numa-node */
#define B200S_NRT_NODE_SCOPE_POD 8u    /* scope == pod (else container) */
#define B200S_NRT_NODE_UNSUPPORTED 16u /* NUMA ids not 0..Z-1 in order, >8 zones, ...: host falls back */
#define B200S_NRT_RES_AFFINE 1u        /* isNUMAAffineResource: cpu, memory, hugepages-* */
#define B200S_NRT_RES_HOST_LEVEL 2u    /* isHostLevelResource: ephemeral-storage, storage, non-native */
typedef struct {
  int32_t n_zones;               /* Z: max zones of any node, 1..8 */
  int32_t n_res;                 /* R: resource slots of the snapshot dictionary, 1..8 */
  const uint8_t* res_flags;      /* [R] B200S_NRT_RES_* */
  const uint8_t* node_flags;     /* [N] B200S_NRT_NODE_* */
  const uint16_t* max_numa;      /* [N] TopologyManager.MaxNUMANodes */
  const uint8_t* n_zones_node;   /* [N] zones of this node */
  const uint8_t* node_res_mask;  /* [N] bit r: resource r reported at node level (util.ResourceList(GetAllocatable())) */
  const uint8_t* zone_res_mask;  /* [Z][N] bit r: zone lists resource r */
  const int64_t* avail;          /* [Z][R][N] zone Available, milli-units */
  const int32_t* cost;           /* [Z][Z][N] Costs[z][z'], -1 = missing; may be NULL unless LeastNUMANodes */
} b200s_nrt_nodes;
int b200s_snapshot_nrt(b200s_ctx* ctx, const b200s_nrt_nodes* nodes);

/* NetworkOverhead.  Region and zone label values share one dictionary of
 * n_names ids (id 0 = empty label), because the reference keeps both in one
 * (origin,destination) map (networkoverhead.go:472-493).  zone_cost / region_cost
 * are [n_names][n_names] int64 with B200S_NETOH_MISSING for absent entries:
 * zone_cost[o][d] = cost listed under topology key zone for origin o,
 * region_cost likewise for topology key region. */
#define B200S_NETOH_MISSING INT64_MIN
int b200s_snapshot_network_overhead(b200s_ctx* ctx, const uint16_t* region_id,
                                    const uint16_t* zone_id, int32_t n_names,
                                    const int64_t* zone_cost, const int64_t* region_cost);

int b200s_snapshot_commit(b200s_ctx* ctx);

/* ---- incremental snapshot: rewrite a few node rows of the resident columns ----
 * Upstream's scheduler cache refreshes its snapshot by per-node generation, the NRT cache carries its
 * own generation (pkg/noderesourcetopology/cache/cache.go:27-39, overreserve.go:101-127) and a bind
 * touches one node's Trimaran bookkeeping (pkg/trimaran/handler.go:131-167): between two cycles only a
 * handful of nodes change.  b200s_snapshot_patch_begin re-opens the COMMITTED snapshot (same node list,
 * same N, same column shapes); each b200s_snapshot_patch_* call rewrites `count` rows of one plugin's
 * columns in place (one host->device copy + one scatter launch); b200s_snapshot_commit closes it and
 * re-derives what depends on the rows (Allocatable's sorted raw scores, NRT's thread permutation,
 * NetworkOverhead's label-pair dictionary).  node_idx[count] are shard-local node indices in [0, N);
 * when an index repeats, the last row wins.  Value arrays hold `count` elements per column, in the
 * layouts of the full calls with N replaced by count.  A plugin whose columns were never uploaded in
 * full returns B200S_ERR_STATE.  The pod batch and the plugin args stay valid; results of earlier evals
 * do not. */
int b200s_snapshot_patch_begin(b200s_ctx* ctx, uint64_t generation);
int b200s_snapshot_patch_allocatable(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, int32_t n_res,
                                     const int64_t* const* alloc);
int b200s_snapshot_patch_tlp(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const double* cpu_util_pct,
                             const int64_t* cap_milli, const int64_t* missing_milli, const uint8_t* flags);
int b200s_snapshot_patch_lvrb(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const double* cpu_avg,
                              const double* cpu_std, const double* mem_avg, const double* mem_std,
                              const int64_t* alloc_cpu_milli, const int64_t* alloc_mem_bytes, const uint8_t* flags);
int b200s_snapshot_patch_peaks(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const double* cpu_util_pct,
                               const int64_t* cap_milli, const uint8_t* flags, const double* k1, const double* k2);
/* a bind / pod deletion changes one node's request and limit sums (GetNodeRequestsAndLimits) */
int b200s_snapshot_patch_low_risk(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const double* cpu_avg,
                                  const double* cpu_std, const double* mem_avg, const double* mem_std,
                                  const int64_t* alloc_cpu_milli, const int64_t* alloc_mem_bytes, const uint8_t* flags,
                                  const int64_t* node_req_cpu_milli, const int64_t* node_req_mem_bytes,
                                  const int64_t* node_lim_cpu_milli, const int64_t* node_lim_mem_bytes);
/* rows->n_zones / n_res must equal the resident snapshot's; rows->res_flags is ignored; rows->cost must be
 * non-NULL iff the snapshot has costs. */
int b200s_snapshot_patch_nrt(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const b200s_nrt_nodes* rows);
/* OverReserve cache (pkg/noderesourcetopology/cache/overreserve.go:101-127, store.go:129-160): a pod assumed on a
 * node is deducted from EVERY zone of that node that lists the resource -- available = available < q ? 0 :
 * available - q -- until the node's NRT is resynced.  deduct [R][count] = the quantities to take off (for several
 * assumed pods: their per-resource sum, which gives the same result as the reference's pod-by-pod loop for
 * non-negative quantities), res_mask[count] bit r = the resource is a key of some assumed pod's request map.
 * Cumulative and in place; a resync is a b200s_snapshot_patch_nrt of the row followed by the deduction of what is
 * still assumed.  node_idx must not repeat. */
int b200s_snapshot_patch_nrt_deduct(b200s_ctx* ctx, int32_t count, const int32_t* node_idx, const uint8_t* res_mask,
                                    const int64_t* deduct);
/* labels only (ids into the resident name dictionary); a changed cost table needs the full call */
int b200s_snapshot_patch_network_overhead(b200s_ctx* ctx, int32_t count, const int32_t* node_idx,
                                          const uint16_t* region_id, const uint16_t* zone_id);

/* ---- plugin args (per ctx = per profile; TLP's package-level globals of the
 * reference, targetloadpacking.go:49-53, become per-instance here) ----------- */
int b200s_config_allocatable(b200s_ctx* ctx, int mode, int32_t n_res, const int64_t* weights);
int b200s_config_tlp(b200s_ctx* ctx, int64_t target_utilization_pct);
int b200s_config_lvrb(b200s_ctx* ctx, double safe_variance_margin, double safe_variance_sensitivity);
/* LowRiskOverCommitmentArgs: SmoothingWindowSize (defaults.go:70, > 0) and RiskLimitWeights[cpu], [memory] in [0, 1] */
int b200s_config_low_risk(b200s_ctx* ctx, int64_t smoothing_window_size, double risk_limit_weight_cpu,
                          double risk_limit_weight_mem);
/* weights[r] per resource slot of the NRT dictionary; values < 1 mean 1 (score.go:49-60) */
int b200s_config_nrt(b200s_ctx* ctx, int strategy, int32_t n_res, const int64_t* weights);
/* NodeResourceTopologyMatch has two evaluation paths with identical results: the DIRECT kernel evaluates every
 * (pod, node) pair from scratch as filter.go / score.go do; the BATCHED path (P >= 32, Least/Most/Balanced, <= 4 zones
 * x <= 4 resource slots, quantities that fit the gcd-scaled 32-bit encoding) builds one score / pod-scope-filter
 * table per DISTINCT request vector of the batch and expands it.  AUTO picks the batched path whenever it applies.
 * The knob exists for the parity tests and A/B timing; b200s_nrt_last_path reports what the last eval ran. */
#define B200S_NRT_PATH_AUTO 0
#define B200S_NRT_PATH_DIRECT 1
#define B200S_NRT_PATH_BATCHED 2 /* also for P < 32; still falls back to DIRECT where the path does not apply */
int b200s_config_nrt_path(b200s_ctx* ctx, int path);
int b200s_nrt_last_path(b200s_ctx* ctx); /* 0 = none yet, else B200S_NRT_PATH_DIRECT / _BATCHED */
/* why the last eval declined the batched path ("" if it ran, or a remark on the variant that ran); a static string,
 * valid for the life of the library */
const char* b200s_nrt_path_note(b200s_ctx* ctx);
/* NetworkOverhead: want_counts != 0 also keeps PreFilterState.satisfiedMap / violatedMap
 * (networkoverhead.go:283-296) for the Filter status message.  apply_own_filter (default 1): the
 * plugin's Filter verdict is ANDed into the feasible set that NormalizeScore runs over, as in the
 * upstream cycle; 0 = normalise over exactly the caller's mask (what NormalizeScore does when it is
 * handed an arbitrary NodeScoreList, networkoverhead.go:389-418) — verdicts are still reported. */
int b200s_config_network_overhead(b200s_ctx* ctx, int want_counts, int apply_own_filter);

/* ---- pod batch ------------------------------------------------------------ */
#define B200S_QOS_GUARANTEED 0
#define B200S_QOS_BURSTABLE 1
#define B200S_QOS_BEST_EFFORT 2
#define B200S_NRT_POD_FILTER_BYPASS 1u /* BestEffort && !IncludeNonNative: Filter passes (filter.go:181) */
#define B200S_NRT_POD_UNSUPPORTED 2u   /* >8 containers or resource outside the dictionary: host falls back */
#define B200S_CONT_APP 0
#define B200S_CONT_INIT 1
#define B200S_CONT_SIDECAR 2

typedef struct {
  const uint8_t* qos;            /* [P] B200S_QOS_* */
  const uint8_t* flags;          /* [P] B200S_NRT_POD_* */
  const uint8_t* n_init;         /* [P] init containers (first in the container list) */
  const uint8_t* n_app;          /* [P] app containers; n_init + n_app <= 8, else B200S_ERR_INVALID (flag the pod UNSUPPORTED with zero counts) */
  const uint8_t* cont_kind;      /* [P][C]   B200S_CONT_* */
  const uint8_t* req_mask;       /* [P][C+1] bit r: resource r is a key of the container's Requests; slot C = pod effective request */
  const int64_t* req;            /* [P][C+1][R] milli-units; slot C = GetPodEffectiveRequest (pkg/util/resource.go:51) */
} b200s_nrt_pods;                /* C = B200S_NRT_MAX_CONT, R = snapshot n_res */

typedef struct {
  int32_t host_node;             /* GLOBAL node index of the placed pod's host */
  uint16_t host_region;          /* name ids of that host's labels, < n_names of the snapshot (validated at upload) */
  uint16_t host_zone;
  int64_t max_network_cost;      /* DependenciesInfo.MaxNetworkCost */
} b200s_netoh_dep;               /* one (placed pod, matching dependency) pair; 16 bytes */

typedef struct {
  const uint8_t* score_equally;  /* [P] PreFilterState.scoreEqually */
  const int32_t* dep_offset;     /* [P+1] CSR offsets into deps */
  const b200s_netoh_dep* deps;   /* [dep_offset[P]] in scheduledList x dependencyList order */
} b200s_netoh_pods;

typedef struct {
  int32_t n_pods;
  /* Upstream feasibility (result of the filters that ran before, in-tree ones
   * included), [P][Npad/64] words for THIS shard, or NULL = every node feasible.
   * NormalizeScore min/max run over this set (allocatable.go:145-155). */
  const uint64_t* feasible;
  const int64_t* tlp_pod_cpu_milli;   /* [P] sum PredictUtilisation + overhead (targetloadpacking.go:122-129) or NULL */
  const int64_t* lvrb_req_cpu_milli;  /* [P] GetResourceRequested (resourcestats.go:110) or NULL */
  const int64_t* lvrb_req_mem_bytes;  /* [P] or NULL */
  const b200s_nrt_pods* nrt;          /* or NULL */
  const b200s_netoh_pods* netoh;      /* or NULL */
  const int64_t* peaks_pod_cpu_milli; /* [P] GetResourceRequestQuantity(pod, cpu).MilliValue() (peaks.go:113-114) or NULL */
  /* [4][P]: request cpu milli, request memory bytes, limit cpu milli, limit memory bytes of the pending pod
   * (CreatePodResourcesStateData, lowriskovercommitment.go:257-267: limits raised to requests) or NULL */
  const int64_t* low_risk_pod;
} b200s_pod_batch;

int b200s_pods_upload(b200s_ctx* ctx, const b200s_pod_batch* batch);

/* ---- evaluation ------------------------------------------------------------ */
/* One plugin over all uploaded pods x all nodes of the shard, normalised as the
 * plugin's NormalizeScore does.  Filter plugins (NRT, NetworkOverhead) also
 * produce their feasibility words and reason codes.  The feasible set used for
 * normalisation is (upstream mask) AND (this plugin's own filter).  With a
 * communicator, per-pod min/max are all-reduced across shards first. */
int b200s_eval(b200s_ctx* ctx, b200s_plugin plugin, b200s_out_dtype dtype);

/* Fetch results of the last b200s_eval(plugin).  Buffers are caller-owned host
 * memory of the sizes given in the header comment. */
int b200s_fetch_scores(b200s_ctx* ctx, b200s_plugin plugin, void* out, size_t bytes);
int b200s_fetch_feasible(b200s_ctx* ctx, b200s_plugin plugin, uint64_t* out, size_t bytes);
int b200s_fetch_reasons(b200s_ctx* ctx, b200s_plugin plugin, uint8_t* out, size_t bytes);
/* NetworkOverhead's PreFilterState after the last eval of that plugin: finalCostMap as
 * [P][Npad] int64 (what Score returns before NormalizeScore, networkoverhead.go:383), and — if
 * enabled — [P][Npad] uint32 = satisfied | violated << 16. */
int b200s_fetch_network_overhead_raw(b200s_ctx* ctx, int64_t* out, size_t bytes);
int b200s_fetch_network_overhead_counts(b200s_ctx* ctx, uint32_t* out, size_t bytes);
/* Device pointers of the same matrices (valid until the next eval of that plugin). */
void* b200s_device_scores(b200s_ctx* ctx, b200s_plugin plugin);
uint64_t* b200s_device_feasible(b200s_ctx* ctx, b200s_plugin plugin);

/* Weighted combination of the enabled plugins (upstream RunScorePlugins:
 * feasible = AND of the filters, total = sum weight_p * score_p) and per-pod
 * top-k under (total desc, global node index asc).  With a communicator the
 * per-shard winners are exchanged with ONE ncclAllGather and folded, so every
 * rank returns the global top-k. */
typedef struct {
  int64_t score;
  int32_t node;                  /* GLOBAL node index, -1 = no feasible node */
  int32_t pad;
} b200s_topk_entry;              /* 16 bytes */
int b200s_eval_combined(b200s_ctx* ctx, uint32_t plugin_mask,
                        const int64_t* weights /* [B200S_PLUGIN_COUNT] */, int32_t k,
                        int write_total_matrix);
int b200s_fetch_topk(b200s_ctx* ctx, b200s_topk_entry* out, size_t bytes); /* [P][k] */
/* Small batches (<= 4 pods -- the real scheduler runs one pod per cycle) on a single GPU go through TWO launches
 * that do the whole cycle (filters, scores, NormalizeScore, weighted sum, top-k) without materialising any
 * per-plugin matrix; b200s_fetch_total_feasible still works, b200s_fetch_scores of the individual plugins does not.
 * on = 0 keeps the plugin-by-plugin path for every batch size (parity tests, A/B timing).  Default on. */
int b200s_config_fused_cycle(b200s_ctx* ctx, int on);
/* on = 1: b200s_pods_upload queues its copies and returns without synchronising the stream, so the host can prepare
 * the next pod chunk while the device evaluates this one (a 50k-pod queue goes through the engine in chunks).  The
 * caller then keeps every column buffer of the batch that is larger than 4 MiB unchanged until the next synchronising
 * call (any b200s_fetch_*, b200s_score_batch, b200s_schedule_*); smaller columns are staged at the call.  Default 0. */
int b200s_config_async_upload(b200s_ctx* ctx, int on);
/* The engine-only profile in ONE call with ONE synchronisation: upload the batch's pod columns (HOST pointers),
 * evaluate the weighted combination, copy the [n_pods][k] winners to topk_out (HOST). */
int b200s_schedule_batch(b200s_ctx* ctx, const b200s_pod_batch* batch, uint32_t plugin_mask,
                         const int64_t* weights /* [B200S_PLUGIN_COUNT] */, int32_t k, b200s_topk_entry* topk_out);
/* Speculative placement of a whole batch, pod by pod, without leaving the device: the cycle of pod i
 * runs on the snapshot as pods 0..i-1 left it; after each cycle the winner is ASSUMED on the device --
 * NodeResourceTopologyMatch: the pod's effective request comes off every zone of the winner node that lists the
 * resource (OverReserve cache: overreserve.go:148-182 -> store.go:101-160); TargetLoadPacking: the node's missing
 * utilisation grows by the pod's predicted CPU (handler.go:131-167) -- and the next pod's cycle sees it.  One call, one
 * synchronisation, winners_out[n_pods] (node = -1: unschedulable).  The resident snapshot columns are modified IN
 * PLACE: when a bind later fails the caller resyncs that node's rows (b200s_snapshot_patch_*).  NetworkOverhead's
 * dependency entries are those of the upload (pods of the batch do not become each other's placed dependencies).
 * Single GPU, <= 4 zones x <= 4 resource slots, Least/Most/BalancedAllocation. */
int b200s_schedule_sequence(b200s_ctx* ctx, const b200s_pod_batch* batch, uint32_t plugin_mask,
                            const int64_t* weights /* [B200S_PLUGIN_COUNT] */, b200s_topk_entry* winners_out);
int b200s_fetch_total(b200s_ctx* ctx, int64_t* out, size_t bytes);          /* [P][Npad] */
int b200s_fetch_total_feasible(b200s_ctx* ctx, uint64_t* out, size_t bytes);

/* ---- the per-call convenience the Go shim uses from PreScore ---------------- */
/* upload + eval + fetch in one call, HOST buffers in and out.  scores_out has
 * P*Npad elements of `dtype`; feasible_out / reasons_out may be NULL.
 * A large batch (>= 96 MB of scores) of a score-only plugin (Allocatable, TLP, LVRB, Peaks) with both optional
 * outputs NULL is pipelined in pod chunks -- the D2H of one chunk overlaps the H2D of the next -- and leaves NO
 * engine-resident result behind: b200s_fetch_* and b200s_device_* fail until the next b200s_eval. */
int b200s_score_batch(b200s_ctx* ctx, b200s_plugin plugin, const b200s_pod_batch* batch,
                      b200s_out_dtype dtype, void* scores_out, uint64_t* feasible_out,
                      uint8_t* reasons_out);

/* Pinned host allocations for the caller's staging buffers (cgo cannot hand Go
 * memory to async copies). */
void* b200s_alloc_pinned(size_t bytes);
void b200s_free_pinned(void* p);

/* Harness support: with profiling on, every b200s_eval brackets its DOMINANT kernel (the
 * P x N pass) with CUDA events on the engine stream.  b200s_kernel_time returns the summed
 * duration and the number of launches since the last reset, then resets.  Synchronises. */
int b200s_set_profiling(b200s_ctx* ctx, int on);
int b200s_kernel_time(b200s_ctx* ctx, b200s_plugin plugin, double* total_ms, uint64_t* launches);
/* Same bookkeeping for the phases of a sharded / combined evaluation that are not one plugin's kernel: the per-pod
 * min/max all-reduce of the normalising plugins, the all-gather of the per-pod top-k winners, and the weighted sum +
 * top-k + fold kernels of b200s_eval_combined. */
#define B200S_PHASE_ALLREDUCE 0
#define B200S_PHASE_ALLGATHER 1
#define B200S_PHASE_COMBINE 2
#define B200S_PHASE_COUNT 3
int b200s_phase_time(b200s_ctx* ctx, int phase, double* total_ms, uint64_t* count);

/* Test hook: the Trimaran kernels divide by per-node / per-launch invariants with a hoisted
 * reciprocal + FMA residual correction; this counts inputs where that differs (bitwise) from the
 * IEEE division x[i]/d[i] on the device.  Must be 0. */
int b200s_debug_div_check(b200s_ctx* ctx, const double* x, const double* d, int32_t n, uint64_t* mismatches);

/* Padded node count of the current snapshot (row pitch of every matrix). */
int32_t b200s_npad(b200s_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200SCHED_H */

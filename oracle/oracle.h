/*
 * oracle.h — CPU restatement of the reference's per-(pod,node) Filter/Score arithmetic.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product
 * (scheduler-plugins_b200/, libb200sched.so) never links, imports or calls it.
 *
 * Plain scalar C, single-threaded, written to follow the reference's Go code statement by
 * statement (file:line cited at each function; paths relative to the reference repo
 * kubernetes-sigs/scheduler-plugins @ 2c75c8b).  Go is not installed in this image, so the
 * reference itself cannot be executed: the oracle is pinned instead against the reference's
 * own unit-test vectors (the JSON files under tests/golden, replayed by tests/test_oracle_golden.py).
 *
 * Go semantics restated here: int64 wraps, `/` truncates toward zero, int64(float64)
 * truncates, math.Round rounds half away from zero, float64 arithmetic is never fused
 * (build with -ffp-contract=off).
 */
#ifndef B200S_ORACLE_H
#define B200S_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Go integer helpers ---- */
int64_t orc_wrap_add(int64_t a, int64_t b);
int64_t orc_wrap_sub(int64_t a, int64_t b);
int64_t orc_wrap_mul(int64_t a, int64_t b);
int64_t orc_go_div(int64_t x, int64_t y);

/* ---- NodeResourcesAllocatable (pkg/noderesources) ---- */
/* resourceScorer + score, allocatable.go:117-140.  alloc[r], w[r] for the R configured resources. */
int64_t orc_alloc_score(const int64_t* alloc, const int64_t* w, int R, int mode);
/* NormalizeScore over one pod's feasible list, in place, allocatable.go:143-168. */
void orc_alloc_normalize(int64_t* scores, int n);
/* All pods x all nodes: cols[r][n]; feasible [P][words] or NULL; out [P][pitch] (0 where infeasible). */
void orc_alloc_batch(const int64_t* const* cols, int R, int N, const int64_t* w, int mode, int P,
                     const uint64_t* feasible, int words, int64_t* out, int pitch);

/* "Go-faithful" CPU baseline (oracle/gofaithful.cpp): the same result through the reference's per-call
 * structure (maps, string switches, NodeScoreList), `threads` scheduling cycles in parallel. */
void orc_gofaithful_alloc_batch(const int64_t* const* cols, const char* const* res_names, int R, int N,
                                const int64_t* w, int mode, int P, const int64_t* pod_cpu_milli,
                                const int64_t* pod_mem_bytes, const uint64_t* feasible, int words, int64_t* out,
                                int pitch, int threads, double* compute_seconds);

/* ---- TargetLoadPacking (pkg/trimaran/targetloadpacking) ---- */
int64_t orc_tlp_score(double cpu_util_pct, int64_t cap_milli, int64_t missing_milli, uint8_t flags,
                      int64_t pod_cpu_milli, int64_t target_pct);
void orc_tlp_batch(const double* util, const int64_t* cap, const int64_t* missing, const uint8_t* flags, int N,
                   const int64_t* pod_cpu, int P, int64_t target, int64_t* out, int pitch);

/* ---- LoadVariationRiskBalancing (pkg/trimaran/loadvariationriskbalancing) ---- */
void orc_lvrb_mu_sigma(double used_avg, double used_std, double req, double capacity, double* mu_out,
                       double* sigma_out);
double orc_lvrb_compute_score(double used_avg, double used_std, double req, double capacity, double margin,
                              double sensitivity);
int64_t orc_lvrb_score(double cpu_avg, double cpu_std, double mem_avg, double mem_std, int64_t alloc_cpu_milli,
                       int64_t alloc_mem_bytes, uint8_t flags, int64_t req_cpu_milli, int64_t req_mem_bytes,
                       double margin, double sensitivity);
void orc_lvrb_batch(const double* cpu_avg, const double* cpu_std, const double* mem_avg, const double* mem_std,
                    const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags, int N,
                    const int64_t* req_cpu, const int64_t* req_mem, int P, double margin, double sens,
                    int64_t* out, int pitch);

/* test entry points of the two NUMA subtraction helpers (numaresources.go:145-182, :184-215) */
int orc_nrt_subtract_from_numa_list(int64_t* avail, const uint8_t* zmask, int Z, int R, const uint8_t* res_flags, int numa_id,
                                    int qos, uint8_t req_mask, const int64_t* req);
void orc_nrt_subtract_from_numas(int64_t* avail, const uint8_t* zmask, int Z, int R, uint32_t zone_mask, uint8_t req_mask,
                                 const int64_t* req);
/* OverReserve cache deduction of one assumed pod (cache/store.go:129-160); see nrt.c */
void orc_nrt_overreserve_deduct(int64_t* avail, const uint8_t* zmask, int Z, int R, uint8_t req_mask, const int64_t* req);

/* ---- Trimaran Peaks + LowRiskOverCommitment (trimaran2.c; see its header for the parity note) ---- */
double orc_go_exp(double x);
int64_t orc_peaks_score(double util_pct, int64_t cap_milli, uint8_t flags, double k1, double k2, int64_t pod_cpu_milli);
void orc_peaks_normalize(int64_t* scores, int n);
void orc_peaks_batch(const double* util, const int64_t* cap, const uint8_t* flags, const double* k1, const double* k2,
                     int N, const int64_t* pod_cpu, int P, const uint64_t* feasible, int words, int64_t* out, int pitch);
double orc_incbet(double a, double b, double x);
double orc_lowrisk_risk_load(int stats_ok, double util, double std, double capacity_f, int64_t capacity,
                             int64_t req_minus_pod, int64_t lim_minus_pod, int64_t window);
double orc_lowrisk_compute_risk(int stats_ok, double util, double std, double capacity_f, int64_t capacity,
                                int64_t node_req, int64_t node_lim, int64_t pod_req, int64_t pod_lim, int64_t window,
                                double weight);
int64_t orc_lowrisk_score(double cpu_avg, double cpu_std, double mem_avg, double mem_std, int64_t alloc_cpu_milli,
                          int64_t alloc_mem_bytes, uint8_t flags, int64_t node_req_cpu, int64_t node_req_mem,
                          int64_t node_lim_cpu, int64_t node_lim_mem, int64_t pod_req_cpu, int64_t pod_req_mem,
                          int64_t pod_lim_cpu, int64_t pod_lim_mem, int64_t window, double w_cpu, double w_mem);
void orc_lowrisk_batch(const double* cpu_avg, const double* cpu_std, const double* mem_avg, const double* mem_std,
                       const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags,
                       const int64_t* node_req_cpu, const int64_t* node_req_mem, const int64_t* node_lim_cpu,
                       const int64_t* node_lim_mem, int N, const int64_t* pod_req_cpu, const int64_t* pod_req_mem,
                       const int64_t* pod_lim_cpu, const int64_t* pod_lim_mem, int P, int64_t window, double w_cpu,
                       double w_mem, int64_t* out, int pitch);

/* ---- NetworkOverhead (pkg/networkaware/networkoverhead) ---- */
#define ORC_NETOH_MISSING INT64_MIN
typedef struct {
  int32_t host_node;
  uint16_t host_region, host_zone;
  int64_t max_network_cost;
} orc_netoh_dep;
typedef struct {
  int n_names;
  const int64_t* zone_cost;   /* [K][K] */
  const int64_t* region_cost; /* [K][K] */
} orc_netoh_topology;
void orc_netoh_node(const orc_netoh_topology* t, int node_global, int region, int zone, const orc_netoh_dep* deps,
                    int n_deps, int64_t* satisfied, int64_t* violated, int64_t* cost_out);
void orc_netoh_normalize(int64_t* scores, int n);
void orc_netoh_batch(const orc_netoh_topology* t, const uint16_t* region_id, const uint16_t* zone_id, int N,
                     int node_offset, const uint8_t* score_equally, const int32_t* dep_offset,
                     const orc_netoh_dep* deps, int P, const uint64_t* feasible, int words, int64_t* out_scores,
                     uint64_t* out_feasible, uint8_t* out_reasons, int pitch);

/* ---- NodeResourceTopologyMatch (pkg/noderesourcetopology) ---- */
#define ORC_QOS_GUARANTEED 0
#define ORC_QOS_BURSTABLE 1
#define ORC_QOS_BEST_EFFORT 2
#define ORC_NRT_NODE_HAS_NRT 1u
#define ORC_NRT_NODE_FRESH 2u
#define ORC_NRT_NODE_SINGLE_NUMA 4u
#define ORC_NRT_NODE_SCOPE_POD 8u
#define ORC_NRT_NODE_UNSUPPORTED 16u
#define ORC_NRT_RES_AFFINE 1u
#define ORC_NRT_RES_HOST_LEVEL 2u
#define ORC_NRT_POD_FILTER_BYPASS 1u
#define ORC_NRT_POD_UNSUPPORTED 2u
#define ORC_CONT_APP 0
#define ORC_CONT_INIT 1
#define ORC_CONT_SIDECAR 2
#define ORC_NRT_MOST_ALLOCATED 0
#define ORC_NRT_BALANCED_ALLOCATION 1
#define ORC_NRT_LEAST_ALLOCATED 2
#define ORC_NRT_LEAST_NUMA_NODES 3
#define ORC_REASON_OK 0
#define ORC_REASON_NRT_INVALID_TOPOLOGY 1
#define ORC_REASON_NRT_ALIGN_POD 2
#define ORC_REASON_NRT_ALIGN_CONTAINER 3
#define ORC_REASON_NRT_ALIGN_INIT 4
#define ORC_REASON_NRT_ALIGN_SIDECAR 5
#define ORC_REASON_NRT_ACCOUNTING 6
#define ORC_REASON_UPSTREAM 8
#define ORC_REASON_UNSUPPORTED 9
typedef struct {
  uint8_t flags;
  uint16_t max_numa;
  uint8_t n_zones;
  uint8_t node_res_mask;
  uint8_t zone_res_mask[8];
  int64_t avail[8][8]; /* [zone][resource slot], milli-units */
  int32_t cost[8][8];  /* -1 = missing */
} orc_nrt_node;
typedef struct {
  uint8_t qos, flags, n_init, n_app;
  uint8_t cont_kind[8];
  uint8_t req_mask[9]; /* slot 8 = pod effective request */
  int64_t req[9][8];
} orc_nrt_pod;
/* returns an ORC_REASON_* code (0 = pass) */
int orc_nrt_filter(const orc_nrt_node* nd, const orc_nrt_pod* pod, const uint8_t* res_flags, int R);
int64_t orc_nrt_score(const orc_nrt_node* nd, const orc_nrt_pod* pod, const uint8_t* res_flags, int R, int strategy,
                      const int64_t* weights);
int64_t orc_nrt_normalize_least_numa(int count, int is_min, int max_numa);
int orc_nrt_only_non_numa(const uint8_t* zone_res_mask, int n_zones, uint8_t req_mask, int R);
int orc_nrt_numa_nodes_required(const orc_nrt_node* nd, const uint8_t* res_flags, int R, int qos, uint8_t req_mask,
                                const int64_t* req, uint32_t* mask_out, int* is_min);
float orc_nrt_min_avg_distance(const int32_t* cost, int n, const int* combos, int n_combos, int k);
typedef struct {
  int32_t n_zones, n_res;
  const uint8_t* res_flags;     /* [R] */
  const uint8_t* node_flags;    /* [N] */
  const uint16_t* max_numa;     /* [N] */
  const uint8_t* n_zones_node;  /* [N] */
  const uint8_t* node_res_mask; /* [N] */
  const uint8_t* zone_res_mask; /* [Z][N] */
  const int64_t* avail;         /* [Z][R][N] */
  const int32_t* cost;          /* [Z][Z][N] or NULL */
} orc_nrt_nodes_soa;
typedef struct {
  const uint8_t* qos;
  const uint8_t* flags;
  const uint8_t* n_init;
  const uint8_t* n_app;
  const uint8_t* cont_kind; /* [P][8] */
  const uint8_t* req_mask;  /* [P][9] */
  const int64_t* req;       /* [P][9][R] */
} orc_nrt_pods_soa;
void orc_nrt_batch(const orc_nrt_nodes_soa* ns, int N, const orc_nrt_pods_soa* ps, int P, int strategy,
                   const int64_t* weights, const uint64_t* feasible, int words, int64_t* out_scores,
                   uint64_t* out_feasible, uint8_t* out_reasons, int pitch);

#ifdef __cplusplus
}
#endif
#endif

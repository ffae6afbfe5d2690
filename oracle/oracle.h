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

#ifdef __cplusplus
}
#endif
#endif

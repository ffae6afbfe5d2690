/* Oracle: NodeResourceTopologyMatch Filter + Score on the dense encoding.
 * TEST INFRASTRUCTURE — see oracle.h.
 *
 * Encoding (what the host layer flattens an NRT object + TopologyManager config into; the same
 * columns the CUDA engine reads):
 *   zones 0..nz-1 are the NUMA cells with NUMA id == zone index (createNUMANodeList,
 *   pluginhelpers.go:105-134; anything else is flagged UNSUPPORTED by the host and falls back);
 *   avail[z][r]   zone "Available" in exact milli-units (pluginhelpers.go:155-161);
 *   zone_res_mask bit r: the zone lists resource r; node_res_mask bit r: reported at node level
 *   (util.ResourceList(GetAllocatable()), filter.go:97);
 *   res_flags[r]: bit0 isNUMAAffineResource, bit1 isHostLevelResource (numaresources.go:105-135);
 *   cost[z1][z2]: Costs of zone z1 towards NUMA id z2, -1 = missing (pluginhelpers.go:136-153).
 * Quantities: Cmp / IsZero / Sub on exact milli values; Value() = ceil(milli/1000) (apimachinery).
 */
#include <math.h>
#include <string.h>

#include "oracle.h"

#define Z_MAX 8
#define R_MAX 8
#define C_MAX 8

typedef struct {
  int nz;
  int64_t avail[Z_MAX][R_MAX];
  uint8_t zmask[Z_MAX];
} zones_t;

static int64_t qty_value(int64_t milli) { /* Quantity.Value(): rounded up */
  if (milli >= 0) return (milli + 999) / 1000;
  return -((-milli) / 1000);
}

static int64_t f2i(double x) { /* Go int64(float64) on amd64 */
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}

/* isResourceSetSuitable: numaresources.go:137-142 */
static int suitable(int qos, uint8_t rflags, int64_t qty, int64_t numa_qty) {
  if (qos != ORC_QOS_GUARANTEED && (rflags & ORC_NRT_RES_AFFINE)) return 1;
  return numa_qty >= qty;
}

/* resourcesAvailableInAnyNUMANodes: filter.go:90-160.  Returns match; *numa_id = lowest NUMA id of
 * the surviving bitmask (0 when no resource constrained it, :154). */
static int available_in_any(const orc_nrt_node* nd, const zones_t* zs, const uint8_t* res_flags, int R, int qos,
                            uint8_t req_mask, const int64_t* req, int* numa_id) {
  uint64_t bitmask = ~0ull; /* bitmask.Fill(), :92-95 */
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1)) continue;
    if (req[r] == 0) continue;                      /* :101-105 */
    if (!((nd->node_res_mask >> r) & 1)) return 0;   /* :107-113 */
    int has_affinity = 0;
    uint64_t res_bits = 0;
    for (int z = 0; z < zs->nz; ++z) {
      if (!((zs->zmask[z] >> r) & 1)) continue;      /* :121-125 */
      has_affinity = 1;
      if (!suitable(qos, res_flags[r], req[r], zs->avail[z][r])) continue; /* :128-131 */
      res_bits |= 1ull << z;                          /* :133 */
    }
    if (!has_affinity && (res_flags[r] & ORC_NRT_RES_HOST_LEVEL)) continue; /* :139-142 */
    bitmask &= res_bits;
    if (bitmask == 0) return 0;                       /* :144-148 */
  }
  *numa_id = __builtin_ctzll(bitmask);                /* :154 */
  return 1;
}

/* subtractResourcesFromNUMANodeList: numaresources.go:145-182.  Returns 0 on "inconsistent
 * resource accounting". */
static int subtract_from_numa(zones_t* zs, const uint8_t* res_flags, int R, int numa_id, int qos, uint8_t req_mask,
                              const int64_t* req) {
  for (int z = 0; z < zs->nz; ++z) {
    if (z != numa_id) continue; /* node.NUMAID != numaID */
    for (int r = 0; r < R; ++r) {
      if (!((req_mask >> r) & 1)) continue;
      if (qos != ORC_QOS_GUARANTEED && (res_flags[r] & ORC_NRT_RES_AFFINE)) continue; /* :156-160 */
      if (req[r] == 0) continue;                                                       /* :161-164 */
      if (!((zs->zmask[z] >> r) & 1)) continue;                                        /* :165-169 */
      int64_t q = zs->avail[z][r] - req[r];
      if (q < 0) return 0;                                                             /* :172-175 */
      zs->avail[z][r] = q;
    }
  }
  return 1;
}

static void load_zones(const orc_nrt_node* nd, zones_t* zs) {
  zs->nz = nd->n_zones;
  for (int z = 0; z < nd->n_zones; ++z) {
    zs->zmask[z] = nd->zone_res_mask[z];
    for (int r = 0; r < R_MAX; ++r) zs->avail[z][r] = nd->avail[z][r];
  }
}

/* TopologyMatch.Filter: filter.go:176-225 with the two handlers :39-78, :162-173. */
int orc_nrt_filter(const orc_nrt_node* nd, const orc_nrt_pod* pod, const uint8_t* res_flags, int R) {
  if (pod->flags & ORC_NRT_POD_FILTER_BYPASS) return ORC_REASON_OK;     /* :180-183 */
  if (!(nd->flags & ORC_NRT_NODE_FRESH)) return ORC_REASON_NRT_INVALID_TOPOLOGY; /* :194-197 */
  if (!(nd->flags & ORC_NRT_NODE_HAS_NRT)) return ORC_REASON_OK;         /* :198-200 */
  if (!(nd->flags & ORC_NRT_NODE_SINGLE_NUMA)) return ORC_REASON_OK;     /* :206-209, :228-230 */
  /* the dense encoding's own escape hatch: only where the reference would start reading zones / containers (:211) */
  if ((nd->flags & ORC_NRT_NODE_UNSUPPORTED) || (pod->flags & ORC_NRT_POD_UNSUPPORTED)) return ORC_REASON_UNSUPPORTED;
  zones_t zs;
  load_zones(nd, &zs);
  int numa_id = 0;
  if (nd->flags & ORC_NRT_NODE_SCOPE_POD) { /* singleNUMAPodLevelHandler */
    if (!available_in_any(nd, &zs, res_flags, R, pod->qos, pod->req_mask[C_MAX], pod->req[C_MAX], &numa_id))
      return ORC_REASON_NRT_ALIGN_POD;
    return ORC_REASON_OK;
  }
  /* singleNUMAContainerLevelHandler */
  for (int c = 0; c < pod->n_init; ++c) { /* :43-55: init containers, no subtraction */
    if (!available_in_any(nd, &zs, res_flags, R, pod->qos, pod->req_mask[c], pod->req[c], &numa_id))
      return pod->cont_kind[c] == ORC_CONT_SIDECAR ? ORC_REASON_NRT_ALIGN_SIDECAR : ORC_REASON_NRT_ALIGN_INIT;
  }
  for (int c = pod->n_init; c < pod->n_init + pod->n_app; ++c) { /* :57-76 */
    if (!available_in_any(nd, &zs, res_flags, R, pod->qos, pod->req_mask[c], pod->req[c], &numa_id))
      return ORC_REASON_NRT_ALIGN_CONTAINER;
    if (!subtract_from_numa(&zs, res_flags, R, numa_id, pod->qos, pod->req_mask[c], pod->req[c]))
      return ORC_REASON_NRT_ACCOUNTING;
  }
  return ORC_REASON_OK;
}

/* ---- scoring strategies ---- */
static int64_t weight_of(const int64_t* w, int r) { return w[r] < 1 ? 1 : w[r]; } /* score.go:49-60 */

/* leastAllocatedScoreStrategy / mostAllocatedScoreStrategy: least_allocated.go:25-55, most_allocated.go:25-54 */
static int64_t strat_least_most(int most, uint8_t req_mask, const int64_t* req, const zones_t* zs, int z,
                                const int64_t* w, int R) {
  int64_t node_score = 0, weight_sum = 0;
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1)) continue;
    int64_t cap = ((zs->zmask[z] >> r) & 1) ? zs->avail[z][r] : 0; /* missing key -> zero Quantity */
    int64_t s;
    if (cap == 0) s = 0;            /* CmpInt64(0) == 0 */
    else if (req[r] > cap) s = 0;   /* requested.Cmp(numaCapacity) > 0 */
    else if (most) s = orc_go_div(orc_wrap_mul(qty_value(req[r]), 100), qty_value(cap));
    else s = orc_go_div(orc_wrap_mul(qty_value(cap) - qty_value(req[r]), 100), qty_value(cap));
    int64_t wt = weight_of(w, r);
    node_score = orc_wrap_add(node_score, orc_wrap_mul(s, wt));
    weight_sum = orc_wrap_add(weight_sum, wt);
  }
  if (weight_sum == 0) return 0; /* empty request map panics in Go (least_allocated.go:38); unreachable for Guaranteed pods */
  return orc_go_div(node_score, weight_sum);
}

/* balancedAllocationScoreStrategy: balanced_allocation.go:27-54 with gonum stat.Variance
 * (unweighted: mean = sum/n; corrected two-pass; / (n-1)) [gonum v0.12.0, not in tree].
 * Summation order = resource-slot order (the reference iterates a Go map: order-dependent for
 * three or more resources, deterministic for the default cpu+memory pair). */
static int64_t strat_balanced(uint8_t req_mask, const int64_t* req, const zones_t* zs, int z, int R) {
  double fr[R_MAX];
  int n = 0;
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1)) continue;
    int64_t cap = ((zs->zmask[z] >> r) & 1) ? zs->avail[z][r] : 0;
    double f = qty_value(cap) == 0 ? 1.0 : (double)qty_value(req[r]) / (double)qty_value(cap);
    if (f > 1) return 0;
    fr[n++] = f;
  }
  double sum = 0;
  for (int i = 0; i < n; ++i) sum += fr[i];
  double mean = sum / (double)n;
  double ss = 0, comp = 0;
  for (int i = 0; i < n; ++i) {
    double d = fr[i] - mean;
    ss += d * d;
    comp += d;
  }
  double variance = (ss - comp * comp / (double)n) / ((double)n - 1);
  return f2i((1 - variance) * 100.0);
}

static int64_t strategy_score(int strategy, uint8_t req_mask, const int64_t* req, const zones_t* zs, int z,
                              const int64_t* w, int R) {
  if (strategy == ORC_NRT_MOST_ALLOCATED) return strat_least_most(1, req_mask, req, zs, z, w, R);
  if (strategy == ORC_NRT_LEAST_ALLOCATED) return strat_least_most(0, req_mask, req, zs, z, w, R);
  return strat_balanced(req_mask, req, zs, z, R);
}

/* scoreForEachNUMANode: score.go:110-124 */
static int64_t score_each_numa(int strategy, uint8_t req_mask, const int64_t* req, const zones_t* zs,
                               const int64_t* w, int R) {
  int64_t min_score = 0;
  for (int z = 0; z < zs->nz; ++z) {
    int64_t s = strategy_score(strategy, req_mask, req, zs, z, w, R);
    if (min_score == 0 || (s != 0 && s < min_score)) min_score = s;
  }
  return min_score;
}

/* ---- LeastNUMANodes: least_numa.go ---- */
static int only_non_numa(const zones_t* zs, uint8_t req_mask, int R) { /* pluginhelpers.go:163-173 */
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1)) continue;
    for (int z = 0; z < zs->nz; ++z)
      if ((zs->zmask[z] >> r) & 1) return 0;
  }
  return 1;
}

static float avg_distance(const orc_nrt_node* nd, const int* combo, int k) { /* :116-138 */
  if (k == 0) return 255.0f;
  int accu = 0;
  for (int i = 0; i < k; ++i)
    for (int j = 0; j < k; ++j) {
      int c = nd->cost[combo[i]][combo[j]];
      if (c < 0) c = 255; /* missing -> maxDistanceValue */
      accu += c;
    }
  return (float)accu / (float)(k * k);
}

static int next_combination(int* idx, int n, int k) { /* lexicographic, as gonum combin.Combinations */
  int i = k - 1;
  while (i >= 0 && idx[i] == n - k + i) --i;
  if (i < 0) return 0;
  ++idx[i];
  for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
  return 1;
}

/* numaNodesRequired :159-174 + findSuitableCombination :179-208.  Returns the number of NUMA
 * nodes (0 = cannot fit), the chosen zone mask and the min-distance flag. */
static int numa_nodes_required(const orc_nrt_node* nd, const zones_t* zs, const uint8_t* res_flags, int R, int qos,
                               uint8_t req_mask, const int64_t* req, uint32_t* mask_out, int* is_min) {
  int n = zs->nz;
  for (int k = 1; k <= n; ++k) {
    int idx[Z_MAX];
    for (int i = 0; i < k; ++i) idx[i] = i;
    float min_avg = 255.0f; /* minAvgDistanceInCombinations :102-114 */
    do {
      float d = avg_distance(nd, idx, k);
      if (d < min_avg) min_avg = d;
    } while (next_combination(idx, n, k));
    for (int i = 0; i < k; ++i) idx[i] = i;
    int have = 0;
    uint32_t best_mask = 0;
    float min_dist = 256.0f;
    do {
      int valid = 1; /* isValidCombineResources :224-233 */
      for (int i = 0; i < k && valid; ++i)
        for (int r = 0; r < R; ++r)
          if (((req_mask >> r) & 1) && !((zs->zmask[idx[i]] >> r) & 1)) { valid = 0; break; }
      if (!valid) continue;
      int fit = 1; /* combineResources + checkResourcesFit :140-157, :210-222 */
      for (int r = 0; r < R && fit; ++r) {
        if (!((req_mask >> r) & 1) || req[r] == 0) continue;
        int64_t sum = 0;
        for (int i = 0; i < k; ++i) sum += zs->avail[idx[i]][r];
        if (!suitable(qos, res_flags[r], req[r], sum)) fit = 0;
      }
      if (!fit) continue;
      float dist = avg_distance(nd, idx, k);
      uint32_t m = 0;
      for (int i = 0; i < k; ++i) m |= 1u << idx[i];
      if (dist == min_avg) { /* :195-198 */
        *mask_out = m;
        *is_min = 1;
        return k;
      }
      if (dist < min_dist) { /* :200-203 */
        min_dist = dist;
        best_mask = m;
        have = 1;
      }
    } while (next_combination(idx, n, k));
    if (have) {
      *mask_out = best_mask;
      *is_min = 0;
      return k;
    }
  }
  return 0;
}

static int64_t normalize_least_numa(int count, int is_min, int max_numa) { /* :91-100 */
  int64_t numa_node_score = 100 / (int64_t)max_numa;
  int64_t score = 100 - (int64_t)count * numa_node_score;
  if (is_min) return score + numa_node_score / 2;
  return score;
}

/* exported for the golden tests of least_numa_test.go (TestNormalizeScore :706-756, TestMinDistance :758-912) */
int64_t orc_nrt_normalize_least_numa(int count, int is_min, int max_numa) {
  return normalize_least_numa(count, is_min, max_numa);
}
float orc_nrt_min_avg_distance(const int32_t* cost /* [n][n], -1 = missing */, int n, const int* combos, int n_combos,
                               int k) { /* minAvgDistanceInCombinations :102-114 */
  orc_nrt_node nd;
  memset(&nd, 0, sizeof(nd));
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) nd.cost[i][j] = cost[i * n + j];
  float min_d = 255.0f;
  for (int c = 0; c < n_combos; ++c) {
    float d = avg_distance(&nd, combos + c * k, k);
    if (d < min_d) min_d = d;
  }
  return min_d;
}

/* subtractFromNUMAs: numaresources.go:184-215 (greedy across the chosen zones in bit order) */
static void subtract_from_numas(zones_t* zs, uint8_t req_mask, const int64_t* req, uint32_t mask, int R) {
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1)) continue;
    int64_t q = req[r];
    for (int z = 0; z < zs->nz; ++z) {
      if (!((mask >> z) & 1)) continue;
      if (q == 0) break;
      if (!((zs->zmask[z] >> r) & 1)) continue;
      int64_t av = zs->avail[z][r];
      if (q >= av) { /* Cmp 0 or 1 */
        q -= av;
        zs->avail[z][r] = 0;
      } else {
        zs->avail[z][r] = av - q;
        q = 0;
      }
    }
  }
}

/* TopologyMatch.Score: score.go:62-102 and the handlers :142-191, least_numa.go:35-89. */
int64_t orc_nrt_score(const orc_nrt_node* nd, const orc_nrt_pod* pod, const uint8_t* res_flags, int R, int strategy,
                      const int64_t* weights) {
  if (pod->qos != ORC_QOS_GUARANTEED) return 100;           /* :72-75 */
  if (!(nd->flags & ORC_NRT_NODE_FRESH)) return 0;           /* :79-82 */
  if (!(nd->flags & ORC_NRT_NODE_HAS_NRT)) return 0;         /* :83-86 */
  if ((nd->flags & ORC_NRT_NODE_UNSUPPORTED) || (pod->flags & ORC_NRT_POD_UNSUPPORTED)) return 0;
  zones_t zs;
  load_zones(nd, &zs);
  const int scope_pod = (nd->flags & ORC_NRT_NODE_SCOPE_POD) != 0;
  const int nc = pod->n_init + pod->n_app;
  if (strategy == ORC_NRT_LEAST_NUMA_NODES) { /* :168-176: no policy check for LeastNUMANodes */
    uint32_t mask;
    int is_min;
    if (scope_pod) { /* leastNUMAPodScopeScore :73-89 */
      if (only_non_numa(&zs, pod->req_mask[C_MAX], R)) return 100;
      int k = numa_nodes_required(nd, &zs, res_flags, R, pod->qos, pod->req_mask[C_MAX], pod->req[C_MAX], &mask, &is_min);
      if (k == 0) return 0;
      return normalize_least_numa(k, is_min, nd->max_numa);
    }
    int max_count = 0, all_min = 1; /* leastNUMAContainerScopeScore :35-71 */
    for (int c = 0; c < nc; ++c) {
      if (only_non_numa(&zs, pod->req_mask[c], R)) continue;
      int k = numa_nodes_required(nd, &zs, res_flags, R, pod->qos, pod->req_mask[c], pod->req[c], &mask, &is_min);
      if (k == 0) return 0;
      if (!is_min) all_min = 0;
      if (k > max_count) max_count = k;
      subtract_from_numas(&zs, pod->req_mask[c], pod->req[c], mask, R);
    }
    if (max_count == 0) return 100;
    return normalize_least_numa(max_count, all_min, nd->max_numa);
  }
  if (!(nd->flags & ORC_NRT_NODE_SINGLE_NUMA)) return 0; /* :177-179 */
  if (scope_pod) /* podScopeScore :142-150 */
    return score_each_numa(strategy, pod->req_mask[C_MAX], pod->req[C_MAX], &zs, weights, R);
  /* containerScopeScore :152-165: mean over init + app containers, no subtraction */
  double sum = 0;
  for (int c = 0; c < nc; ++c) sum += (double)score_each_numa(strategy, pod->req_mask[c], pod->req[c], &zs, weights, R);
  return f2i(sum / (double)nc); /* stat.Mean = Sum/len */
}

/* Batch driver over the SoA columns (same layout the engine takes). */
void orc_nrt_batch(const orc_nrt_nodes_soa* ns, int N, const orc_nrt_pods_soa* ps, int P, int strategy,
                   const int64_t* weights, const uint64_t* feasible, int words, int64_t* out_scores,
                   uint64_t* out_feasible, uint8_t* out_reasons, int pitch) {
  const int Z = ns->n_zones, R = ns->n_res, owords = pitch / 64;
  for (int p = 0; p < P; ++p) {
    orc_nrt_pod pod;
    memset(&pod, 0, sizeof(pod));
    pod.qos = ps->qos[p];
    pod.flags = ps->flags[p];
    pod.n_init = ps->n_init[p];
    pod.n_app = ps->n_app[p];
    for (int c = 0; c < C_MAX; ++c) pod.cont_kind[c] = ps->cont_kind[(size_t)p * C_MAX + c];
    for (int c = 0; c <= C_MAX; ++c) {
      pod.req_mask[c] = ps->req_mask[(size_t)p * (C_MAX + 1) + c];
      for (int r = 0; r < R; ++r) pod.req[c][r] = ps->req[((size_t)p * (C_MAX + 1) + c) * R + r];
    }
    for (int w = 0; w < owords; ++w) out_feasible[(size_t)p * owords + w] = 0;
    for (int n = 0; n < pitch; ++n) {
      out_scores[(size_t)p * pitch + n] = 0;
      out_reasons[(size_t)p * pitch + n] = 0;
    }
    for (int n = 0; n < N; ++n) {
      orc_nrt_node nd;
      memset(&nd, 0, sizeof(nd));
      nd.flags = ns->node_flags[n];
      nd.max_numa = ns->max_numa[n];
      nd.n_zones = ns->n_zones_node[n];
      nd.node_res_mask = ns->node_res_mask[n];
      for (int z = 0; z < Z; ++z) {
        nd.zone_res_mask[z] = ns->zone_res_mask[(size_t)z * N + n];
        for (int r = 0; r < R; ++r) nd.avail[z][r] = ns->avail[((size_t)z * R + r) * N + n];
        for (int z2 = 0; z2 < Z; ++z2) nd.cost[z][z2] = ns->cost ? ns->cost[((size_t)z * Z + z2) * N + n] : -1;
      }
      int reason = orc_nrt_filter(&nd, &pod, ns->res_flags, R);
      int up = !feasible || ((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull);
      out_reasons[(size_t)p * pitch + n] = (uint8_t)(reason != 0 ? reason : (up ? 0 : ORC_REASON_UPSTREAM));
      if (reason != 0 || !up) continue;
      out_feasible[(size_t)p * owords + (n >> 6)] |= 1ull << (n & 63);
      out_scores[(size_t)p * pitch + n] = orc_nrt_score(&nd, &pod, ns->res_flags, R, strategy, weights);
    }
  }
}

/* resourceStore.UpdateNRT: pkg/noderesourcetopology/cache/store.go:129-160 -- the OverReserve cache's pessimistic
 * deduction of ONE assumed pod's effective request from EVERY zone that lists the resource (GetCachedNRTCopy,
 * overreserve.go:101-127, applies it once per assumed pod of the node).  avail [Z][R] in place; zmask[z] bit r = zone
 * lists r; req_mask bit r = r is a key of the pod's request map. */
void orc_nrt_overreserve_deduct(int64_t* avail, const uint8_t* zmask, int Z, int R, uint8_t req_mask, const int64_t* req) {
  for (int z = 0; z < Z; ++z)
    for (int r = 0; r < R; ++r) {
      if (!((zmask[z] >> r) & 1)) continue;   /* the zone does not report the resource */
      if (!((req_mask >> r) & 1)) continue;    /* :141-147 the pod does not ask for it */
      int64_t* a = &avail[z * R + r];
      if (*a < req[r]) {                       /* :148-155 "cannot decrement resource": zeroed */
        *a = 0;
        continue;
      }
      *a -= req[r];                            /* :157 */
    }
}

/* Test entry points for the two subtraction helpers, so that their reference vectors
 * (numaresources_test.go:117-373 and :375-462) pin them directly and not only through Filter / Score.
 * avail [Z][R_MAX-wide rows of R used], zmask[z] bit r = zone lists r. */
int orc_nrt_subtract_from_numa_list(int64_t* avail, const uint8_t* zmask, int Z, int R, const uint8_t* res_flags, int numa_id,
                                    int qos, uint8_t req_mask, const int64_t* req) {
  zones_t zs;
  zs.nz = Z;
  for (int z = 0; z < Z; ++z) {
    zs.zmask[z] = zmask[z];
    for (int r = 0; r < R; ++r) zs.avail[z][r] = avail[z * R + r];
  }
  const int ok = subtract_from_numa(&zs, res_flags, R, numa_id, qos, req_mask, req);
  for (int z = 0; z < Z; ++z)
    for (int r = 0; r < R; ++r) avail[z * R + r] = zs.avail[z][r];
  return ok;
}

void orc_nrt_subtract_from_numas(int64_t* avail, const uint8_t* zmask, int Z, int R, uint32_t zone_mask, uint8_t req_mask,
                                 const int64_t* req) {
  zones_t zs;
  zs.nz = Z;
  for (int z = 0; z < Z; ++z) {
    zs.zmask[z] = zmask[z];
    for (int r = 0; r < R; ++r) zs.avail[z][r] = avail[z * R + r];
  }
  subtract_from_numas(&zs, req_mask, req, zone_mask, R);
  for (int z = 0; z < Z; ++z)
    for (int r = 0; r < R; ++r) avail[z * R + r] = zs.avail[z][r];
}

/* Test entry point for numaNodesRequired (least_numa.go:159-208): the zone bitmask that can host the request with the
 * fewest zones, and whether it is a minimum-distance combination.  Returns the zone count, 0 = cannot fit. */
/* onlyNonNUMAResources (pluginhelpers.go:163-173) on zone resource masks: 1 iff no zone lists any requested resource */
int orc_nrt_only_non_numa(const uint8_t* zone_res_mask, int n_zones, uint8_t req_mask, int R) {
  zones_t zs;
  memset(&zs, 0, sizeof(zs));
  zs.nz = n_zones;
  for (int z = 0; z < n_zones && z < Z_MAX; ++z) zs.zmask[z] = zone_res_mask[z];
  return only_non_numa(&zs, req_mask, R);
}

int orc_nrt_numa_nodes_required(const orc_nrt_node* nd, const uint8_t* res_flags, int R, int qos, uint8_t req_mask,
                                const int64_t* req, uint32_t* mask_out, int* is_min) {
  zones_t zs;
  load_zones(nd, &zs);
  *mask_out = 0;
  *is_min = 0;
  return numa_nodes_required(nd, &zs, res_flags, R, qos, req_mask, req, mask_out, is_min);
}

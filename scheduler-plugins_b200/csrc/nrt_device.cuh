// Device side of NodeResourceTopologyMatch for ONE (pod, node) pair: Filter (filter.go:176-225) and Score
// (score.go:62-191) on the dense encoding, the node's zones x resources block in registers.  Shared by the batched
// direct kernel (nrt.cu) and the fused single-cycle kernel (cycle.cu); moved here verbatim from nrt.cu.
#pragma once
#include "engine.h"

namespace b200s {
namespace nrtdev {

constexpr int C_MAX = B200S_NRT_MAX_CONT;

// Combination tables for LeastNUMANodes: for n zones, all non-empty subsets ordered by size then
// lexicographically by index tuple (gonum combin.Combinations order, least_numa.go:161).
static __constant__ uint8_t c_combo_mask[8][255];
static __constant__ uint8_t c_combo_off[8][10];  // [n-1][k] = first index of size-k subsets; [n-1][n+1] = end
template <int R>
struct PodS {  // one pod of the tile, in shared memory
  int64_t req[C_MAX + 1][R];
  int64_t reqv[C_MAX + 1][R];  // Quantity.Value() of req (ceil to whole units), precomputed once per tile
  // Filter operands (see nrt_filter): eff = threshold a zone's availability must reach (INT64_MIN: resource does
  // not constrain), sub = what an app container takes from its zone, need = requested with a non-zero quantity
  int64_t eff[C_MAX + 1][R];
  int64_t sub[C_MAX + 1][R];
  uint8_t need[C_MAX + 1];
  uint8_t req_mask[C_MAX + 1];
  uint8_t kind[C_MAX];
  uint8_t qos, flags, n_init, n_app;
};

template <int Z, int R>
struct Zones {
  int64_t avail[Z][R];
  uint32_t zmask[Z];
  int nz;
};

struct NrtCfg {
  int strategy;
  int64_t w[B200S_NRT_MAX_RES];
  uint8_t res_flags[B200S_NRT_MAX_RES];
};

__device__ __forceinline__ int64_t qty_value(int64_t milli) { return (milli + 999) / 1000; }  // Quantity.Value(), >= 0
__device__ __forceinline__ int64_t f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
// floor(a*100/cv) for 0 <= a <= cv (quotient in 0..100): fp32 estimate + exact integer fix-up instead of
// a 64-bit division; identical to Go's (a*100)/cv.  Absurdly large capacities keep the wrapping Go path.
// generic path of div100 (quantities beyond 2^52 or out-of-range numerators): kept out of line so that the 16
// unrolled (zone, resource) call sites stay small
static __device__ __noinline__ int64_t div100_slow(int64_t a, int64_t cv) { return go_div(wrap_mul(a, 100), cv); }

__device__ __forceinline__ int64_t div100(int64_t a, int64_t cv) {
  if (cv >= (1ll << 52) || a < 0 || a > cv) return div100_slow(a, cv);
  const int64_t num = a * 100;
  int64_t q = (int64_t)__float2int_rd(__fdividef(__ll2float_rn(num), __ll2float_rn(cv)));
  int64_t rem = num - q * cv;
  // the quotient is <= 100 and the fp32 estimate carries a relative error below 2^-21: floor() is off by at most
  // one either way, so one predicated step each way replaces the correction loops (straight-line code that the
  // unrolled resource loop can interleave); anything else would be a logic error and takes the exact path
  const bool under = rem < 0;
  q -= under ? 1 : 0;
  rem += under ? cv : 0;
  const bool over = rem >= cv;
  q += over ? 1 : 0;
  rem -= over ? cv : 0;
  if (rem < 0 || rem >= cv) return div100_slow(a, cv);
  return q;
}

__device__ __forceinline__ bool suitable(int qos, uint32_t rflags, int64_t qty, int64_t numa_qty) {
  if (qos != B200S_QOS_GUARANTEED && (rflags & B200S_NRT_RES_AFFINE)) return true;
  return numa_qty >= qty;
}

// TopologyMatch.Filter -> reason code (filter.go:176-225, handlers :39-78 / :162-173).
//
// resourcesAvailableInAnyNUMANodes (:90-160) as a branch-free dominance test.  The node side is pre-encoded once
// per tile (kernel prologue): avail[z][r] = Available where the zone lists r; INT64_MIN where it does not but
// another zone does (never suitable, :121-125); INT64_MAX where NO zone lists r and r is host-level (no constraint,
// :139-142); INT64_MIN where no zone lists a NUMA-bound resource (bitmask becomes empty, :144-148).  The pod side
// is pre-encoded per tile in shared memory: eff[r] = INT64_MIN if r is not requested or zero (:101-105: passes
// everywhere), INT64_MIN+1 if the pod is not Guaranteed and r is NUMA-affine (isResourceSetSuitable :137-142:
// every LISTING zone is suitable), else the quantity.  Zone z survives iff avail[z][r] >= eff[r] for all r; the chosen NUMA id is the lowest
// surviving zone (an unconstrained request keeps the all-ones mask -> id 0, :154).
template <int Z, int R>
__device__ int nrt_filter(const Zones<Z, R>& node_zs, uint32_t nflags, uint32_t node_res_mask, const NrtCfg& cfg,
                          const PodS<R>& pod) {
  if (pod.flags & B200S_NRT_POD_FILTER_BYPASS) return B200S_REASON_OK;
  if (!(nflags & B200S_NRT_NODE_FRESH)) return B200S_REASON_NRT_INVALID_TOPOLOGY;
  if (!(nflags & B200S_NRT_NODE_HAS_NRT)) return B200S_REASON_OK;
  if (!(nflags & B200S_NRT_NODE_SINGLE_NUMA)) return B200S_REASON_OK;
  // shapes outside the dense encoding only matter where the reference would look at the zones / containers: after
  // its own gates (filter.go:194-209), so a stale or policy-less node answers as the reference does
  if ((nflags & B200S_NRT_NODE_UNSUPPORTED) || (pod.flags & B200S_NRT_POD_UNSUPPORTED)) return B200S_REASON_UNSUPPORTED;
  // One loop for both scopes: pod scope = a single step on the pod-effective request (slot C_MAX,
  // singleNUMAPodLevelHandler :162-173); container scope = init containers without subtraction, then app
  // containers with it (:39-78).
  const bool scope_pod = nflags & B200S_NRT_NODE_SCOPE_POD;
  const int n_init = pod.n_init, steps = scope_pod ? 1 : n_init + pod.n_app;
  Zones<Z, R> zs = node_zs;  // working copy: app containers subtract what they take
  for (int s = 0; s < steps; ++s) {
    const int c = scope_pod ? C_MAX : s;
    uint32_t ok = 0;
    if (!(pod.need[c] & ~node_res_mask)) {  // every requested resource is reported at node level (:107-113)
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        bool fits = true;
#pragma unroll
        for (int r = 0; r < R; ++r) fits &= zs.avail[z][r] >= pod.eff[c][r];
        ok |= (fits ? 1u : 0u) << z;
      }
    }
    if (ok == 0) {
      if (scope_pod) return B200S_REASON_NRT_ALIGN_POD;
      if (s >= n_init) return B200S_REASON_NRT_ALIGN_CONTAINER;
      return pod.kind[c] == B200S_CONT_SIDECAR ? B200S_REASON_NRT_ALIGN_SIDECAR : B200S_REASON_NRT_ALIGN_INIT;
    }
    if (scope_pod || s < n_init) continue;
    // subtractResourcesFromNUMANodeList (numaresources.go:145-182) on the zone with the lowest surviving id
    const int numa_id = __ffs(ok) - 1;
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      if (z != numa_id) continue;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t q = pod.sub[c][r];
        if (q == 0 || !((zs.zmask[z] >> r) & 1u)) continue;  // zero / QoS-exempt / resource missing in the zone
        const int64_t left = zs.avail[z][r] - q;
        if (left < 0) return B200S_REASON_NRT_ACCOUNTING;
        zs.avail[z][r] = left;
      }
    }
  }
  return B200S_REASON_OK;
}

// one zone, Least/Most/Balanced strategies
template <int Z, int R, int SC>
__device__ __forceinline__ int64_t strategy_score(const Zones<Z, R>& zs, int z, const NrtCfg& cfg, uint32_t req_mask,
                                                  const int64_t* req, const int64_t* reqv) {
  if constexpr (SC == 1) {
    // fractions of every requested resource first (independent divisions, statically indexed), one exit test after
    double fr[R];
    bool over = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      fr[r] = 0;
      if (!((req_mask >> r) & 1u)) continue;
      const int64_t cap = ((zs.zmask[z] >> r) & 1u) ? zs.avail[z][r] : 0;
      const int64_t cv = qty_value(cap);
      const double q = (double)reqv[r] / (double)(cv == 0 ? 1 : cv);
      const double f = cv == 0 ? 1.0 : q;
      over |= f > 1;
      fr[r] = f;
    }
    if (over) return 0;
    const int n = __popc(req_mask & ((1u << R) - 1u));
    double sum = 0;  // summation in resource-slot order over the requested resources, as before
#pragma unroll
    for (int r = 0; r < R; ++r)
      if ((req_mask >> r) & 1u) sum += fr[r];
    const double mean = sum / (double)n;
    double ss = 0, comp = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if ((req_mask >> r) & 1u) {
        const double d = fr[r] - mean;
        ss += d * d;
        comp += d;
      }
    const double variance = (ss - comp * comp / (double)n) / ((double)n - 1);
    return f2i((1 - variance) * 100.0);
  }
  const bool most = cfg.strategy == B200S_NRT_MOST_ALLOCATED;
  int64_t node_score = 0, weight_sum = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (!((req_mask >> r) & 1u)) continue;
    const int64_t cap = ((zs.zmask[z] >> r) & 1u) ? zs.avail[z][r] : 0;
    int64_t s;
    {  // evaluate unconditionally and select afterwards: the unrolled resources become independent straight-line
       // chains the scheduler can interleave (A/B on B200: 7.75 -> 7.29 ms at c4 against the branchy form)
      const int64_t cv = qty_value(cap) | (cap == 0 ? 1 : 0), rv = reqv[r];
      const bool zero = cap == 0 || req[r] > cap;
      const int64_t a = zero ? 0 : (most ? rv : cv - rv);
      s = zero ? 0 : div100(a, cv);
    }
    node_score = wrap_add(node_score, wrap_mul(s, cfg.w[r]));
    weight_sum = wrap_add(weight_sum, cfg.w[r]);
  }
  if (weight_sum == 0) return 0;
  // sum of (score <= 100) x weight over the weight sum: both fit 31 bits unless the weights are absurd
  if ((uint64_t)(node_score | weight_sum) < (1ull << 31)) return (int64_t)((uint32_t)node_score / (uint32_t)weight_sum);
  return go_div(node_score, weight_sum);
}

template <int Z, int R, int SC>
__device__ __forceinline__ int64_t score_each_numa(const Zones<Z, R>& zs, const NrtCfg& cfg, uint32_t req_mask,
                                                   const int64_t* req, const int64_t* reqv) {
  int64_t min_score = 0;
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    if (z >= zs.nz) continue;
    const int64_t s = strategy_score<Z, R, SC>(zs, z, cfg, req_mask, req, reqv);
    if (min_score == 0 || (s != 0 && s < min_score)) min_score = s;
  }
  return min_score;
}

template <int Z, int R>
__device__ __forceinline__ bool only_non_numa(const Zones<Z, R>& zs, uint32_t req_mask) {
  uint32_t any = 0;
#pragma unroll
  for (int z = 0; z < Z; ++z)
    if (z < zs.nz) any |= zs.zmask[z];
  return (any & req_mask) == 0;
}

template <int Z>
__device__ __forceinline__ float avg_distance(const int32_t (&cost)[Z][Z], uint32_t m, int k) {
  int accu = 0;
#pragma unroll
  for (int i = 0; i < Z; ++i) {
    if (!((m >> i) & 1u)) continue;
#pragma unroll
    for (int j = 0; j < Z; ++j) {
      if (!((m >> j) & 1u)) continue;
      const int c = cost[i][j];
      accu += c < 0 ? 255 : c;
    }
  }
  return (float)accu / (float)(k * k);
}

// numaNodesRequired + findSuitableCombination (least_numa.go:159-208); returns k (0 = cannot fit)
template <int Z, int R>
__device__ int numa_nodes_required(const Zones<Z, R>& zs, const int32_t (&cost)[Z][Z], const NrtCfg& cfg, int qos,
                                   uint32_t req_mask, const int64_t* req, uint32_t& mask_out, bool& is_min) {
  const int n = zs.nz;
  if (n == 0) return 0;
  for (int k = 1; k <= n; ++k) {
    const int lo = c_combo_off[n - 1][k], hi = c_combo_off[n - 1][k + 1];
    float min_avg = 255.0f;
    for (int i = lo; i < hi; ++i) {
      const float d = avg_distance<Z>(cost, c_combo_mask[n - 1][i], k);
      if (d < min_avg) min_avg = d;
    }
    bool have = false;
    uint32_t best = 0;
    float min_dist = 256.0f;
    for (int i = lo; i < hi; ++i) {
      const uint32_t m = c_combo_mask[n - 1][i];
      bool valid = true;
      int64_t sum[R];
#pragma unroll
      for (int r = 0; r < R; ++r) sum[r] = 0;
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        if (!((m >> z) & 1u)) continue;
        if ((zs.zmask[z] & req_mask) != req_mask) valid = false;  // isValidCombineResources
#pragma unroll
        for (int r = 0; r < R; ++r)
          if ((zs.zmask[z] >> r) & 1u) sum[r] += zs.avail[z][r];  // unlisted cells hold filter sentinels
      }
      if (!valid) continue;
      bool fit = true;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!((req_mask >> r) & 1u) || req[r] == 0) continue;
        if (!suitable(qos, cfg.res_flags[r], req[r], sum[r])) fit = false;
      }
      if (!fit) continue;
      const float dist = avg_distance<Z>(cost, m, k);
      if (dist == min_avg) {
        mask_out = m;
        is_min = true;
        return k;
      }
      if (dist < min_dist) {
        min_dist = dist;
        best = m;
        have = true;
      }
    }
    if (have) {
      mask_out = best;
      is_min = false;
      return k;
    }
  }
  return 0;
}

__device__ __forceinline__ int64_t normalize_least_numa(int count, bool is_min, int max_numa) {
  const int64_t unit = 100 / (int64_t)max_numa;
  const int64_t s = 100 - (int64_t)count * unit;
  return is_min ? s + unit / 2 : s;
}

template <int Z, int R, int SC>
__device__ int64_t nrt_score(const Zones<Z, R>& node_zs, const int32_t (&cost)[Z][Z], uint32_t nflags, int max_numa,
                             const NrtCfg& cfg, const PodS<R>& pod) {
  if (pod.qos != B200S_QOS_GUARANTEED) return 100;
  if (!(nflags & B200S_NRT_NODE_FRESH) || !(nflags & B200S_NRT_NODE_HAS_NRT)) return 0;
  if ((nflags & B200S_NRT_NODE_UNSUPPORTED) || (pod.flags & B200S_NRT_POD_UNSUPPORTED)) return 0;
  const bool scope_pod = nflags & B200S_NRT_NODE_SCOPE_POD;
  const int nc = pod.n_init + pod.n_app;
  const int steps = scope_pod ? 1 : nc;  // pod scope: one step on the pod-effective request (slot C_MAX)
  if constexpr (SC == 2) {
    // leastNUMAPodScopeScore :73-89 / leastNUMAContainerScopeScore :35-71 as one loop
    Zones<Z, R> zs = node_zs;
    int max_count = 0;
    bool all_min = true;
    for (int s = 0; s < steps; ++s) {
      const int c = scope_pod ? C_MAX : s;
      const uint32_t rm = pod.req_mask[c];
      if (only_non_numa<Z, R>(zs, rm)) continue;
      uint32_t mask = 0;
      bool is_min = false;
      const int k = numa_nodes_required<Z, R>(zs, cost, cfg, pod.qos, rm, pod.req[c], mask, is_min);
      if (k == 0) return 0;
      if (!is_min) all_min = false;
      if (k > max_count) max_count = k;
      if (scope_pod) break;
      // subtractFromNUMAs (numaresources.go:184-215)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!((rm >> r) & 1u)) continue;
        int64_t q = pod.req[c][r];
#pragma unroll
        for (int z = 0; z < Z; ++z) {
          if (z >= zs.nz || !((mask >> z) & 1u) || q == 0) continue;
          if (!((zs.zmask[z] >> r) & 1u)) continue;
          const int64_t av = zs.avail[z][r];
          if (q >= av) {
            q -= av;
            zs.avail[z][r] = 0;
          } else {
            zs.avail[z][r] = av - q;
            q = 0;
          }
        }
      }
    }
    return max_count == 0 ? 100 : normalize_least_numa(max_count, all_min, max_numa);
  } else {
    if (!(nflags & B200S_NRT_NODE_SINGLE_NUMA)) return 0;
    // podScopeScore :142-150 / containerScopeScore :152-165 (mean over init + app, no subtraction)
    double sum = 0;
    int64_t last = 0;
    for (int s = 0; s < steps; ++s) {
      const int c = scope_pod ? C_MAX : s;
      last = score_each_numa<Z, R, SC>(node_zs, cfg, pod.req_mask[c], pod.req[c], pod.reqv[c]);
      sum += (double)last;
    }
    return scope_pod ? last : f2i(sum / (double)nc);
  }
}

}  // namespace nrtdev
}  // namespace b200s

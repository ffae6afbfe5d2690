// NodeResourceTopologyMatch: Filter (single-numa-node, pod/container scope) + Score (Least/Most/
// Balanced/LeastNUMANodes), all pods x all nodes on the dense encoding of include/b200sched.h.
//
// Reference semantics (pkg/noderesourcetopology):
//   Filter   filter.go:176-225, handlers :39-78 / :162-173, resourcesAvailableInAnyNUMANodes :90-160,
//            isResourceSetSuitable numaresources.go:137-142, subtractResourcesFromNUMANodeList :145-182
//   Score    score.go:62-102, scoreForEachNUMANode :110-124, pod/container scope :142-165,
//            least_allocated.go:25-55, most_allocated.go:25-54, balanced_allocation.go:27-54,
//            least_numa.go:35-233, subtractFromNUMAs numaresources.go:184-215
// The reference re-parses zone names and deep-copies every Quantity map per (pod,node) call.
//
// B200 design (round-1 profile: the register-resident, fully unrolled version was 64 KB of SASS,
// instruction-fetch bound and ran 9 of 32 lanes): one thread owns one node; the CTA's 128 nodes'
// zones x resources blocks (exact milli-units) are staged ONCE per pod tile into shared memory as
// [zone][resource][thread] planes (conflict-free: consecutive lanes, consecutive 8-byte words), plus a
// working plane for the container-scope state machine.  Zone/resource loops stay rolled (small code,
// few registers, 4-8 CTAs per SM), any Z <= 8, R <= 8 without per-shape instantiation.  Pod records
// live in shared memory too and are warp-uniform (broadcast) reads.
#include "engine.h"

namespace b200s {

namespace {

constexpr int C_MAX = B200S_NRT_MAX_CONT;
constexpr int TPB = 128;  // threads (= nodes) per CTA

// Combination tables for LeastNUMANodes: for n zones, all non-empty subsets ordered by size then
// lexicographically by index tuple (gonum combin.Combinations order, least_numa.go:161).
__constant__ uint8_t c_combo_mask[8][255];
__constant__ uint8_t c_combo_off[8][10];  // [n-1][k] = first index of size-k subsets; [n-1][n+1] = end
bool g_combo_ready = false;

int ensure_combos() {
  if (g_combo_ready) return 0;
  static uint8_t mask[8][255];
  static uint8_t off[8][10];
  for (int n = 1; n <= 8; ++n) {
    int cnt = 0;
    for (int k = 1; k <= n; ++k) {
      off[n - 1][k] = (uint8_t)cnt;
      int idx[8];
      for (int i = 0; i < k; ++i) idx[i] = i;
      while (true) {
        uint8_t m = 0;
        for (int i = 0; i < k; ++i) m |= (uint8_t)(1u << idx[i]);
        mask[n - 1][cnt++] = m;
        int i = k - 1;
        while (i >= 0 && idx[i] == n - k + i) --i;
        if (i < 0) break;
        ++idx[i];
        for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
      }
    }
    off[n - 1][n + 1] = (uint8_t)cnt;
    off[n - 1][0] = 0;
  }
  if (cudaMemcpyToSymbol(c_combo_mask, mask, sizeof(mask)) != cudaSuccess) return -1;
  if (cudaMemcpyToSymbol(c_combo_off, off, sizeof(off)) != cudaSuccess) return -1;
  g_combo_ready = true;
  return 0;
}

struct PodS {  // one pod of the tile, in shared memory (R_MAX columns; columns >= R are zero)
  int64_t req[C_MAX + 1][B200S_NRT_MAX_RES];
  int64_t reqv[C_MAX + 1][B200S_NRT_MAX_RES];  // Quantity.Value() of req (ceil to whole units)
  uint8_t req_mask[C_MAX + 1];
  uint8_t kind[C_MAX];
  uint8_t qos, flags, n_init, n_app;
};

struct NrtCfg {
  int strategy;
  int64_t w[B200S_NRT_MAX_RES];
  uint8_t res_flags[B200S_NRT_MAX_RES];
};

// This thread's view of its node in shared memory: planes of TPB int64, element (z, r) at [(z*R + r)*TPB + tid].
struct NodeView {
  const int64_t* base;   // zone Available as published (never modified)
  int64_t* work;         // working copy for the subtracting state machines
  const uint8_t* zmask;  // [Z][TPB]
  const int32_t* cost;   // [Z][Z][TPB] (LeastNUMANodes only)
  int R, nz, tid;
  uint32_t node_res_mask, flags;
  int max_numa;
  __device__ __forceinline__ int64_t av(const int64_t* buf, int z, int r) const { return buf[(z * R + r) * TPB + tid]; }
  __device__ __forceinline__ uint32_t zm(int z) const { return zmask[z * TPB + tid]; }
};

__device__ __forceinline__ int64_t qty_value(int64_t milli) { return (milli + 999) / 1000; }  // Quantity.Value(), >= 0
__device__ __forceinline__ int64_t f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
// floor(a*100/cv) for 0 <= a <= cv (quotient in 0..100): fp32 estimate + exact integer fix-up instead of
// a 64-bit division; identical to Go's (a*100)/cv.  Absurdly large capacities keep the wrapping Go path.
__device__ __forceinline__ int64_t div100(int64_t a, int64_t cv) {
  if (cv >= (1ll << 52) || a < 0 || a > cv) return go_div(wrap_mul(a, 100), cv);
  const int64_t num = a * 100;
  int64_t q = (int64_t)__float2int_rd(__fdividef(__ll2float_rn(num), __ll2float_rn(cv)));
  int64_t rem = num - q * cv;
  while (rem < 0) {
    --q;
    rem += cv;
  }
  while (rem >= cv) {
    ++q;
    rem -= cv;
  }
  return q;
}

__device__ __forceinline__ bool suitable(int qos, uint32_t rflags, int64_t qty, int64_t numa_qty) {
  if (qos != B200S_QOS_GUARANTEED && (rflags & B200S_NRT_RES_AFFINE)) return true;
  return numa_qty >= qty;
}

// resourcesAvailableInAnyNUMANodes (filter.go:90-160) against the planes in `buf`
__device__ bool available_in_any(const NodeView& nv, const int64_t* buf, const NrtCfg& cfg, int qos, uint32_t req_mask,
                                 const int64_t* req, int& numa_id) {
  uint32_t bitmask = 0xffffffffu;  // only bits < nz <= 8 can be cleared; all-ones <=> untouched
  for (int r = 0; r < nv.R; ++r) {
    if (!((req_mask >> r) & 1u)) continue;
    const int64_t q = req[r];
    if (q == 0) continue;                                 // :101-105
    if (!((nv.node_res_mask >> r) & 1u)) return false;    // :107-113
    bool has_affinity = false;
    uint32_t res_bits = 0;
    for (int z = 0; z < nv.nz; ++z) {
      if (!((nv.zm(z) >> r) & 1u)) continue;              // :121-125
      has_affinity = true;
      if (suitable(qos, cfg.res_flags[r], q, nv.av(buf, z, r))) res_bits |= 1u << z;  // :128-133
    }
    if (!has_affinity && (cfg.res_flags[r] & B200S_NRT_RES_HOST_LEVEL)) continue;      // :139-142
    bitmask &= res_bits;
    if (bitmask == 0) return false;                       // :144-148
  }
  numa_id = __ffs(bitmask) - 1;                           // :154
  return true;
}

__device__ __forceinline__ void reset_work(const NodeView& nv) {
  const int n = nv.nz * nv.R;
  for (int i = 0; i < n; ++i) nv.work[i * TPB + nv.tid] = nv.base[i * TPB + nv.tid];
}

// TopologyMatch.Filter -> reason code.  One loop for both scopes: pod scope = a single step on the
// pod-effective request (slot C_MAX, singleNUMAPodLevelHandler :162-173); container scope = init
// containers without subtraction, then app containers with it (:39-78).
__device__ int nrt_filter(const NodeView& nv, const NrtCfg& cfg, const PodS& pod) {
  if (pod.flags & B200S_NRT_POD_FILTER_BYPASS) return B200S_REASON_OK;  // :180-183
  if ((nv.flags & B200S_NRT_NODE_UNSUPPORTED) || (pod.flags & B200S_NRT_POD_UNSUPPORTED)) return B200S_REASON_UNSUPPORTED;
  if (!(nv.flags & B200S_NRT_NODE_FRESH)) return B200S_REASON_NRT_INVALID_TOPOLOGY;  // :194-197
  if (!(nv.flags & B200S_NRT_NODE_HAS_NRT)) return B200S_REASON_OK;                  // :198-200
  if (!(nv.flags & B200S_NRT_NODE_SINGLE_NUMA)) return B200S_REASON_OK;              // :206-209
  const bool scope_pod = nv.flags & B200S_NRT_NODE_SCOPE_POD;
  const int n_init = pod.n_init, steps = scope_pod ? 1 : n_init + pod.n_app;
  const bool subtracts = !scope_pod && pod.n_app > 1;  // the last app container's subtraction is never read
  if (subtracts) reset_work(nv);
  const int64_t* buf = subtracts ? nv.work : nv.base;
  for (int s = 0; s < steps; ++s) {
    const int c = scope_pod ? C_MAX : s;
    const uint32_t rm = pod.req_mask[c];
    int numa_id = 0;
    if (!available_in_any(nv, buf, cfg, pod.qos, rm, pod.req[c], numa_id)) {
      if (scope_pod) return B200S_REASON_NRT_ALIGN_POD;
      if (s >= n_init) return B200S_REASON_NRT_ALIGN_CONTAINER;
      return pod.kind[c] == B200S_CONT_SIDECAR ? B200S_REASON_NRT_ALIGN_SIDECAR : B200S_REASON_NRT_ALIGN_INIT;
    }
    if (scope_pod || s < n_init) continue;
    // subtractResourcesFromNUMANodeList (numaresources.go:145-182): zone with NUMA id == numa_id
    if (numa_id >= nv.nz) continue;
    for (int r = 0; r < nv.R; ++r) {
      if (!((rm >> r) & 1u)) continue;
      if (pod.qos != B200S_QOS_GUARANTEED && (cfg.res_flags[r] & B200S_NRT_RES_AFFINE)) continue;
      const int64_t q = pod.req[c][r];
      if (q == 0) continue;
      if (!((nv.zm(numa_id) >> r) & 1u)) continue;
      const int64_t left = nv.av(buf, numa_id, r) - q;
      if (left < 0) return B200S_REASON_NRT_ACCOUNTING;
      if (subtracts) nv.work[(numa_id * nv.R + r) * TPB + nv.tid] = left;
    }
  }
  return B200S_REASON_OK;
}

// one zone, Least/Most (SC 0) or Balanced (SC 1)
template <int SC>
__device__ int64_t strategy_score(const NodeView& nv, int z, const NrtCfg& cfg, uint32_t req_mask, const int64_t* req,
                                  const int64_t* reqv) {
  const uint32_t zmk = nv.zm(z);
  if constexpr (SC == 1) {
    // balancedAllocationScoreStrategy + gonum stat.Variance (corrected two-pass, / (n-1)); resource-slot order
    double fr[B200S_NRT_MAX_RES];
    int n = 0;
    for (int r = 0; r < nv.R; ++r) {
      if (!((req_mask >> r) & 1u)) continue;
      const int64_t cap = ((zmk >> r) & 1u) ? nv.av(nv.base, z, r) : 0;
      const int64_t cv = qty_value(cap);
      const double f = cv == 0 ? 1.0 : (double)reqv[r] / (double)cv;
      if (f > 1) return 0;
      fr[n++] = f;
    }
    double sum = 0;
    for (int i = 0; i < n; ++i) sum += fr[i];
    const double mean = sum / (double)n;
    double ss = 0, comp = 0;
    for (int i = 0; i < n; ++i) {
      const double d = fr[i] - mean;
      ss += d * d;
      comp += d;
    }
    const double variance = (ss - comp * comp / (double)n) / ((double)n - 1);
    return f2i((1 - variance) * 100.0);
  } else {
    const bool most = cfg.strategy == B200S_NRT_MOST_ALLOCATED;
    int64_t node_score = 0, weight_sum = 0;
    for (int r = 0; r < nv.R; ++r) {
      if (!((req_mask >> r) & 1u)) continue;
      const int64_t cap = ((zmk >> r) & 1u) ? nv.av(nv.base, z, r) : 0;  // missing key -> zero Quantity
      int64_t s;
      if (cap == 0 || req[r] > cap) {
        s = 0;
      } else {
        const int64_t cv = qty_value(cap), rv = reqv[r];
        s = most ? div100(rv, cv) : div100(cv - rv, cv);
      }
      node_score = wrap_add(node_score, wrap_mul(s, cfg.w[r]));
      weight_sum = wrap_add(weight_sum, cfg.w[r]);
    }
    if (weight_sum == 0) return 0;
    return go_div(node_score, weight_sum);
  }
}

// scoreForEachNUMANode (score.go:110-124): minimum of the non-zero zone scores, 0 if every zone scores 0
template <int SC>
__device__ int64_t score_each_numa(const NodeView& nv, const NrtCfg& cfg, uint32_t req_mask, const int64_t* req,
                                   const int64_t* reqv) {
  int64_t min_score = 0;
  for (int z = 0; z < nv.nz; ++z) {
    const int64_t s = strategy_score<SC>(nv, z, cfg, req_mask, req, reqv);
    if (min_score == 0 || (s != 0 && s < min_score)) min_score = s;
  }
  return min_score;
}

__device__ bool only_non_numa(const NodeView& nv, uint32_t req_mask) {  // pluginhelpers.go:163-173
  uint32_t any = 0;
  for (int z = 0; z < nv.nz; ++z) any |= nv.zm(z);
  return (any & req_mask) == 0;
}

__device__ float avg_distance(const NodeView& nv, int Z, uint32_t m, int k) {  // least_numa.go:116-138
  int accu = 0;
  for (int i = 0; i < nv.nz; ++i) {
    if (!((m >> i) & 1u)) continue;
    for (int j = 0; j < nv.nz; ++j) {
      if (!((m >> j) & 1u)) continue;
      const int c = nv.cost[(i * Z + j) * TPB + nv.tid];
      accu += c < 0 ? 255 : c;
    }
  }
  return (float)accu / (float)(k * k);
}

// numaNodesRequired + findSuitableCombination (least_numa.go:159-208) on the WORK planes; returns k (0 = cannot fit)
__device__ int numa_nodes_required(const NodeView& nv, int Z, const NrtCfg& cfg, int qos, uint32_t req_mask,
                                   const int64_t* req, uint32_t& mask_out, bool& is_min) {
  const int n = nv.nz;
  if (n == 0) return 0;
  for (int k = 1; k <= n; ++k) {
    const int lo = c_combo_off[n - 1][k], hi = c_combo_off[n - 1][k + 1];
    float min_avg = 255.0f;
    for (int i = lo; i < hi; ++i) {
      const float d = avg_distance(nv, Z, c_combo_mask[n - 1][i], k);
      if (d < min_avg) min_avg = d;
    }
    bool have = false;
    uint32_t best = 0;
    float min_dist = 256.0f;
    for (int i = lo; i < hi; ++i) {
      const uint32_t m = c_combo_mask[n - 1][i];
      bool valid = true;  // isValidCombineResources :224-233
      for (int z = 0; z < n; ++z)
        if (((m >> z) & 1u) && (nv.zm(z) & req_mask) != req_mask) valid = false;
      if (!valid) continue;
      bool fit = true;  // combineResources + checkResourcesFit
      for (int r = 0; r < nv.R && fit; ++r) {
        if (!((req_mask >> r) & 1u) || req[r] == 0) continue;
        int64_t sum = 0;
        for (int z = 0; z < n; ++z)
          if ((m >> z) & 1u) sum += nv.av(nv.work, z, r);
        if (!suitable(qos, cfg.res_flags[r], req[r], sum)) fit = false;
      }
      if (!fit) continue;
      const float dist = avg_distance(nv, Z, m, k);
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

__device__ __forceinline__ int64_t normalize_least_numa(int count, bool is_min, int max_numa) {  // :91-100
  const int64_t unit = 100 / (int64_t)max_numa;
  const int64_t s = 100 - (int64_t)count * unit;
  return is_min ? s + unit / 2 : s;
}

template <int SC>
__device__ int64_t nrt_score(const NodeView& nv, int Z, const NrtCfg& cfg, const PodS& pod) {
  if (pod.qos != B200S_QOS_GUARANTEED) return 100;  // score.go:72-75
  if ((nv.flags & B200S_NRT_NODE_UNSUPPORTED) || (pod.flags & B200S_NRT_POD_UNSUPPORTED)) return 0;
  if (!(nv.flags & B200S_NRT_NODE_FRESH) || !(nv.flags & B200S_NRT_NODE_HAS_NRT)) return 0;  // :79-86
  const bool scope_pod = nv.flags & B200S_NRT_NODE_SCOPE_POD;
  const int nc = pod.n_init + pod.n_app;
  const int steps = scope_pod ? 1 : nc;  // pod scope: one step on the pod-effective request (slot C_MAX)
  if constexpr (SC == 2) {
    // leastNUMAPodScopeScore :73-89 / leastNUMAContainerScopeScore :35-71 as one loop
    reset_work(nv);
    int max_count = 0;
    bool all_min = true;
    for (int s = 0; s < steps; ++s) {
      const int c = scope_pod ? C_MAX : s;
      const uint32_t rm = pod.req_mask[c];
      if (only_non_numa(nv, rm)) continue;
      uint32_t mask = 0;
      bool is_min = false;
      const int k = numa_nodes_required(nv, Z, cfg, pod.qos, rm, pod.req[c], mask, is_min);
      if (k == 0) return 0;
      if (!is_min) all_min = false;
      if (k > max_count) max_count = k;
      if (scope_pod) break;
      // subtractFromNUMAs (numaresources.go:184-215): greedy across the chosen zones in bit order
      for (int r = 0; r < nv.R; ++r) {
        if (!((rm >> r) & 1u)) continue;
        int64_t q = pod.req[c][r];
        for (int z = 0; z < nv.nz && q != 0; ++z) {
          if (!((mask >> z) & 1u) || !((nv.zm(z) >> r) & 1u)) continue;
          int64_t* cell = &nv.work[(z * nv.R + r) * TPB + nv.tid];
          const int64_t avl = *cell;
          if (q >= avl) {
            q -= avl;
            *cell = 0;
          } else {
            *cell = avl - q;
            q = 0;
          }
        }
      }
    }
    return max_count == 0 ? 100 : normalize_least_numa(max_count, all_min, nv.max_numa);
  } else {
    if (!(nv.flags & B200S_NRT_NODE_SINGLE_NUMA)) return 0;  // :177-179
    // podScopeScore :142-150 / containerScopeScore :152-165 (mean over init + app, no subtraction)
    double sum = 0;
    int64_t last = 0;
    for (int s = 0; s < steps; ++s) {
      const int c = scope_pod ? C_MAX : s;
      last = score_each_numa<SC>(nv, cfg, pod.req_mask[c], pod.req[c], pod.reqv[c]);
      sum += (double)last;
    }
    return scope_pod ? last : f2i(sum / (double)nc);
  }
}

struct NrtNodeCols {
  const uint8_t* node_flags;
  const uint16_t* max_numa;
  const uint8_t* nz;
  const uint8_t* node_res_mask;
  const uint8_t* zone_res_mask;  // [Z][Npad]
  const int64_t* avail;          // [Z][R][Npad]
  const int32_t* cost;           // [Z][Z][Npad] or null
  int Z, R;
};
struct NrtPodCols {
  const uint8_t* qos;
  const uint8_t* flags;
  const uint8_t* n_init;
  const uint8_t* n_app;
  const uint8_t* kind;      // [P][8]
  const uint8_t* req_mask;  // [P][9]
  const int64_t* req;       // [P][9][R]
};

constexpr int PT = 16;

__host__ __device__ inline size_t nrt_smem_bytes(int Z, int R, bool with_cost) {
  size_t b = (size_t)2 * Z * R * TPB * 8;          // base + work planes
  b += with_cost ? (size_t)Z * Z * TPB * 4 : 0;    // cost planes
  b += (size_t)Z * TPB;                            // zone masks
  b = (b + 15) & ~(size_t)15;
  b += sizeof(PodS) * PT;
  return b;
}

template <int SC, class OutT>
__global__ void __launch_bounds__(TPB)
nrt_kernel(NrtNodeCols nc, NrtPodCols pc, NrtCfg cfg, const uint64_t* __restrict__ upstream, int words, int N,
           int Npad, int P, OutT* __restrict__ out, uint32_t* __restrict__ feas_out32, uint8_t* __restrict__ reasons) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int Z = nc.Z, R = nc.R, tid = threadIdx.x;
  int64_t* s_base = reinterpret_cast<int64_t*>(smem);
  int64_t* s_work = s_base + (size_t)Z * R * TPB;
  int32_t* s_cost = reinterpret_cast<int32_t*>(s_work + (size_t)Z * R * TPB);
  uint8_t* s_zmask = reinterpret_cast<uint8_t*>(s_cost + (SC == 2 ? (size_t)Z * Z * TPB : 0));
  PodS* sp = reinterpret_cast<PodS*>(smem + nrt_smem_bytes(Z, R, SC == 2) - sizeof(PodS) * PT);

  const int n = blockIdx.x * TPB + tid;  // Npad is a multiple of TPB: every thread has a (possibly padding) node
  const int p0 = blockIdx.y * PT, pend = min(PT, P - p0);
  // ---- stage the pod tile
  for (int i = tid; i < pend * (C_MAX + 1) * B200S_NRT_MAX_RES; i += TPB) {
    const int pp = i / ((C_MAX + 1) * B200S_NRT_MAX_RES), rest = i % ((C_MAX + 1) * B200S_NRT_MAX_RES);
    const int c = rest / B200S_NRT_MAX_RES, r = rest % B200S_NRT_MAX_RES;
    const int64_t q = r < R ? pc.req[((size_t)(p0 + pp) * (C_MAX + 1) + c) * R + r] : 0;
    sp[pp].req[c][r] = q;
    sp[pp].reqv[c][r] = q >= 0 ? (q + 999) / 1000 : -((-q) / 1000);
  }
  for (int i = tid; i < pend; i += TPB) {
    const int p = p0 + i;
    sp[i].qos = pc.qos[p];
    sp[i].flags = pc.flags[p];
    sp[i].n_init = pc.n_init[p];
    sp[i].n_app = pc.n_app[p];
    for (int c = 0; c < C_MAX; ++c) sp[i].kind[c] = pc.kind[(size_t)p * C_MAX + c];
    for (int c = 0; c <= C_MAX; ++c) sp[i].req_mask[c] = pc.req_mask[(size_t)p * (C_MAX + 1) + c];
  }
  // ---- stage this thread's node: coalesced column reads -> [z][r][tid] planes
  NodeView nv;
  nv.base = s_base;
  nv.work = s_work;
  nv.zmask = s_zmask;
  nv.cost = s_cost;
  nv.R = R;
  nv.tid = tid;
  nv.flags = n < N ? nc.node_flags[n] : 0;
  nv.node_res_mask = nc.node_res_mask[n];
  nv.max_numa = max((int)nc.max_numa[n], 1);
  nv.nz = min((int)nc.nz[n], Z);
  for (int z = 0; z < Z; ++z) {
    s_zmask[z * TPB + tid] = nc.zone_res_mask[(size_t)z * Npad + n];
    for (int r = 0; r < R; ++r) s_base[(z * R + r) * TPB + tid] = nc.avail[((size_t)z * R + r) * Npad + n];
    if constexpr (SC == 2)
      for (int z2 = 0; z2 < Z; ++z2) s_cost[(z * Z + z2) * TPB + tid] = nc.cost[((size_t)z * Z + z2) * Npad + n];
  }
  __syncthreads();
  const int lane = tid & 31;
  const int word = n >> 6, half = (n >> 5) & 1;
  for (int pp = 0; pp < pend; ++pp) {
    const int p = p0 + pp;
    const PodS& pod = sp[pp];
    int reason = 0;
    bool feasible = false;
    int64_t score = 0;
    if (n < N) {
      reason = nrt_filter(nv, cfg, pod);
      const bool up = upstream ? ((upstream[(size_t)p * words + word] >> (n & 63)) & 1ull) : true;
      feasible = reason == 0 && up;
      if (reason == 0 && !up) reason = B200S_REASON_UPSTREAM;
      if (feasible) score = nrt_score<SC>(nv, Z, cfg, pod);
    }
    const uint32_t fw = __ballot_sync(0xffffffffu, feasible);
    if (lane == 0) feas_out32[((size_t)p * words + word) * 2 + half] = fw;
    out[(size_t)p * Npad + n] = (OutT)score;
    reasons[(size_t)p * Npad + n] = (uint8_t)reason;
  }
}

template <int SC, class OutT>
int launch_one(b200s_ctx* c, const NrtNodeCols& nc, const NrtPodCols& pc, const NrtCfg& cfg, OutT* out) {
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  const size_t smem = nrt_smem_bytes(nc.Z, nc.R, SC == 2);
  auto kern = nrt_kernel<SC, OutT>;
  B200S_CUDA_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(Npad / TPB, (P + PT - 1) / PT);
  kern<<<grid, TPB, smem, c->stream>>>(nc, pc, cfg, c->upstream_mask(), words, N, Npad, P, out, o.feas.as<uint32_t>(),
                                       o.reasons.as<uint8_t>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

template <int SC>
int launch_sc(b200s_ctx* c, int dtype, const NrtNodeCols& nc, const NrtPodCols& pc, const NrtCfg& cfg) {
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  if (dtype == B200S_OUT_I64) return launch_one<SC, int64_t>(c, nc, pc, cfg, o.scores.as<int64_t>());
  return launch_one<SC, uint8_t>(c, nc, pc, cfg, o.scores.as<uint8_t>());
}

}  // namespace

int nrt_eval(b200s_ctx* c, int dtype) {
  if (!c->has_nrt) return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: snapshot has no NRT columns");
  if (!c->has_nrt_pods) return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: pod batch has no NRT columns");
  if (c->nrt_strategy == B200S_NRT_LEAST_NUMA_NODES && !c->nrt_has_cost)
    return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: LeastNUMANodes needs the zone cost columns");
  if (ensure_combos() != 0) return c->set_err(B200S_ERR_CUDA, "NodeResourceTopologyMatch: combination table upload failed");
  B200S_TRY(ensure_out(c, B200S_PLUGIN_NRT, dtype, true, true));
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  if (c->P == 0) {
    o.valid = true;
    return B200S_OK;
  }
  NrtNodeCols nc{c->nrt_node_flags.as<uint8_t>(), c->nrt_max_numa.as<uint16_t>(), c->nrt_nz.as<uint8_t>(),
                 c->nrt_node_res_mask.as<uint8_t>(), c->nrt_zone_res_mask.as<uint8_t>(), c->nrt_avail.as<int64_t>(),
                 c->nrt_has_cost ? c->nrt_cost.as<int32_t>() : nullptr, c->nrt_Z, c->nrt_R};
  NrtPodCols pc{c->nrt_pod_qos.as<uint8_t>(), c->nrt_pod_flags.as<uint8_t>(), c->nrt_pod_ninit.as<uint8_t>(),
                c->nrt_pod_napp.as<uint8_t>(), c->nrt_pod_kind.as<uint8_t>(), c->nrt_pod_req_mask.as<uint8_t>(),
                c->nrt_pod_req.as<int64_t>()};
  NrtCfg cfg;
  cfg.strategy = c->nrt_strategy;
  for (int r = 0; r < B200S_NRT_MAX_RES; ++r) {
    cfg.w[r] = c->nrt_w[r];
    cfg.res_flags[r] = c->nrt_res_flags[r];
  }
  KernelTimer kt(c, B200S_PLUGIN_NRT);
  int rc;
  if (cfg.strategy == B200S_NRT_LEAST_NUMA_NODES)
    rc = launch_sc<2>(c, dtype, nc, pc, cfg);
  else if (cfg.strategy == B200S_NRT_BALANCED_ALLOCATION)
    rc = launch_sc<1>(c, dtype, nc, pc, cfg);
  else
    rc = launch_sc<0>(c, dtype, nc, pc, cfg);
  if (rc != B200S_OK) return rc;
  o.valid = true;
  return B200S_OK;
}

}  // namespace b200s

// The scheduling cycle in TWO launches (one graph launch): all enabled plugins, chained filters, NormalizeScore, weighted sum and the
// per-pod top-k for a handful of pods (the real scheduler's shape is P = 1) over all nodes of the shard.
//
// combined.cu evaluates a profile plugin by plugin (13 launches for the five plugins at P = 1: each plugin writes its
// u8 row, combine_topk_kernel reads them back).  At P = 1 those kernels are a few microseconds each and the cycle is
// launch latency.  Here two small back-to-back launches do the whole upstream cycle (SURVEY App. B "Upstream
// combination"; a single cooperative kernel with two grid barriers was measured first: its launch + barrier cost,
// 21 us at 50k nodes, exceeded the two extra launches, and its co-residency limit kept the grid at one CTA per SM):
//   phase 1  per node: upstream bit AND NodeResourceTopologyMatch.Filter (filter.go:176) AND NetworkOverhead.Filter
//            (networkoverhead.go:326) -> feasible; the scores that need no normalisation (TopologyMatch.Score,
//            TargetLoadPacking.Score, LoadVariationRiskBalancing.Score) and NetworkOverhead's accumulated cost go to a
//            small L2-resident scratch row; per-pod min/max of the Allocatable raw score and of the NetworkOverhead
//            cost over the FEASIBLE nodes (allocatable.go:145-155, networkoverhead.go:421-435) by block reduction +
//            one atomic per block
//   phase 2  NormalizeScore of the two normalising plugins, sum of weight x score (upstream RunScorePlugins), per-block
//            top-k under (total desc, node asc)
//            ... and the CTA of phase 2 that finishes last folds the per-block winners.
// No per-plugin matrix is materialised.  Single GPU (a sharded cycle needs the min/max exchange between phases 1 and
// 2 and keeps the plugin-by-plugin path), Least/Most/BalancedAllocation, <= 4 zones x <= 4 resource slots, P <= 4.
#include <algorithm>

#include "engine.h"
#include "netoh_device.cuh"
#include "nrt_device.cuh"
#include "trimaran_device.cuh"

namespace b200s {

namespace {

using namespace nrtdev;
using namespace netdev;
using namespace tridev;

constexpr int PMAX = 4, KMAX = 16, Z = 4, R = 4;

struct CycleArgs {
  uint32_t mask;
  int P, N, Npad, words, node_off, k;
  int64_t w[B200S_PLUGIN_COUNT];
  const uint64_t* upstream;
  // NodeResourcesAllocatable
  const int64_t* alloc_raw;
  // TargetLoadPacking
  const double* tlp_util;
  const int64_t *tlp_cap, *tlp_missing, *tlp_pod;
  const uint8_t* tlp_flags;
  int64_t tlp_target;
  // LoadVariationRiskBalancing
  const double* lvrb_f64;
  const int64_t *lvrb_i64, *lvrb_req_cpu, *lvrb_req_mem;
  const uint8_t* lvrb_flags;
  double margin, sens;
  // NodeResourceTopologyMatch
  const uint8_t *nrt_node_flags, *nrt_nz, *nrt_node_res_mask, *nrt_zone_res_mask;
  const int64_t* nrt_avail;
  int nrt_Zs, nrt_Rs;
  const uint8_t *pod_qos, *pod_flags, *pod_ninit, *pod_napp, *pod_kind, *pod_req_mask;
  const int64_t* pod_req;
  NrtCfg cfg;
  // NetworkOverhead
  Topo topo;
  const uint16_t *region, *zone;
  const uint8_t* equal;
  const int32_t* dep_off;
  const b200s_netoh_dep* deps;
  // scratch + output
  unsigned long long* lohi;  // [PMAX][4] order-preserving unsigned images, zero-initialised: ~lo / hi of Allocatable, NetworkOverhead
  uint32_t* part;            // [P][Npad] feasible << 24 | nrt << 16 | lvrb << 8 | tlp
  int64_t* cost;             // [P][Npad] NetworkOverhead accumulated cost
  b200s_topk_entry* block_best;  // [P][gridDim.x][k]
  b200s_topk_entry* out;         // [P][k]
  uint32_t* feas32;              // [P][Npad/32] the final feasible set (what b200s_fetch_total_feasible returns)
  const int32_t* perm;           // [Npad] thread slot -> node, nodes grouped by NodeResourceTopologyMatch control-flow class
  int blocks2;                   // CTAs of phase 2 (rows of block_best)
  unsigned int* done;            // ticket counter of phase 2 (zero between cycles)
};

__device__ __forceinline__ unsigned long long ord(int64_t x) { return (unsigned long long)x ^ 0x8000000000000000ull; }
__device__ __forceinline__ int64_t unord(unsigned long long u) { return (int64_t)(u ^ 0x8000000000000000ull); }
__device__ __forceinline__ bool better(int64_t s1, int32_t n1, int64_t s2, int32_t n2) {
  return s1 > s2 || (s1 == s2 && n1 < n2);
}

// TargetLoadPacking.Score (targetloadpacking.go:107-187), the arithmetic of tlp_kernel's IEEE branch
__device__ __forceinline__ int64_t tlp_score(double util, int64_t cap, int64_t missing, uint32_t flags, double pod_cpu, double t) {
  if (!(flags & B200S_TLP_HAS_METRICS) || !(flags & B200S_TLP_CPU_FOUND)) return 0;
  const double ncap = (double)cap;
  const double base = (util / 100) * ncap;
  double predicted = 0;
  if (ncap != 0) predicted = 100 * (base + pod_cpu + (double)missing) / ncap;
  const double hmt = 100 - t;
  const bool over = predicted > t;
  const double num = over ? t * (100 - predicted) : hmt * predicted;
  double quo = num / (over ? hmt : t);
  quo = over ? quo : quo + t;
  double s = go_round(quo);
  s = (over && predicted > 100) ? 0.0 : s;
  return go_f2i(s);
}

// Allocatable.NormalizeScore (allocatable.go:143-168) for one node
__device__ __forceinline__ int64_t alloc_norm(int64_t s, int64_t lo, int64_t hi, bool any) {
  if (!any) return 0;
  const int64_t range = wrap_sub(hi, lo);
  return range == 0 ? 0 : go_div(wrap_mul(wrap_sub(s, lo), 100), range);
}
// NetworkOverhead.NormalizeScore (networkoverhead.go:389-418) for one node
__device__ __forceinline__ int64_t netoh_norm(int64_t s, int64_t lo, int64_t hi, bool any) {
  if (!any || (lo == 0 && hi == 0)) return 0;  // empty list / both zero: scores stay as they are (all 0)
  if (lo == hi) return 100;
  const double norm = 100.0 * (double)wrap_sub(s, lo) / (double)wrap_sub(hi, lo);
  const int64_t tr = (norm >= -9223372036854775808.0 && norm < 9223372036854775808.0) ? (int64_t)norm : INT64_MIN;
  return wrap_sub(100, tr);
}

// phase 1: one thread per node slot (through the class-sorted permutation when NodeResourceTopologyMatch is on, so the
// lanes of a warp share their control flow)
template <int SC>
__global__ void __launch_bounds__(128, 3) cycle_phase1_kernel(CycleArgs a) {
  __shared__ PodS<R> sp[PMAX];
  __shared__ unsigned long long s_red[PMAX][4];
  const int t = threadIdx.x, P = a.P;
  const bool use_alloc = a.mask & (1u << B200S_PLUGIN_ALLOCATABLE), use_tlp = a.mask & (1u << B200S_PLUGIN_TLP);
  const bool use_lvrb = a.mask & (1u << B200S_PLUGIN_LVRB), use_nrt = a.mask & (1u << B200S_PLUGIN_NRT);
  const bool use_net = a.mask & (1u << B200S_PLUGIN_NETWORK_OVERHEAD);
  // ---- stage the pods' NodeResourceTopologyMatch records (as nrt_kernel does)
  if (use_nrt) {
    for (int i = t; i < P * (C_MAX + 1) * R; i += 128) {
      const int pp = i / ((C_MAX + 1) * R), rest = i % ((C_MAX + 1) * R), c = rest / R, r = rest % R;
      const int64_t q = r < a.nrt_Rs ? a.pod_req[((size_t)pp * (C_MAX + 1) + c) * a.nrt_Rs + r] : 0;
      sp[pp].req[c][r] = q;
      sp[pp].reqv[c][r] = q >= 0 ? (q + 999) / 1000 : -((-q) / 1000);
      const uint32_t m = r < a.nrt_Rs ? a.pod_req_mask[(size_t)pp * (C_MAX + 1) + c] : 0u;
      const bool needed = ((m >> r) & 1u) && q != 0;
      const bool exempt = a.pod_qos[pp] != B200S_QOS_GUARANTEED && (a.cfg.res_flags[r] & B200S_NRT_RES_AFFINE);
      sp[pp].eff[c][r] = needed ? (exempt ? INT64_MIN + 1 : q) : INT64_MIN;
      sp[pp].sub[c][r] = (needed && !exempt) ? q : 0;
    }
    for (int i = t; i < P; i += 128) {
      sp[i].qos = a.pod_qos[i];
      sp[i].flags = a.pod_flags[i];
      sp[i].n_init = a.pod_ninit[i];
      sp[i].n_app = a.pod_napp[i];
      for (int c = 0; c < C_MAX; ++c) sp[i].kind[c] = a.pod_kind[(size_t)i * C_MAX + c];
      for (int c = 0; c <= C_MAX; ++c) {
        const uint32_t m = a.pod_req_mask[(size_t)i * (C_MAX + 1) + c];
        sp[i].req_mask[c] = (uint8_t)m;
        uint32_t need = 0;
        for (int r = 0; r < a.nrt_Rs; ++r)
          if (((m >> r) & 1u) && a.pod_req[((size_t)i * (C_MAX + 1) + c) * a.nrt_Rs + r] != 0) need |= 1u << r;
        sp[i].need[c] = (uint8_t)need;
      }
    }
  }
  if (t < PMAX * 4) s_red[t / 4][t % 4] = 0;
  __syncthreads();

  // ---- phase 1: filters, un-normalised scores, block min/max over the feasible nodes
  const int cost_dummy[Z][Z] = {};
  (void)cost_dummy;
  {
    const int slot = blockIdx.x * 128 + t;  // < Npad: the grid covers exactly Npad / 128 CTAs
    const int n = a.perm ? a.perm[slot] : slot;
    const bool in = n < a.N;
    Zones<Z, R> zs;
    uint32_t nflags = 0, node_res_mask = 0;
    zs.nz = 0;
    if (use_nrt && in) {
      nflags = a.nrt_node_flags[n];
      node_res_mask = a.nrt_node_res_mask[n];
      zs.nz = min((int)a.nrt_nz[n], Z);
    }
    if (use_nrt) {
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        zs.zmask[z] = (in && z < a.nrt_Zs && z < zs.nz) ? a.nrt_zone_res_mask[(size_t)z * a.Npad + n] : 0;
#pragma unroll
        for (int r = 0; r < R; ++r)
          zs.avail[z][r] = (in && z < a.nrt_Zs && r < a.nrt_Rs) ? a.nrt_avail[((size_t)z * a.nrt_Rs + r) * a.Npad + n] : 0;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {  // filter encoding of the node (see nrt_filter)
        uint32_t any = 0;
#pragma unroll
        for (int z = 0; z < Z; ++z) any |= (zs.zmask[z] >> r) & 1u;
        const int64_t none = (!any && (a.cfg.res_flags[r] & B200S_NRT_RES_HOST_LEVEL)) ? INT64_MAX : INT64_MIN;
#pragma unroll
        for (int z = 0; z < Z; ++z)
          if (!((zs.zmask[z] >> r) & 1u)) zs.avail[z][r] = none;
      }
    }
    LvrbNode lcpu, lmem;
    uint32_t lfl = 0;
    if (use_lvrb && in) {
      lcpu = lvrb_node(a.lvrb_f64[n], a.lvrb_f64[(size_t)a.Npad + n], (double)a.lvrb_i64[n], a.margin, a.sens);
      double mcap = (double)a.lvrb_i64[(size_t)a.Npad + n];
      mcap *= 1. / 1024. / 1024.;  // resourcestats.go:62-63
      lmem = lvrb_node(a.lvrb_f64[2 * (size_t)a.Npad + n], a.lvrb_f64[3 * (size_t)a.Npad + n], mcap, a.margin, a.sens);
      lfl = a.lvrb_flags[n];
    }
    const int rg = (use_net && in) ? a.region[n] : 0, zn = (use_net && in) ? a.zone[n] : 0;
    const int64_t araw = (use_alloc && in) ? a.alloc_raw[n] : 0;
    for (int p = 0; p < P; ++p) {
      bool feasible = in;
      if (a.upstream) feasible = feasible && ((a.upstream[(size_t)p * a.words + (n >> 6)] >> (n & 63)) & 1ull);
      uint32_t s_nrt = 0, s_tlp = 0, s_lvrb = 0;
      int64_t cost = 0;
      if (use_nrt && in) {  // chained: NodeResourceTopologyMatch first (its Score only where it is still feasible)
        const int reason = nrt_filter<Z, R>(zs, nflags, node_res_mask, a.cfg, sp[p]);
        feasible = feasible && reason == 0;
        if (feasible) s_nrt = (uint32_t)nrt_score<Z, R, SC>(zs, cost_dummy, nflags, 8, a.cfg, sp[p]);
      }
      if (use_net && in && !a.equal[p]) {
        int64_t sat, viol;
        eval_node(a.topo, a.node_off + n, rg, zn, a.deps + a.dep_off[p], a.dep_off[p + 1] - a.dep_off[p], sat, viol, cost);
        feasible = feasible && !(viol > sat);  // networkoverhead.go:349-357
      }
      if (use_tlp && in)
        s_tlp = (uint32_t)tlp_score(a.tlp_util[n], a.tlp_cap[n], a.tlp_missing[n], a.tlp_flags[n], (double)a.tlp_pod[p],
                                    (double)a.tlp_target);
      if (use_lvrb && in) {
        const bool cpu_ok = lfl & B200S_LVRB_CPU_OK, mem_ok = lfl & B200S_LVRB_MEM_OK;
        const double rc = go_max((double)a.lvrb_req_cpu[p], 0), rm = go_max((double)a.lvrb_req_mem[p] * (1. / 1024. / 1024.), 0);
        const double cs = cpu_ok ? lvrb_res_score(lcpu, rc) : 0.0, ms = mem_ok ? lvrb_res_score(lmem, rm) : 0.0;
        const double total = (mem_ok && cpu_ok) ? go_min(ms, cs) : go_max(ms, cs);
        s_lvrb = (lfl & B200S_LVRB_HAS_METRICS) ? (uint32_t)go_f2i(go_round(total)) : 0u;
      }
      if (n < a.Npad) {
        a.part[(size_t)p * a.Npad + n] = ((feasible ? 1u : 0u) << 24) | ((s_nrt & 255u) << 16) | ((s_lvrb & 255u) << 8) | (s_tlp & 255u);
        if (use_net) a.cost[(size_t)p * a.Npad + n] = cost;
      }
      // min / max over the feasible nodes: warp shuffle, then one shared-memory atomic per warp
      unsigned long long v[4] = {feasible ? ~ord(araw) : 0ull, feasible ? ord(araw) : 0ull, feasible ? ~ord(cost) : 0ull,
                                 feasible ? ord(cost) : 0ull};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        for (int o = 16; o; o >>= 1) v[j] = max(v[j], __shfl_xor_sync(0xffffffffu, v[j], o));
        if ((t & 31) == 0 && v[j]) atomicMax(&s_red[p][j], v[j]);
      }
    }
  }
  __syncthreads();
  // most CTAs do not move the extrema: read first, so that the four cells do not serialise ~400 atomics each
  if (t < P * 4 && s_red[t / 4][t % 4] > *reinterpret_cast<volatile unsigned long long*>(&a.lohi[t]))
    atomicMax(&a.lohi[t], s_red[t / 4][t % 4]);
}

// the last CTA of phase 2 folds the per-block winners (k rounds of arg-max, strictly after the previous winner) and
// resets the min/max cells for the next cycle
__device__ void cycle_fold(const CycleArgs& a, int64_t* ss, int32_t* sn, int64_t& win_s, int32_t& win_n) {
  const int t = threadIdx.x, P = a.P;
  for (int p = 0; p < P; ++p) {
    int64_t last_s = 0;
    int32_t last_n = -1;
    bool have_last = false;
    const b200s_topk_entry* cand = a.block_best + (size_t)p * a.blocks2 * a.k;
    const int ncand = a.blocks2 * a.k;
    for (int round = 0; round < a.k; ++round) {
      int64_t s = INT64_MIN;
      int32_t nn = INT32_MAX;
      for (int i = t; i < ncand; i += 256) {
        const b200s_topk_entry e = cand[i];
        if (e.node < 0) continue;
        if (have_last && !better(last_s, last_n, e.score, e.node)) continue;
        if (better(e.score, e.node, s, nn)) {
          s = e.score;
          nn = e.node;
        }
      }
      for (int o = 16; o; o >>= 1) {
        const int64_t os = __shfl_xor_sync(0xffffffffu, s, o);
        const int32_t on = __shfl_xor_sync(0xffffffffu, nn, o);
        if (better(os, on, s, nn)) {
          s = os;
          nn = on;
        }
      }
      if ((t & 31) == 0) {
        ss[t >> 5] = s;
        sn[t >> 5] = nn;
      }
      __syncthreads();
      if (t == 0) {
        for (int i = 1; i < 8; ++i)
          if (better(ss[i], sn[i], s, nn)) {
            s = ss[i];
            nn = sn[i];
          }
        win_s = s;
        win_n = nn;
        b200s_topk_entry e;
        e.score = nn == INT32_MAX ? 0 : s;
        e.node = nn == INT32_MAX ? -1 : nn;
        e.pad = 0;
        a.out[(size_t)p * a.k + round] = e;
      }
      __syncthreads();
      last_s = win_s;
      last_n = win_n;
      have_last = win_n != INT32_MAX;
      if (!have_last) {  // nothing left: the remaining ranks are empty
        if (t == 0)
          for (int j = round + 1; j < a.k; ++j) {
            b200s_topk_entry e;
            e.score = 0;
            e.node = -1;
            e.pad = 0;
            a.out[(size_t)p * a.k + j] = e;
          }
        break;
      }
      __syncthreads();
    }
    __syncthreads();
  }
  if (t < PMAX * 4) a.lohi[t] = 0;
  if (t == 0) *a.done = 0;
}

// phase 2: NormalizeScore, weighted sum, per-block top-k, the final feasibility words (natural node order)
__global__ void __launch_bounds__(256) cycle_phase2_kernel(CycleArgs a) {
  __shared__ int64_t ss[8];
  __shared__ int32_t sn[8];
  __shared__ int64_t win_s;
  __shared__ int32_t win_n;
  const int t = threadIdx.x, P = a.P;
  const bool use_alloc = a.mask & (1u << B200S_PLUGIN_ALLOCATABLE), use_tlp = a.mask & (1u << B200S_PLUGIN_TLP);
  const bool use_lvrb = a.mask & (1u << B200S_PLUGIN_LVRB), use_nrt = a.mask & (1u << B200S_PLUGIN_NRT);
  const bool use_net = a.mask & (1u << B200S_PLUGIN_NETWORK_OVERHEAD);
  for (int p = 0; p < P; ++p) {
    const unsigned long long ulo_a = a.lohi[p * 4 + 0], uhi_a = a.lohi[p * 4 + 1], ulo_n = a.lohi[p * 4 + 2], uhi_n = a.lohi[p * 4 + 3];
    const bool any = uhi_a != 0 || ulo_a != 0;  // at least one feasible node (ord() never maps to 0 AND ~0 at once)
    const int64_t lo_a = unord(~ulo_a), hi_a = unord(uhi_a), lo_n = unord(~ulo_n), hi_n = unord(uhi_n);
    int64_t bs[KMAX];
    int32_t bn[KMAX];
    for (int i = 0; i < a.k; ++i) {
      bs[i] = INT64_MIN;
      bn[i] = INT32_MAX;
    }
    {
      const int n = blockIdx.x * 256 + t;  // the grid covers Npad in 256-node CTAs (Npad is a multiple of 128)
      const uint32_t pr = n < a.Npad ? a.part[(size_t)p * a.Npad + n] : 0u;
      const bool feasible = n < a.N && (pr >> 24);
      const uint32_t fbits = __ballot_sync(0xffffffffu, feasible);
      if ((t & 31) == 0 && n < a.Npad) a.feas32[((size_t)p * a.Npad + n) >> 5] = fbits;
      if (feasible) {
      int64_t total = 0;
      if (use_alloc) total = wrap_add(total, wrap_mul(a.w[B200S_PLUGIN_ALLOCATABLE], alloc_norm(a.alloc_raw[n], lo_a, hi_a, any)));
      if (use_tlp) total = wrap_add(total, wrap_mul(a.w[B200S_PLUGIN_TLP], (int64_t)(pr & 255u)));
      if (use_lvrb) total = wrap_add(total, wrap_mul(a.w[B200S_PLUGIN_LVRB], (int64_t)((pr >> 8) & 255u)));
      if (use_nrt) total = wrap_add(total, wrap_mul(a.w[B200S_PLUGIN_NRT], (int64_t)((pr >> 16) & 255u)));
      if (use_net)
        total = wrap_add(total, wrap_mul(a.w[B200S_PLUGIN_NETWORK_OVERHEAD], netoh_norm(a.cost[(size_t)p * a.Npad + n], lo_n, hi_n, any)));
      const int32_t g = a.node_off + n;
      if (better(total, g, bs[a.k - 1], bn[a.k - 1])) {  // insertion into the sorted per-thread list
        bs[a.k - 1] = total;
        bn[a.k - 1] = g;
        for (int i = a.k - 1; i > 0 && better(bs[i], bn[i], bs[i - 1], bn[i - 1]); --i) {
          const int64_t ts = bs[i]; bs[i] = bs[i - 1]; bs[i - 1] = ts;
          const int32_t tn = bn[i]; bn[i] = bn[i - 1]; bn[i - 1] = tn;
        }
      }
      }
    }
    for (int round = 0; round < a.k; ++round) {  // k rounds of block arg-max over the heads of the per-thread lists
      int64_t s = bs[0];
      int32_t nn = bn[0];
      for (int o = 16; o; o >>= 1) {
        const int64_t os = __shfl_xor_sync(0xffffffffu, s, o);
        const int32_t on = __shfl_xor_sync(0xffffffffu, nn, o);
        if (better(os, on, s, nn)) {
          s = os;
          nn = on;
        }
      }
      if ((t & 31) == 0) {
        ss[t >> 5] = s;
        sn[t >> 5] = nn;
      }
      __syncthreads();
      if (t == 0) {
        for (int i = 1; i < 8; ++i)
          if (better(ss[i], sn[i], s, nn)) {
            s = ss[i];
            nn = sn[i];
          }
        win_s = s;
        win_n = nn;
        b200s_topk_entry e;
        e.score = nn == INT32_MAX ? 0 : s;
        e.node = nn == INT32_MAX ? -1 : nn;
        e.pad = 0;
        a.block_best[((size_t)p * a.blocks2 + blockIdx.x) * a.k + round] = e;
      }
      __syncthreads();
      if (bn[0] == win_n && win_n != INT32_MAX) {  // the owner pops its head
        for (int i = 0; i < a.k - 1; ++i) {
          bs[i] = bs[i + 1];
          bn[i] = bn[i + 1];
        }
        bs[a.k - 1] = INT64_MIN;
        bn[a.k - 1] = INT32_MAX;
      }
      __syncthreads();
    }
  }
  // the CTA that finishes last folds (threadfence + ticket: its reads see every CTA's winners)
  __shared__ bool last;
  __threadfence();
  if (t == 0) last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  cycle_fold(a, ss, sn, win_s, win_n);
}

}  // namespace

// Whether the fused cycle applies to (snapshot, batch, args); see the header comment.
bool cycle_applies(b200s_ctx* c, uint32_t mask, int k, int write_total, bool any_p) {
  if (!c->fused_cycle || write_total || c->P < 1 || (!any_p && c->P > PMAX) || k < 1 || k > KMAX || comm_world(c) > 1) return false;
  if (mask & ~0x1Fu) return false;  // Peaks / LowRiskOverCommitment keep the plugin-by-plugin path
  if (mask & (1u << B200S_PLUGIN_NRT)) {
    if (c->nrt_strategy == B200S_NRT_LEAST_NUMA_NODES || c->nrt_Z > Z || c->nrt_R > R) return false;
  }
  if ((mask & (1u << B200S_PLUGIN_NETWORK_OVERHEAD)) && !c->netoh_apply_filter) return false;
  return true;
}

// The two launches of pods [p0, p0 + np) of the uploaded batch: arguments (winners to out[np][k]), kernel variant and
// grid.  May queue snapshot-time work (Allocatable's raw score / sort, the first zeroing of the min/max cells) on the
// engine stream; the caller launches after it.
static int cycle_build(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, int p0, int np, b200s_topk_entry* out,
                       CycleArgs& a, int& sc, int& blocks) {
  const int P = np, Npad = c->Npad;
  sc = (mask & (1u << B200S_PLUGIN_NRT)) && c->nrt_strategy == B200S_NRT_BALANCED_ALLOCATION ? 1 : 0;
  blocks = (Npad + 255) / 256;  // phase 2 CTAs = rows of the per-block winners
  B200S_CUDA_TRY(c, c->cycle_scratch.ensure((size_t)256 + (size_t)P * Npad * 12 + (size_t)P * blocks * k * sizeof(b200s_topk_entry) + 64));
  char* base = c->cycle_scratch.as<char>();
  memset(&a, 0, sizeof(a));
  a.mask = mask;
  a.P = P;
  a.N = c->N;
  a.Npad = Npad;
  a.words = Npad / 64;
  a.node_off = c->node_off;
  a.k = k;
  for (int j = 0; j < B200S_PLUGIN_COUNT; ++j) a.w[j] = weights[j];
  a.upstream = c->has_feasible ? c->feasible_in.as<uint64_t>() + (size_t)p0 * (Npad / 64) : nullptr;
  a.lohi = reinterpret_cast<unsigned long long*>(base);
  a.done = reinterpret_cast<unsigned int*>(base + PMAX * 4 * 8);
  a.cost = reinterpret_cast<int64_t*>(base + 256);
  a.part = reinterpret_cast<uint32_t*>(base + 256 + (size_t)P * Npad * 8);
  a.block_best = reinterpret_cast<b200s_topk_entry*>(base + 256 + (size_t)P * Npad * 12);
  a.out = out;
  B200S_CUDA_TRY(c, c->total_feas.ensure((size_t)c->P * (Npad / 64) * 8));
  a.feas32 = c->total_feas.as<uint32_t>() + (size_t)p0 * (Npad / 32);
  if (mask & (1u << B200S_PLUGIN_ALLOCATABLE)) {
    if (!c->has_alloc || !c->alloc_cfg) return c->set_err(B200S_ERR_STATE, "NodeResourcesAllocatable: snapshot columns / args missing");
    B200S_TRY(alloc_prepare(c));
    a.alloc_raw = c->alloc_raw.as<int64_t>();
  }
  if (mask & (1u << B200S_PLUGIN_TLP)) {
    if (!c->has_tlp || !c->has_tlp_pods) return c->set_err(B200S_ERR_STATE, "TargetLoadPacking: snapshot / pod columns missing");
    a.tlp_util = c->tlp_util.as<double>();
    a.tlp_cap = c->tlp_cap.as<int64_t>();
    a.tlp_missing = c->tlp_missing.as<int64_t>();
    a.tlp_flags = c->tlp_flags.as<uint8_t>();
    a.tlp_pod = c->tlp_pod_cpu.as<int64_t>() + p0;
    a.tlp_target = c->tlp_target;
  }
  if (mask & (1u << B200S_PLUGIN_LVRB)) {
    if (!c->has_lvrb || !c->has_lvrb_pods) return c->set_err(B200S_ERR_STATE, "LoadVariationRiskBalancing: snapshot / pod columns missing");
    a.lvrb_f64 = c->lvrb_f64.as<double>();
    a.lvrb_i64 = c->lvrb_i64.as<int64_t>();
    a.lvrb_flags = c->lvrb_flags.as<uint8_t>();
    a.lvrb_req_cpu = c->lvrb_req_cpu.as<int64_t>() + p0;
    a.lvrb_req_mem = c->lvrb_req_mem.as<int64_t>() + p0;
    a.margin = c->lvrb_margin;
    a.sens = c->lvrb_sens;
  }
  if (mask & (1u << B200S_PLUGIN_NRT)) {
    if (!c->has_nrt || !c->has_nrt_pods) return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: snapshot / pod columns missing");
    a.nrt_node_flags = c->nrt_node_flags.as<uint8_t>();
    a.nrt_nz = c->nrt_nz.as<uint8_t>();
    a.nrt_node_res_mask = c->nrt_node_res_mask.as<uint8_t>();
    a.nrt_zone_res_mask = c->nrt_zone_res_mask.as<uint8_t>();
    a.nrt_avail = c->nrt_avail.as<int64_t>();
    a.nrt_Zs = c->nrt_Z;
    a.nrt_Rs = c->nrt_R;
    a.pod_qos = c->nrt_pod_qos.as<uint8_t>() + p0;
    a.pod_flags = c->nrt_pod_flags.as<uint8_t>() + p0;
    a.pod_ninit = c->nrt_pod_ninit.as<uint8_t>() + p0;
    a.pod_napp = c->nrt_pod_napp.as<uint8_t>() + p0;
    a.pod_kind = c->nrt_pod_kind.as<uint8_t>() + (size_t)p0 * C_MAX;
    a.pod_req_mask = c->nrt_pod_req_mask.as<uint8_t>() + (size_t)p0 * (C_MAX + 1);
    a.pod_req = c->nrt_pod_req.as<int64_t>() + (size_t)p0 * (C_MAX + 1) * c->nrt_R;
    a.cfg.strategy = c->nrt_strategy;
    for (int r = 0; r < B200S_NRT_MAX_RES; ++r) {
      a.cfg.w[r] = c->nrt_w[r];
      a.cfg.res_flags[r] = c->nrt_res_flags[r];
    }
  }
  if (mask & (1u << B200S_PLUGIN_NETWORK_OVERHEAD)) {
    if (!c->has_netoh || !c->has_netoh_pods) return c->set_err(B200S_ERR_STATE, "NetworkOverhead: snapshot / pod columns missing");
    a.topo.zc = c->netoh_zone_cost.as<int64_t>();
    a.topo.rc = c->netoh_region_cost.as<int64_t>();
    a.topo.K = c->netoh_K;
    a.region = c->netoh_region.as<uint16_t>();
    a.zone = c->netoh_zone.as<uint16_t>();
    a.equal = c->netoh_equal.as<uint8_t>() + p0;
    a.dep_off = c->netoh_dep_off.as<int32_t>() + p0;
    a.deps = c->netoh_deps.as<b200s_netoh_dep>();
  }
  if (!c->cycle_cells_zero || c->cycle_scratch.p != c->cycle_cells_base) {  // first use / reallocation: the fold resets them afterwards
    B200S_CUDA_TRY(c, cudaMemsetAsync(base, 0, 256, c->stream));
    c->cycle_cells_zero = true;
    c->cycle_cells_base = c->cycle_scratch.p;
  }
  a.blocks2 = blocks;
  a.perm = (mask & (1u << B200S_PLUGIN_NRT)) ? c->nrt_perm.as<int32_t>() : nullptr;
  return B200S_OK;
}

// pods [p0, p0 + np) of the uploaded batch; winners to out[np][k] (device)
static int cycle_launch(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, int p0, int np, b200s_topk_entry* out) {
  CycleArgs a;
  int sc, blocks;
  B200S_TRY(cycle_build(c, mask, weights, k, p0, np, out, a, sc, blocks));
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_COMBINE);
  if (sc)
    cycle_phase1_kernel<1><<<c->Npad / 128, 128, 0, c->stream>>>(a);
  else
    cycle_phase1_kernel<0><<<c->Npad / 128, 128, 0, c->stream>>>(a);
  cycle_phase2_kernel<<<blocks, 256, 0, c->stream>>>(a);
  c->launches += 2;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

// ---- b200s_schedule_batch's cycle as ONE graph launch ------------------------------------------------------------------
// A scheduling cycle of one pod is bound by the host's driver calls, not by the device: copy of the pod columns, two
// kernel launches, copy of the winners.  The three device-side operations are a CUDA graph that is instantiated once
// and re-launched; the pod columns always travel through the same pinned staging block into the same arena, so between
// two cycles normally nothing but the CONTENT of that block changes.  Whatever does change (the size of the copy when a
// pod has more NetworkOverhead dependencies, buffer addresses after a snapshot upload, weights, strategy) is patched
// into the instantiated graph with the ExecSetParams calls -- compared bytewise, so a stale parameter cannot survive.
// The winners are written by the folding CTA straight into the engine's pinned, device-mapped result page.
namespace {
struct CycleGraph {
  cudaGraph_t g = nullptr;
  cudaGraphExec_t x = nullptr;
  cudaGraphNode_t copy = nullptr, k1 = nullptr, k2 = nullptr;
  CycleArgs a;
  int sc = -1, grid1 = 0, grid2 = 0;
  void* cp_dst = nullptr;
  const void* cp_src = nullptr;
  size_t cp_bytes = 0;
};

void fill_kernel_params(cudaKernelNodeParams& kp, void** argv, int sc, int phase, int grid) {
  memset(&kp, 0, sizeof(kp));
  kp.func = phase == 1 ? (sc ? (void*)cycle_phase1_kernel<1> : (void*)cycle_phase1_kernel<0>) : (void*)cycle_phase2_kernel;
  kp.gridDim = dim3(grid);
  kp.blockDim = dim3(phase == 1 ? 128 : 256);
  kp.sharedMemBytes = 0;
  kp.kernelParams = argv;
  kp.extra = nullptr;
}
}  // namespace

void cycle_graph_free(b200s_ctx* c) {
  CycleGraph* G = static_cast<CycleGraph*>(c->cycle_graph);
  if (!G) return;
  if (G->x) cudaGraphExecDestroy(G->x);
  if (G->g) cudaGraphDestroy(G->g);
  delete G;
  c->cycle_graph = nullptr;
}

// The held copy of the pod columns (c->held_*), both launches, winners to host_out (pinned + mapped, [P][k]).
int cycle_graph_run(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, b200s_topk_entry* host_out) {
  b200s_topk_entry* dev_out = nullptr;
  B200S_CUDA_TRY(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&dev_out), host_out, 0));
  CycleArgs a;
  int sc, blocks;
  B200S_TRY(cycle_build(c, mask, weights, k, 0, c->P, dev_out, a, sc, blocks));
  const int grid1 = c->Npad / 128;
  CycleGraph* G = static_cast<CycleGraph*>(c->cycle_graph);
  void* argv[1] = {&a};
  cudaKernelNodeParams kp;
  if (G && G->sc != sc) {  // another kernel function: rebuild (a strategy change, once)
    cycle_graph_free(c);
    G = nullptr;
  }
  if (!G) {
    G = new CycleGraph();
    c->cycle_graph = G;
    cudaError_t e = cudaGraphCreate(&G->g, 0);
    if (e == cudaSuccess)
      e = cudaGraphAddMemcpyNode1D(&G->copy, G->g, nullptr, 0, c->held_dst, c->held_src, c->held_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      fill_kernel_params(kp, argv, sc, 1, grid1);
      e = cudaGraphAddKernelNode(&G->k1, G->g, &G->copy, 1, &kp);
    }
    if (e == cudaSuccess) {
      fill_kernel_params(kp, argv, sc, 2, blocks);
      e = cudaGraphAddKernelNode(&G->k2, G->g, &G->k1, 1, &kp);
    }
    if (e == cudaSuccess) e = cudaGraphInstantiate(&G->x, G->g, 0);
    if (e != cudaSuccess) {
      cycle_graph_free(c);
      return c->set_err(B200S_ERR_CUDA, std::string("schedule_batch: building the cycle graph: ") + cudaGetErrorString(e));
    }
    G->a = a;
    G->sc = sc;
    G->grid1 = grid1;
    G->grid2 = blocks;
    G->cp_dst = c->held_dst;
    G->cp_src = c->held_src;
    G->cp_bytes = c->held_bytes;
  } else {
    cudaError_t e = cudaSuccess;
    if (G->cp_dst != c->held_dst || G->cp_src != c->held_src || G->cp_bytes != c->held_bytes) {
      e = cudaGraphExecMemcpyNodeSetParams1D(G->x, G->copy, c->held_dst, c->held_src, c->held_bytes, cudaMemcpyHostToDevice);
      G->cp_dst = c->held_dst;
      G->cp_src = c->held_src;
      G->cp_bytes = c->held_bytes;
    }
    if (e == cudaSuccess && (memcmp(&G->a, &a, sizeof(a)) != 0 || G->grid1 != grid1 || G->grid2 != blocks)) {
      fill_kernel_params(kp, argv, sc, 1, grid1);
      e = cudaGraphExecKernelNodeSetParams(G->x, G->k1, &kp);
      if (e == cudaSuccess) {
        fill_kernel_params(kp, argv, sc, 2, blocks);
        e = cudaGraphExecKernelNodeSetParams(G->x, G->k2, &kp);
      }
      G->a = a;
      G->grid1 = grid1;
      G->grid2 = blocks;
    }
    if (e != cudaSuccess) {
      cycle_graph_free(c);
      return c->set_err(B200S_ERR_CUDA, std::string("schedule_batch: updating the cycle graph: ") + cudaGetErrorString(e));
    }
  }
  c->held_bytes = 0;  // the graph's first node is the copy
  B200S_CUDA_TRY(c, cudaGraphLaunch(G->x, c->stream));
  c->launches += 2;
  c->topk_valid = c->total_valid = false;  // the winners went to the caller, not to the engine's top-k buffer
  c->feas_valid = true;                    // ... the final feasible set is resident
  return B200S_OK;
}

int cycle_eval(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k) {
  B200S_CUDA_TRY(c, c->topk_final.ensure((size_t)c->P * k * sizeof(b200s_topk_entry)));
  B200S_TRY(cycle_launch(c, mask, weights, k, 0, c->P, c->topk_final.as<b200s_topk_entry>()));
  c->topk_k = k;
  c->topk_valid = true;
  c->total_valid = false;
  return B200S_OK;
}

namespace {
// "Assume" of the pod that was just placed, on the device: what the scheduler's caches do between two cycles.
//   NodeResourceTopologyMatch (OverReserve cache): ReserveNodeResources records the pod's effective request
//     (cache/store.go:101-112, pkg/util/resource.go:51-85) and GetCachedNRTCopy -> UpdateNRT takes it off EVERY zone of
//     the node that lists the resource: available < q ? 0 : available - q (store.go:129-160)
//   TargetLoadPacking: the bind handler adds the pod to ScheduledPodsCache (handler.go:131-167), so the node's
//     "missing" utilisation grows by the pod's predicted CPU (targetloadpacking.go:151-167)
// One CTA, thread (z, r) of the winner node.
__global__ void assume_kernel(const b200s_topk_entry* __restrict__ winner, int node_off, int N, size_t npad, int Zs, int Rs,
                              const uint8_t* __restrict__ zmask, int64_t* __restrict__ avail, const int64_t* __restrict__ eff_req,
                              uint32_t eff_mask_ptr_valid, const uint8_t* __restrict__ eff_mask, int64_t* __restrict__ tlp_missing,
                              const int64_t* __restrict__ tlp_pod) {
  const int n = winner->node - node_off;
  if (winner->node < 0 || n < 0 || n >= N) return;  // unschedulable pod, or placed on another shard
  const int t = threadIdx.x;
  if (avail && t < Zs * Rs) {
    const int z = t / Rs, r = t % Rs;
    if (((zmask[(size_t)z * npad + n] >> r) & 1u) && ((eff_mask[0] >> r) & 1u)) {
      int64_t* cell = avail + ((size_t)z * Rs + r) * npad + n;
      const int64_t q = eff_req[r], v = *cell;
      *cell = v < q ? 0 : v - q;
    }
  }
  if (tlp_missing && t == 0) tlp_missing[n] = wrap_add(tlp_missing[n], tlp_pod[0]);
  (void)eff_mask_ptr_valid;
}
}  // namespace

// Speculative placement of a whole batch, pod by pod, without leaving the device: cycle of pod i on the snapshot as
// pods 0..i-1 left it, winner, assume, next.  The snapshot columns are modified IN PLACE (the caller resyncs rows
// whose bind failed with b200s_snapshot_patch_*).  winners: [P] (k = 1), device.
int cycle_sequence(b200s_ctx* c, uint32_t mask, const int64_t* weights, b200s_topk_entry* winners) {
  const bool nrt = mask & (1u << B200S_PLUGIN_NRT), tlp = mask & (1u << B200S_PLUGIN_TLP);
  for (int p = 0; p < c->P; ++p) {
    B200S_TRY(cycle_launch(c, mask, weights, 1, p, 1, winners + p));
    if (!nrt && !tlp) continue;
    assume_kernel<<<1, 64, 0, c->stream>>>(
        winners + p, c->node_off, c->N, (size_t)c->Npad, c->nrt_Z, c->nrt_R, nrt ? c->nrt_zone_res_mask.as<uint8_t>() : nullptr,
        nrt ? c->nrt_avail.as<int64_t>() : nullptr,
        nrt ? c->nrt_pod_req.as<int64_t>() + ((size_t)p * (C_MAX + 1) + C_MAX) * c->nrt_R : nullptr, 1u,
        nrt ? c->nrt_pod_req_mask.as<uint8_t>() + (size_t)p * (C_MAX + 1) + C_MAX : nullptr,
        tlp ? c->tlp_missing.as<int64_t>() : nullptr, tlp ? c->tlp_pod_cpu.as<int64_t>() + p : nullptr);
    c->launches++;
  }
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

}  // namespace b200s

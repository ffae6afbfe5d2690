// NetworkOverhead: PreFilter node loop + Filter + Score + NormalizeScore, all pods x all nodes.
//
// Reference semantics (pkg/networkaware/networkoverhead/networkoverhead.go):
//   per node: (satisfied, violated) = checkMaxNetworkCostRequirements :500-573
//             cost                  = getAccumulatedCost              :576-638
//   Filter  : !scoreEqually && violated > satisfied -> Unschedulable   :326-359
//   Score   : scoreEqually ? 0 : cost                                  :362-386
//   Normalize over the feasible list: 100 - int64(100.0*float64(s-min)/float64(max-min)) :389-418
// The reference keeps the per-node cost map in a Go map keyed by (origin,destination) strings and
// rebuilds (and re-sorts) it for every node; here region/zone labels are dictionary ids and the
// two cost lists are dense [K][K] int64 matrices with a MISSING sentinel (L2/L1 resident).
//
// Kernels: (1) raw pass: cost + own filter -> raw int64 matrix, feasibility words, reason codes;
//          (2) per-pod min/max over the feasible bits; [NCCL min/max all-reduce when sharded]
//          (3) normalise pass -> score matrix (int64 or u8).
// Warp layout of the P x N passes: a warp covers 64 consecutive nodes, lane l owns nodes l and
// l+32, so the warp's two ballots ARE the 64-bit feasibility word and every store is a fully
// used 256 B segment.
#include "engine.h"
#include "netoh_device.cuh"

namespace b200s {

namespace {

using namespace netdev;

template <int PT>
__global__ void __launch_bounds__(256)
netoh_raw_kernel(Topo t, const uint16_t* __restrict__ region, const uint16_t* __restrict__ zone, int node_off,
                 const uint8_t* __restrict__ equal, const int32_t* __restrict__ dep_off,
                 const b200s_netoh_dep* __restrict__ deps, const uint64_t* __restrict__ upstream, int words, int N,
                 int Npad, int P, int64_t* __restrict__ raw, uint64_t* __restrict__ feas_out,
                 uint8_t* __restrict__ reasons, uint32_t* __restrict__ counts, bool apply_filter) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbase = (blockIdx.x * 8 + warp) * 64;  // this warp's 64-node group
  if (nbase >= Npad) return;
  const int p0 = blockIdx.y * PT;
  const int pend = min(p0 + PT, P);
  const int n0 = nbase + lane, n1 = nbase + lane + 32;
  const int r0 = region[n0], z0 = zone[n0], r1 = region[n1], z1 = zone[n1];
  const bool v0 = n0 < N, v1 = n1 < N;
  const int word = nbase >> 6;
  for (int p = p0; p < pend; ++p) {
    const bool eq = equal[p] != 0;
    const int d0 = dep_off[p], nd = dep_off[p + 1] - d0;
    int64_t s0 = 0, w0 = 0, c0 = 0, s1 = 0, w1 = 0, c1 = 0;
    if (!eq) {
      eval_node(t, node_off + n0, r0, z0, deps + d0, nd, s0, w0, c0);
      eval_node(t, node_off + n1, r1, z1, deps + d0, nd, s1, w1, c1);
    }
    const bool pass0 = v0 && (eq || !(w0 > s0)), pass1 = v1 && (eq || !(w1 > s1));
    uint64_t up = upstream ? upstream[(size_t)p * words + word] : ~0ull;
    const bool f0 = (apply_filter ? pass0 : v0) && ((up >> lane) & 1ull);
    const bool f1 = (apply_filter ? pass1 : v1) && ((up >> (lane + 32)) & 1ull);
    const uint64_t fw = (uint64_t)__ballot_sync(0xffffffffu, f0) | ((uint64_t)__ballot_sync(0xffffffffu, f1) << 32);
    if (lane == 0) feas_out[(size_t)p * words + word] = fw;
    // raw = PreFilterState.finalCostMap (written for every node, filtered or not: Score reads it, :383)
    int64_t* row = raw + (size_t)p * Npad;
    row[n0] = (v0 && !eq) ? c0 : 0;
    row[n1] = (v1 && !eq) ? c1 : 0;
    if (counts) {  // satisfiedMap / violatedMap for the Filter message (:355-356); counts <= #deps
      uint32_t* cr = counts + (size_t)p * Npad;
      cr[n0] = (uint32_t)(s0 & 0xffff) | ((uint32_t)(w0 & 0xffff) << 16);
      cr[n1] = (uint32_t)(s1 & 0xffff) | ((uint32_t)(w1 & 0xffff) << 16);
    }
    if (reasons) {
      uint8_t* rr = reasons + (size_t)p * Npad;
      rr[n0] = !v0 ? 0 : (!pass0 ? B200S_REASON_NETOH_VIOLATED : (f0 ? B200S_REASON_OK : B200S_REASON_UPSTREAM));
      rr[n1] = !v1 ? 0 : (!pass1 ? B200S_REASON_NETOH_VIOLATED : (f1 ? B200S_REASON_OK : B200S_REASON_UPSTREAM));
    }
  }
}

// One CTA per pod: min/max of raw over the feasible bits (getMinMaxScores :421-435).
__global__ void __launch_bounds__(256)
row_minmax_kernel(const int64_t* __restrict__ raw, const uint64_t* __restrict__ feas, int words, int Npad,
                  int64_t* __restrict__ lo, int64_t* __restrict__ hi) {
  const int p = blockIdx.x;
  const int64_t* row = raw + (size_t)p * Npad;
  const uint64_t* fr = feas + (size_t)p * words;
  int64_t mn = INT64_MAX, mx = INT64_MIN;
  for (int n = threadIdx.x; n < Npad; n += 256) {
    if ((fr[n >> 6] >> (n & 63)) & 1ull) {
      int64_t v = row[n];
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
  }
  for (int o = 16; o; o >>= 1) {
    int64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
  }
  __shared__ int64_t smn[8], smx[8];
  if ((threadIdx.x & 31) == 0) {
    smn[threadIdx.x >> 5] = mn;
    smx[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      mn = smn[i] < mn ? smn[i] : mn;
      mx = smx[i] > mx ? smx[i] : mx;
    }
    lo[p] = mn;
    hi[p] = mx;
  }
}

// mode 0: all 0 (min == max == 0, or nothing feasible)       :400-402
// mode 1: 100 - floor((s-min)*100/range) with the 32-bit reciprocal (== the float64 formula: for
//         range*100 < 2^32 the fp64 quotient cannot round across an integer)
// mode 2: the float64 formula verbatim                        :406-410
// mode 3: max == min != 0 -> 100 - int64(float64(0)) = 100    :411-413
__global__ void netoh_params_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi, int P,
                                    NormParam* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  NormParam q;
  q.lo = lo[p];
  q.range = wrap_sub(hi[p], lo[p]);
  q.magic = q.shift = q.pad = 0;
  if (lo[p] > hi[p] || (lo[p] == 0 && hi[p] == 0)) {
    q.mode = 0;
  } else if (lo[p] == hi[p]) {
    q.mode = 3;
  } else if (q.range > 0 && q.range <= (int64_t)(0xffffffffu / 100u)) {
    uint32_t r = (uint32_t)q.range, s = 31 - __clz(r);
    uint64_t m = ((1ull << 32) << s) / r;
    q.magic = m > 0xffffffffull ? 0xffffffffu : (uint32_t)m;
    q.shift = s;
    q.mode = 1;
  } else {
    q.mode = 2;
  }
  out[p] = q;
}

__device__ __forceinline__ int64_t netoh_norm_one(const NormParam& np, int64_t s) {
  if (np.mode == 1) {
    uint32_t n100 = ((uint32_t)(uint64_t)s - (uint32_t)(uint64_t)np.lo) * 100u;
    uint32_t q0 = __umulhi(n100, np.magic) >> np.shift;
    uint32_t rem = n100 - q0 * (uint32_t)np.range;
    q0 += rem >= (uint32_t)np.range ? 1u : 0u;
    return 100 - (int64_t)q0;
  }
  if (np.mode == 0) return 0;
  if (np.mode == 3) return 100;
  double norm = 100.0 * (double)wrap_sub(s, np.lo) / (double)np.range;
  int64_t tr = (norm >= -9223372036854775808.0 && norm < 9223372036854775808.0) ? (int64_t)norm : INT64_MIN;
  return wrap_sub(100, tr);
}

template <class OutT, int PT>
__global__ void __launch_bounds__(256)
netoh_norm_kernel(const int64_t* __restrict__ raw, const uint64_t* __restrict__ feas, const NormParam* __restrict__ params,
                  int words, int Npad, int P, OutT* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbase = (blockIdx.x * 8 + warp) * 64;
  if (nbase >= Npad) return;
  const int p0 = blockIdx.y * PT, pend = min(p0 + PT, P);
  const int n0 = nbase + lane, n1 = n0 + 32, word = nbase >> 6;
  for (int p = p0; p < pend; ++p) {
    const NormParam np = params[p];
    const uint64_t fw = feas[(size_t)p * words + word];
    const int64_t* row = raw + (size_t)p * Npad;
    int64_t a = ((fw >> lane) & 1ull) ? netoh_norm_one(np, row[n0]) : 0;
    int64_t b = ((fw >> (lane + 32)) & 1ull) ? netoh_norm_one(np, row[n1]) : 0;
    OutT* orow = out + (size_t)p * Npad;
    orow[n0] = (OutT)a;
    orow[n1] = (OutT)b;
  }
}

// ---------------------------------------------------------------------------------------------
// Fast path (B200 design, not in the reference): (satisfied, violated, cost) of a (pod, node) depend on
// the node only through its (region, zone) label pair — except for the dependencies hosted on that very
// node.  So per pod the dependency loop runs once per distinct PAIR (tens) instead of once per NODE
// (hundreds of thousands); the P x N passes become a 16-byte table lookup, and the rare hosted nodes
// (a per-tile shared-memory bitmap marks them) are evaluated directly.  Two light passes (min/max, then
// normalise + store) replace the raw-matrix round trip: HBM traffic = the output matrix.
__global__ void netoh_pair_kernel(Topo t, const uint16_t* __restrict__ pair_r, const uint16_t* __restrict__ pair_z, int NQ,
                                  const uint8_t* __restrict__ equal, const int32_t* __restrict__ dep_off,
                                  const b200s_netoh_dep* __restrict__ deps, int P, int64_t* __restrict__ pair_cost,
                                  uint32_t* __restrict__ pair_sv) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)P * NQ) return;
  const int p = (int)(i / NQ), q = (int)(i % NQ);
  int64_t s = 0, w = 0, cst = 0;
  if (!equal[p]) eval_node(t, INT32_MIN /* no node: the same-host case is handled per node */, pair_r[q], pair_z[q], deps + dep_off[p], dep_off[p + 1] - dep_off[p], s, w, cst);
  pair_cost[i] = cst;
  pair_sv[i] = (uint32_t)s | ((uint32_t)w << 16);
}

__global__ void fill_lohi_kernel(int64_t* lo, int64_t* hi, int P) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) {
    lo[p] = INT64_MAX;  // getMinMaxScores :422-423
    hi[p] = INT64_MIN;
  }
}

// PASS 1: filter verdict, feasibility words, reason codes, per-pod min/max.  PASS 2: normalised scores.
template <int PASS, class OutT, int PT>
__global__ void __launch_bounds__(256)
netoh_fast_kernel(Topo t, const int32_t* __restrict__ pair_id, int NQ, const int64_t* __restrict__ pair_cost,
                  const uint32_t* __restrict__ pair_sv, int node_off, const uint8_t* __restrict__ equal,
                  const int32_t* __restrict__ dep_off, const b200s_netoh_dep* __restrict__ deps,
                  const uint16_t* __restrict__ region, const uint16_t* __restrict__ zone,
                  const uint64_t* __restrict__ upstream, int words, int N, int Npad, int P, bool apply_filter,
                  uint64_t* __restrict__ feas, uint8_t* __restrict__ reasons, int64_t* __restrict__ lo,
                  int64_t* __restrict__ hi, const NormParam* __restrict__ params, OutT* __restrict__ out) {
  __shared__ uint32_t bm[PT][16];  // hosted-node bitmap of this CTA's 512 nodes, per pod of the tile
  __shared__ long long s_lo[PT], s_hi[PT];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncta = blockIdx.x * 512;
  const int nbase = ncta + warp * 64;
  const int p0 = blockIdx.y * PT, pend = min(p0 + PT, P);
  for (int i = threadIdx.x; i < PT * 16; i += 256) bm[i / 16][i % 16] = 0;
  if (PASS == 1)
    for (int i = threadIdx.x; i < PT; i += 256) {
      s_lo[i] = INT64_MAX;
      s_hi[i] = INT64_MIN;
    }
  __syncthreads();
  for (int p = p0; p < pend; ++p) {
    if (equal[p]) continue;
    const int d0 = dep_off[p], d1 = dep_off[p + 1];
    for (int i = d0 + threadIdx.x; i < d1; i += 256) {
      const int local = deps[i].host_node - node_off - ncta;
      if (local >= 0 && local < 512) atomicOr(&bm[p - p0][local >> 5], 1u << (local & 31));
    }
  }
  __syncthreads();
  const bool active = nbase < Npad;
  const int n0 = nbase + lane, n1 = n0 + 32;
  int q0 = 0, q1 = 0;
  if (active) {
    q0 = pair_id[n0];
    q1 = pair_id[n1];
  }
  const bool v0 = active && n0 < N, v1 = active && n1 < N;
  const int word = nbase >> 6;
  for (int p = p0; p < pend && active; ++p) {
    const int pp = p - p0;
    const bool eq = equal[p] != 0;
    int64_t c0 = 0, c1 = 0, s0 = 0, w0 = 0, s1 = 0, w1 = 0;
    if (!eq) {
      const uint32_t h0 = bm[pp][warp * 2], h1 = bm[pp][warp * 2 + 1];
      if ((h0 >> lane) & 1u) {  // a dependency is hosted on this very node: evaluate directly
        eval_node(t, node_off + n0, region[n0], zone[n0], deps + dep_off[p], dep_off[p + 1] - dep_off[p], s0, w0, c0);
      } else {
        const size_t k = (size_t)p * NQ + q0;
        c0 = __ldg(&pair_cost[k]);
        const uint32_t sv = __ldg(&pair_sv[k]);
        s0 = sv & 0xffff;
        w0 = sv >> 16;
      }
      if ((h1 >> lane) & 1u) {
        eval_node(t, node_off + n1, region[n1], zone[n1], deps + dep_off[p], dep_off[p + 1] - dep_off[p], s1, w1, c1);
      } else {
        const size_t k = (size_t)p * NQ + q1;
        c1 = __ldg(&pair_cost[k]);
        const uint32_t sv = __ldg(&pair_sv[k]);
        s1 = sv & 0xffff;
        w1 = sv >> 16;
      }
    }
    if (PASS == 1) {
      const bool pass0 = v0 && (eq || !(w0 > s0)), pass1 = v1 && (eq || !(w1 > s1));
      const uint64_t up = upstream ? upstream[(size_t)p * words + word] : ~0ull;
      const bool f0 = (apply_filter ? pass0 : v0) && ((up >> lane) & 1ull);
      const bool f1 = (apply_filter ? pass1 : v1) && ((up >> (lane + 32)) & 1ull);
      const uint64_t fw = (uint64_t)__ballot_sync(0xffffffffu, f0) | ((uint64_t)__ballot_sync(0xffffffffu, f1) << 32);
      if (lane == 0) feas[(size_t)p * words + word] = fw;
      uint8_t* rr = reasons + (size_t)p * Npad;
      rr[n0] = !v0 ? 0 : (!pass0 ? B200S_REASON_NETOH_VIOLATED : (f0 ? B200S_REASON_OK : B200S_REASON_UPSTREAM));
      rr[n1] = !v1 ? 0 : (!pass1 ? B200S_REASON_NETOH_VIOLATED : (f1 ? B200S_REASON_OK : B200S_REASON_UPSTREAM));
      if (fw) {  // warp-uniform
        int64_t mn = INT64_MAX, mx = INT64_MIN;
        if (f0) mn = mx = c0;
        if (f1) {
          mn = c1 < mn ? c1 : mn;
          mx = c1 > mx ? c1 : mx;
        }
        for (int o = 16; o; o >>= 1) {
          const int64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
          mn = a < mn ? a : mn;
          mx = b > mx ? b : mx;
        }
        if (lane == 0) {
          atomicMin(&s_lo[pp], (long long)mn);
          atomicMax(&s_hi[pp], (long long)mx);
        }
      }
    } else {
      const NormParam np = params[p];
      const uint64_t fw = feas[(size_t)p * words + word];
      OutT* orow = out + (size_t)p * Npad;
      orow[n0] = (OutT)(((fw >> lane) & 1ull) ? netoh_norm_one(np, c0) : 0);
      orow[n1] = (OutT)(((fw >> (lane + 32)) & 1ull) ? netoh_norm_one(np, c1) : 0);
    }
  }
  if (PASS == 1) {
    __syncthreads();
    for (int i = threadIdx.x; i < pend - p0; i += 256) {
      if (s_lo[i] != INT64_MAX || s_hi[i] != INT64_MIN) {
        atomicMin(reinterpret_cast<long long*>(lo) + p0 + i, s_lo[i]);
        atomicMax(reinterpret_cast<long long*>(hi) + p0 + i, s_hi[i]);
      }
    }
  }
}

// The same two passes with the pods' pair tables staged in shared memory and FOUR consecutive nodes per thread: the
// table lookups leave L2, the per-eval byte stores become one 32-bit store per thread (reason codes, u8 scores) or two
// 16-byte stores (int64 scores), the feasibility words are assembled nibble by nibble.  PASS 2 stages the NORMALISED
// score of every pair instead of its cost (the normalisation depends on (pod, cost) only).  Needs PT x NQ x 12 bytes of
// shared memory; larger label-pair dictionaries keep netoh_fast_kernel.
template <int PASS, class OutT, int PT>
__global__ void __launch_bounds__(256)
netoh_fast4_kernel(Topo t, const int32_t* __restrict__ pair_id, int NQ, const int64_t* __restrict__ pair_cost,
                   const uint32_t* __restrict__ pair_sv, int node_off, const uint8_t* __restrict__ equal,
                   const int32_t* __restrict__ dep_off, const b200s_netoh_dep* __restrict__ deps,
                   const uint16_t* __restrict__ region, const uint16_t* __restrict__ zone,
                   const uint64_t* __restrict__ upstream, int words, int N, int Npad, int P, bool apply_filter,
                   uint64_t* __restrict__ feas, uint8_t* __restrict__ reasons, int64_t* __restrict__ lo,
                   int64_t* __restrict__ hi, const NormParam* __restrict__ params, OutT* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char dyn[];
  int64_t* s_val = reinterpret_cast<int64_t*>(dyn);                        // [PT][NQ] cost (PASS 1) / normalised score (PASS 2)
  uint32_t* s_sv = reinterpret_cast<uint32_t*>(dyn + (size_t)PT * NQ * 8);  // [PT][NQ] satisfied | violated << 16 (PASS 1)
  __shared__ uint32_t bm[PT][32];  // hosted-node bitmap of this CTA's 1024 nodes, per pod of the tile
  __shared__ long long s_lo[PT], s_hi[PT];
  const int tid = threadIdx.x, lane = tid & 31;
  const int ncta = blockIdx.x * 1024;
  const int p0 = blockIdx.y * PT, pend = min(p0 + PT, P);
  for (int i = tid; i < PT * 32; i += 256) bm[i / 32][i % 32] = 0;
  if (PASS == 1)
    for (int i = tid; i < PT; i += 256) {
      s_lo[i] = INT64_MAX;
      s_hi[i] = INT64_MIN;
    }
  for (int i = tid; i < (pend - p0) * NQ; i += 256) {
    const int pp = i / NQ;
    const size_t k = (size_t)p0 * NQ + i;
    const int64_t cost = pair_cost[k];
    if (PASS == 1) {
      s_val[i] = cost;
      s_sv[i] = pair_sv[k];
    } else {
      s_val[i] = netoh_norm_one(params[p0 + pp], cost);
    }
  }
  __syncthreads();
  for (int p = p0; p < pend; ++p) {
    if (equal[p]) continue;
    const int d0 = dep_off[p], d1 = dep_off[p + 1];
    for (int i = d0 + tid; i < d1; i += 256) {
      const int local = deps[i].host_node - node_off - ncta;
      if (local >= 0 && local < 1024) atomicOr(&bm[p - p0][local >> 5], 1u << (local & 31));
    }
  }
  __syncthreads();
  const int n4 = ncta + tid * 4;
  const bool active = n4 < Npad;  // Npad is a multiple of 128: uniform over the warp
  int q[4] = {0, 0, 0, 0};
  if (active) {
    const int4 qq = *reinterpret_cast<const int4*>(pair_id + n4);
    q[0] = qq.x, q[1] = qq.y, q[2] = qq.z, q[3] = qq.w;
  }
  uint8_t* feas8 = reinterpret_cast<uint8_t*>(feas);
  const size_t row_bytes = (size_t)words * 8;
  for (int p = p0; p < pend && active; ++p) {
    const int pp = p - p0;
    const bool eq = equal[p] != 0;
    const uint32_t hb = eq ? 0u : (bm[pp][tid >> 3] >> ((tid & 7) * 4)) & 15u;
    int64_t val[4];
    uint32_t sv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      val[e] = eq ? (PASS == 1 ? 0 : s_val[pp * NQ + q[e]]) : s_val[pp * NQ + q[e]];
      sv[e] = (PASS == 1 && !eq) ? s_sv[pp * NQ + q[e]] : 0u;
    }
    if (hb) {  // a dependency is hosted on one of these nodes: evaluate it directly
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!((hb >> e) & 1u) || n4 + e >= N) continue;
        int64_t s1 = 0, w1 = 0, c1 = 0;
        eval_node(t, node_off + n4 + e, region[n4 + e], zone[n4 + e], deps + dep_off[p], dep_off[p + 1] - dep_off[p], s1, w1, c1);
        val[e] = PASS == 1 ? c1 : netoh_norm_one(params[p], c1);
        sv[e] = (uint32_t)s1 | ((uint32_t)w1 << 16);
      }
    }
    if (PASS == 1) {
      const uint32_t up = upstream ? (uint32_t)(upstream[(size_t)p * words + (n4 >> 6)] >> (n4 & 63)) & 15u : 15u;
      uint32_t nib = 0, r4 = 0;
      int64_t mn = INT64_MAX, mx = INT64_MIN;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool v = n4 + e < N;
        const bool pass = v && (eq || !((sv[e] >> 16) > (sv[e] & 0xffffu)));  // networkoverhead.go:349-357
        const bool f = (apply_filter ? pass : v) && ((up >> e) & 1u);
        nib |= (f ? 1u : 0u) << e;
        const uint32_t code = !v ? 0u : (!pass ? (uint32_t)B200S_REASON_NETOH_VIOLATED : (f ? (uint32_t)B200S_REASON_OK : (uint32_t)B200S_REASON_UPSTREAM));
        r4 |= code << (8 * e);
        mn = f && val[e] < mn ? val[e] : mn;
        mx = f && val[e] > mx ? val[e] : mx;
      }
      const uint32_t other = __shfl_xor_sync(0xffffffffu, nib, 1);
      if (!(lane & 1)) feas8[(size_t)p * row_bytes + (n4 >> 3)] = (uint8_t)(nib | (other << 4));
      *reinterpret_cast<uint32_t*>(reasons + (size_t)p * Npad + n4) = r4;
      if (__any_sync(0xffffffffu, nib != 0)) {
        for (int o = 16; o; o >>= 1) {
          const int64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
          mn = a < mn ? a : mn;
          mx = b > mx ? b : mx;
        }
        if (lane == 0) {
          atomicMin(&s_lo[pp], (long long)mn);
          atomicMax(&s_hi[pp], (long long)mx);
        }
      }
    } else {
      const uint32_t fn = ((uint32_t)feas8[(size_t)p * row_bytes + (n4 >> 3)] >> (n4 & 4)) & 15u;
      int64_t o4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o4[e] = ((fn >> e) & 1u) ? val[e] : 0;
      OutT* orow = out + (size_t)p * Npad + n4;
      if constexpr (sizeof(OutT) == 1) {
        st_stream_u32(orow, ((uint32_t)o4[0] & 255u) | (((uint32_t)o4[1] & 255u) << 8) | (((uint32_t)o4[2] & 255u) << 16) |
                                (((uint32_t)o4[3] & 255u) << 24));
      } else {
        st_stream_v2(reinterpret_cast<int64_t*>(orow), o4[0], o4[1]);
        st_stream_v2(reinterpret_cast<int64_t*>(orow) + 2, o4[2], o4[3]);
      }
    }
  }
  if (PASS == 1) {
    __syncthreads();
    for (int i = tid; i < pend - p0; i += 256) {
      if (s_lo[i] != INT64_MAX || s_hi[i] != INT64_MIN) {
        atomicMin(reinterpret_cast<long long*>(lo) + p0 + i, s_lo[i]);
        atomicMax(reinterpret_cast<long long*>(hi) + p0 + i, s_hi[i]);
      }
    }
  }
}

int netoh_eval_fast(b200s_ctx* c, int dtype) {
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64, NQ = c->netoh_NQ;
  PluginOut& o = c->out[B200S_PLUGIN_NETWORK_OVERHEAD];
  B200S_CUDA_TRY(c, c->netoh_pair_cost.ensure((size_t)P * NQ * 8));
  B200S_CUDA_TRY(c, c->netoh_pair_sv.ensure((size_t)P * NQ * 4));
  B200S_CUDA_TRY(c, c->pod_lo.ensure((size_t)P * 16));
  int64_t* const lo_buf = c->pod_lo.as<int64_t>();
  int64_t* const hi_buf = lo_buf + P;
  B200S_CUDA_TRY(c, c->norm_params.ensure((size_t)P * sizeof(NormParam)));
  Topo t{c->netoh_zone_cost.as<int64_t>(), c->netoh_region_cost.as<int64_t>(), c->netoh_K};
  const size_t pairs = (size_t)P * NQ;
  netoh_pair_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(
      t, c->netoh_pair_r.as<uint16_t>(), c->netoh_pair_z.as<uint16_t>(), NQ, c->netoh_equal.as<uint8_t>(),
      c->netoh_dep_off.as<int32_t>(), c->netoh_deps.as<b200s_netoh_dep>(), P, c->netoh_pair_cost.as<int64_t>(),
      c->netoh_pair_sv.as<uint32_t>());
  fill_lohi_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(lo_buf, hi_buf, P);
  c->launches += 2;
  B200S_CUDA_TRY(c, cudaGetLastError());
  constexpr int PT = 32;
  // pair tables in shared memory + 4 nodes per thread when they fit (two CTAs per SM)
  const size_t dyn = (size_t)PT * NQ * 12;
  static const bool fast4_enabled = []() { const char* e = getenv("B200S_NETOH_FAST4"); return !(e && e[0] == '0'); }();
  const bool fast4 = fast4_enabled && dyn <= 96 * 1024;
  if (fast4) {
    if (!c->netoh_attr_set) {  // per device
      B200S_CUDA_TRY(c, cudaFuncSetAttribute(netoh_fast4_kernel<1, uint8_t, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      B200S_CUDA_TRY(c, cudaFuncSetAttribute(netoh_fast4_kernel<2, uint8_t, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      B200S_CUDA_TRY(c, cudaFuncSetAttribute(netoh_fast4_kernel<2, int64_t, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      c->netoh_attr_set = true;
    }
  }
  dim3 grid((Npad + 511) / 512, (P + PT - 1) / PT);
  dim3 grid4((Npad + 1023) / 1024, (P + PT - 1) / PT);
#define NETOH_FAST_ARGS(outp)                                                                                          \
  t, c->netoh_pair_id.as<int32_t>(), NQ, c->netoh_pair_cost.as<int64_t>(), c->netoh_pair_sv.as<uint32_t>(), c->node_off, \
      c->netoh_equal.as<uint8_t>(), c->netoh_dep_off.as<int32_t>(), c->netoh_deps.as<b200s_netoh_dep>(),                 \
      c->netoh_region.as<uint16_t>(), c->netoh_zone.as<uint16_t>(), c->upstream_mask(), words, N, Npad, P,               \
      c->netoh_apply_filter, o.feas.as<uint64_t>(), o.reasons.as<uint8_t>(), lo_buf, hi_buf,                             \
      c->norm_params.as<NormParam>(), outp
  {
    KernelTimer kt(c, B200S_PLUGIN_NETWORK_OVERHEAD);
    if (fast4)
      netoh_fast4_kernel<1, uint8_t, PT><<<grid4, 256, dyn, c->stream>>>(NETOH_FAST_ARGS((uint8_t*)nullptr));
    else
      netoh_fast_kernel<1, uint8_t, PT><<<grid, 256, 0, c->stream>>>(NETOH_FAST_ARGS((uint8_t*)nullptr));
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  B200S_TRY(comm_allreduce_minmax(c, lo_buf, hi_buf, P));
  netoh_params_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(lo_buf, hi_buf, P, c->norm_params.as<NormParam>());
  c->launches++;
  if (fast4 && dtype == B200S_OUT_I64)
    netoh_fast4_kernel<2, int64_t, PT><<<grid4, 256, dyn, c->stream>>>(NETOH_FAST_ARGS(o.scores.as<int64_t>()));
  else if (fast4)
    netoh_fast4_kernel<2, uint8_t, PT><<<grid4, 256, dyn, c->stream>>>(NETOH_FAST_ARGS(o.scores.as<uint8_t>()));
  else if (dtype == B200S_OUT_I64)
    netoh_fast_kernel<2, int64_t, PT><<<grid, 256, 0, c->stream>>>(NETOH_FAST_ARGS(o.scores.as<int64_t>()));
  else
    netoh_fast_kernel<2, uint8_t, PT><<<grid, 256, 0, c->stream>>>(NETOH_FAST_ARGS(o.scores.as<uint8_t>()));
#undef NETOH_FAST_ARGS
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  o.valid = true;
  c->netoh_raw_P = -1;  // finalCostMap is not materialised on this path
  return B200S_OK;
}

}  // namespace

int netoh_eval(b200s_ctx* c, int dtype) {
  if (!c->has_netoh) return c->set_err(B200S_ERR_STATE, "NetworkOverhead: snapshot has no region/zone columns");
  if (!c->has_netoh_pods) return c->set_err(B200S_ERR_STATE, "NetworkOverhead: pod batch has no dependency lists");
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_NETWORK_OVERHEAD, dtype, true, true));
  PluginOut& o = c->out[B200S_PLUGIN_NETWORK_OVERHEAD];
  if (P == 0) {
    o.valid = true;
    return B200S_OK;
  }
  // throughput path: per-pod pair table; the PreFilterState maps (raw cost / counts, opt-in via
  // b200s_config_network_overhead) keep the materialising 3-pass path below
  if (!c->netoh_want_counts && c->netoh_NQ <= 4096 && c->netoh_max_deps < 65535 &&
      (size_t)P * c->netoh_NQ <= (size_t)1 << 31)
    return netoh_eval_fast(c, dtype);
  B200S_CUDA_TRY(c, c->raw_scores.ensure((size_t)P * Npad * 8));
  if (c->netoh_want_counts) B200S_CUDA_TRY(c, c->netoh_counts.ensure((size_t)P * Npad * 4));
  B200S_CUDA_TRY(c, c->pod_lo.ensure((size_t)P * 16));  // [lo | hi] contiguous: one all-reduce when sharded
  int64_t* const lo_buf = c->pod_lo.as<int64_t>();
  int64_t* const hi_buf = lo_buf + P;
  B200S_CUDA_TRY(c, c->norm_params.ensure((size_t)P * sizeof(NormParam)));
  Topo t{c->netoh_zone_cost.as<int64_t>(), c->netoh_region_cost.as<int64_t>(), c->netoh_K};
  constexpr int PT = 32;
  dim3 grid((Npad / 64 + 7) / 8, (P + PT - 1) / PT);
  {
    KernelTimer kt(c, B200S_PLUGIN_NETWORK_OVERHEAD);
    netoh_raw_kernel<PT><<<grid, 256, 0, c->stream>>>(
        t, c->netoh_region.as<uint16_t>(), c->netoh_zone.as<uint16_t>(), c->node_off, c->netoh_equal.as<uint8_t>(),
        c->netoh_dep_off.as<int32_t>(), c->netoh_deps.as<b200s_netoh_dep>(),
        c->upstream_mask(), words, N, Npad, P, c->raw_scores.as<int64_t>(),
        o.feas.as<uint64_t>(), o.reasons.as<uint8_t>(), c->netoh_want_counts ? c->netoh_counts.as<uint32_t>() : nullptr, c->netoh_apply_filter);
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  row_minmax_kernel<<<P, 256, 0, c->stream>>>(c->raw_scores.as<int64_t>(), o.feas.as<uint64_t>(), words, Npad,
                                              lo_buf, hi_buf);
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  B200S_TRY(comm_allreduce_minmax(c, lo_buf, hi_buf, P));
  netoh_params_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(lo_buf, hi_buf, P,
                                                               c->norm_params.as<NormParam>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  if (dtype == B200S_OUT_I64)
    netoh_norm_kernel<int64_t, PT><<<grid, 256, 0, c->stream>>>(c->raw_scores.as<int64_t>(), o.feas.as<uint64_t>(),
                                                                c->norm_params.as<NormParam>(), words, Npad, P,
                                                                o.scores.as<int64_t>());
  else
    netoh_norm_kernel<uint8_t, PT><<<grid, 256, 0, c->stream>>>(c->raw_scores.as<int64_t>(), o.feas.as<uint64_t>(),
                                                                c->norm_params.as<NormParam>(), words, Npad, P,
                                                                o.scores.as<uint8_t>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  o.valid = true;
  c->netoh_raw_P = P;
  return B200S_OK;
}

}  // namespace b200s

// NodeResourceTopologyMatch: Filter (single-numa-node, pod/container scope) + Score (Least/Most/
// Balanced/LeastNUMANodes), all pods x all nodes on the dense encoding of include/b200sched.h.
//
// Reference semantics (pkg/noderesourcetopology):
//   Filter   filter.go:176-225, handlers :39-78 / :162-173, resourcesAvailableInAnyNUMANodes :90-160,
//            isResourceSetSuitable numaresources.go:137-142, subtractResourcesFromNUMANodeList :145-182
//   Score    score.go:62-102, scoreForEachNUMANode :110-124, pod/container scope :142-165,
//            least_allocated.go:25-55, most_allocated.go:25-54, balanced_allocation.go:27-54,
//            least_numa.go:35-233, subtractFromNUMAs numaresources.go:184-215
// The reference re-parses zone names and deep-copies every Quantity map per (pod,node) call; here
// a node's zones x resources block (exact milli-units) is loaded ONCE into registers and reused
// for the whole pod tile.  One thread owns one node (through the class/capacity-sorted permutation); the
// feasibility words are rebuilt from the reason codes by nrt_feas_kernel.
// Kernels are instantiated for the common shapes (<=2 zones x <=4 resources, <=4 x <=4, <=8 x <=8)
// so the zone/resource loops unroll and the working copy used by the container-scope state
// machine stays in registers.
#include "engine.h"
#include "nrt_device.cuh"

namespace b200s {

namespace {

using namespace nrtdev;

bool g_combo_ready[64] = {false};  // __constant__ memory is per device (one process may drive several GPUs)

int ensure_combos(int device) {
  if (device < 0 || device >= 64) return -1;
  if (g_combo_ready[device]) return 0;
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
  g_combo_ready[device] = true;
  return 0;
}

struct NrtNodeCols {
  const uint8_t* node_flags;
  const uint16_t* max_numa;
  const uint8_t* nz;
  const uint8_t* node_res_mask;
  const uint8_t* zone_res_mask;  // [Zs][Npad]
  const int64_t* avail;          // [Zs][Rs][Npad]
  const int32_t* cost;           // [Zs][Zs][Npad] or null
  const int32_t* perm;           // [Npad] thread slot -> node
  int Zs, Rs;
};
struct NrtPodCols {
  const uint8_t* qos;
  const uint8_t* flags;
  const uint8_t* n_init;
  const uint8_t* n_app;
  const uint8_t* kind;      // [P][8]
  const uint8_t* req_mask;  // [P][9]
  const int64_t* req;       // [P][9][Rs]
};

// tuning knobs (defaults = the measured best; overridable with -D for A/B builds)
#ifndef B200S_NRT_MIN_BLOCKS
#define B200S_NRT_MIN_BLOCKS 3
#endif
#ifndef B200S_NRT_MIN_BLOCKS_SC2  // LeastNUMANodes instantiation
#define B200S_NRT_MIN_BLOCKS_SC2 B200S_NRT_MIN_BLOCKS
#endif
#ifndef B200S_NRT_POD_UNROLL
#define B200S_NRT_POD_UNROLL 1
#endif
#ifndef B200S_NRT_PT
#define B200S_NRT_PT 32
#endif
#define B200S_PRAGMA(x) _Pragma(#x)
#define B200S_UNROLL(n) B200S_PRAGMA(unroll n)

template <int Z, int R, int SC, class OutT, int PT>
__global__ void __launch_bounds__(128, (Z <= 4 ? (SC == 2 ? B200S_NRT_MIN_BLOCKS_SC2 : B200S_NRT_MIN_BLOCKS) : 1))
nrt_kernel(NrtNodeCols nc, NrtPodCols pc, NrtCfg cfg, const uint64_t* __restrict__ upstream, int words, int N,
           int Npad, int P, OutT* __restrict__ out, uint32_t* __restrict__ feas_out32, uint8_t* __restrict__ reasons) {
  __shared__ PodS<R> sp[PT];
  // thread slot -> node through the class-sorted permutation (Npad is a multiple of 128: every slot has a node)
  const int n = nc.perm[blockIdx.x * 128 + threadIdx.x];
  const int p0 = blockIdx.y * PT, pend = min(PT, P - p0);
  // stage the pod tile
  for (int i = threadIdx.x; i < pend * (C_MAX + 1) * R; i += 128) {
    const int pp = i / ((C_MAX + 1) * R), rest = i % ((C_MAX + 1) * R), c = rest / R, r = rest % R;
    const int64_t q = r < nc.Rs ? pc.req[((size_t)(p0 + pp) * (C_MAX + 1) + c) * nc.Rs + r] : 0;
    sp[pp].req[c][r] = q;
    sp[pp].reqv[c][r] = q >= 0 ? (q + 999) / 1000 : -((-q) / 1000);
    {
      const uint32_t m = r < nc.Rs ? pc.req_mask[(size_t)(p0 + pp) * (C_MAX + 1) + c] : 0u;
      const bool needed = ((m >> r) & 1u) && q != 0;
      const bool exempt = pc.qos[p0 + pp] != B200S_QOS_GUARANTEED && (cfg.res_flags[r] & B200S_NRT_RES_AFFINE);
      sp[pp].eff[c][r] = needed ? (exempt ? INT64_MIN + 1 : q) : INT64_MIN;
      sp[pp].sub[c][r] = (needed && !exempt) ? q : 0;
    }
  }
  for (int i = threadIdx.x; i < pend; i += 128) {
    const int p = p0 + i;
    sp[i].qos = pc.qos[p];
    sp[i].flags = pc.flags[p];
    sp[i].n_init = pc.n_init[p];
    sp[i].n_app = pc.n_app[p];
    for (int c = 0; c < C_MAX; ++c) sp[i].kind[c] = pc.kind[(size_t)p * C_MAX + c];
    for (int c = 0; c <= C_MAX; ++c) {
      const uint32_t m = pc.req_mask[(size_t)p * (C_MAX + 1) + c];
      sp[i].req_mask[c] = (uint8_t)m;
      uint32_t need = 0;
      for (int r = 0; r < nc.Rs; ++r)
        if (((m >> r) & 1u) && pc.req[((size_t)p * (C_MAX + 1) + c) * nc.Rs + r] != 0) need |= 1u << r;
      sp[i].need[c] = (uint8_t)need;
    }
  }
  // this thread's node: zones x resources block into registers, once for the whole pod tile
  Zones<Z, R> zs;
  int32_t cost[Z][Z];
  uint32_t nflags = 0, node_res_mask = 0;
  int max_numa = 8;
  zs.nz = 0;
  const bool in = n < Npad;
  if (in) {
    nflags = n < N ? nc.node_flags[n] : 0;
    node_res_mask = nc.node_res_mask[n];
    max_numa = nc.max_numa[n];
    if (max_numa < 1) max_numa = 1;
    zs.nz = min((int)nc.nz[n], Z);
  }
#pragma unroll
  for (int z = 0; z < Z; ++z) {
    zs.zmask[z] = (in && z < nc.Zs) ? nc.zone_res_mask[(size_t)z * Npad + n] : 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
      zs.avail[z][r] = (in && z < nc.Zs && r < nc.Rs) ? nc.avail[((size_t)z * nc.Rs + r) * Npad + n] : 0;
    if constexpr (SC == 2) {
#pragma unroll
      for (int z2 = 0; z2 < Z; ++z2)
        cost[z][z2] = (in && nc.cost && z < nc.Zs && z2 < nc.Zs) ? nc.cost[((size_t)z * nc.Zs + z2) * Npad + n] : -1;
    }
  }
  // filter encoding of the node (see nrt_filter): zones beyond nz list nothing
#pragma unroll
  for (int z = 0; z < Z; ++z)
    if (z >= zs.nz) zs.zmask[z] = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    uint32_t any = 0;
#pragma unroll
    for (int z = 0; z < Z; ++z) any |= (zs.zmask[z] >> r) & 1u;
    const int64_t none = (!any && (cfg.res_flags[r] & B200S_NRT_RES_HOST_LEVEL)) ? INT64_MAX : INT64_MIN;
#pragma unroll
    for (int z = 0; z < Z; ++z)
      if (!((zs.zmask[z] >> r) & 1u)) zs.avail[z][r] = none;
  }
  __syncthreads();
  if (!in) return;
  const int word = n >> 6;
  // The warp walks the pods of the tile together: pod data is a shared-memory broadcast and every branch on it is
  // warp-uniform.  (Measured alternative: per-lane work lists of the surviving pods keep all lanes busy but turn
  // those broadcasts and uniform branches into divergent ones: 14.5 ms instead of 11.5 ms at c4.)
  B200S_UNROLL(B200S_NRT_POD_UNROLL)
  for (int pp = 0; pp < pend; ++pp) {
    const int p = p0 + pp;
    const PodS<R>& pod = sp[pp];
    int reason = 0;
    bool feasible = false;
    int64_t score = 0;
    if (n < N) {
      reason = nrt_filter<Z, R>(zs, nflags, node_res_mask, cfg, pod);
      const bool up = upstream ? ((upstream[(size_t)p * words + word] >> (n & 63)) & 1ull) : true;
      feasible = reason == 0 && up;
      if (reason == 0 && !up) reason = B200S_REASON_UPSTREAM;
      if (feasible) score = nrt_score<Z, R, SC>(zs, cost, nflags, max_numa, cfg, pod);
    }
    // scattered (permuted) stores; the feasibility words are rebuilt from the reason codes afterwards
    out[(size_t)p * Npad + n] = (OutT)score;
    reasons[(size_t)p * Npad + n] = (uint8_t)reason;
  }
}

// feasible <=> reason == OK (own rejects, UNSUPPORTED and upstream-infeasible all carry a non-zero code)
__global__ void nrt_feas_kernel(const uint8_t* __restrict__ reasons, int N, int Npad, size_t total,
                                uint32_t* __restrict__ feas32) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool f = false;
  if (i < total) f = reasons[i] == B200S_REASON_OK && (int)(i % Npad) < N;
  const uint32_t w = __ballot_sync(0xffffffffu, f);
  if ((threadIdx.x & 31) == 0 && i < total) feas32[i >> 5] = w;
}

template <int Z, int R, int SC>
int launch_sc(b200s_ctx* c, int dtype, const NrtNodeCols& nc, const NrtPodCols& pc, const NrtCfg& cfg) {
  constexpr int PT = (R <= 4 ? B200S_NRT_PT : 16);
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  const uint64_t* up = c->upstream_mask();
  dim3 grid((Npad + 127) / 128, (P + PT - 1) / PT);
  if (dtype == B200S_OUT_I64)
    nrt_kernel<Z, R, SC, int64_t, PT><<<grid, 128, 0, c->stream>>>(nc, pc, cfg, up, words, N, Npad, P,
                                                               o.scores.as<int64_t>(), o.feas.as<uint32_t>(),
                                                               o.reasons.as<uint8_t>());
  else
    nrt_kernel<Z, R, SC, uint8_t, PT><<<grid, 128, 0, c->stream>>>(nc, pc, cfg, up, words, N, Npad, P,
                                                               o.scores.as<uint8_t>(), o.feas.as<uint32_t>(),
                                                               o.reasons.as<uint8_t>());
  const size_t total = (size_t)P * Npad;
  nrt_feas_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(o.reasons.as<uint8_t>(), N, Npad, total,
                                                                         o.feas.as<uint32_t>());
  c->launches += 2;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

template <int Z, int R>
int launch(b200s_ctx* c, int dtype, const NrtNodeCols& nc, const NrtPodCols& pc, const NrtCfg& cfg) {
  if (cfg.strategy == B200S_NRT_LEAST_NUMA_NODES) return launch_sc<Z, R, 2>(c, dtype, nc, pc, cfg);
  if (cfg.strategy == B200S_NRT_BALANCED_ALLOCATION) return launch_sc<Z, R, 1>(c, dtype, nc, pc, cfg);
  return launch_sc<Z, R, 0>(c, dtype, nc, pc, cfg);
}

}  // namespace

int nrt_eval(b200s_ctx* c, int dtype) {
  if (!c->has_nrt) return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: snapshot has no NRT columns");
  if (!c->has_nrt_pods) return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: pod batch has no NRT columns");
  if (c->nrt_strategy == B200S_NRT_LEAST_NUMA_NODES && !c->nrt_has_cost)
    return c->set_err(B200S_ERR_STATE, "NodeResourceTopologyMatch: LeastNUMANodes needs the zone cost columns");
  if (ensure_combos(c->device) != 0) return c->set_err(B200S_ERR_CUDA, "NodeResourceTopologyMatch: combination table upload failed");
  B200S_TRY(ensure_out(c, B200S_PLUGIN_NRT, dtype, true, true));
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  if (c->P == 0) {
    o.valid = true;
    return B200S_OK;
  }
  NrtNodeCols nc{c->nrt_node_flags.as<uint8_t>(), c->nrt_max_numa.as<uint16_t>(), c->nrt_nz.as<uint8_t>(),
                 c->nrt_node_res_mask.as<uint8_t>(), c->nrt_zone_res_mask.as<uint8_t>(), c->nrt_avail.as<int64_t>(),
                 c->nrt_has_cost ? c->nrt_cost.as<int32_t>() : nullptr, c->nrt_perm.as<int32_t>(), c->nrt_Z, c->nrt_R};
  NrtPodCols pc{c->nrt_pod_qos.as<uint8_t>(), c->nrt_pod_flags.as<uint8_t>(), c->nrt_pod_ninit.as<uint8_t>(),
                c->nrt_pod_napp.as<uint8_t>(), c->nrt_pod_kind.as<uint8_t>(), c->nrt_pod_req_mask.as<uint8_t>(),
                c->nrt_pod_req.as<int64_t>()};
  NrtCfg cfg;
  cfg.strategy = c->nrt_strategy;
  for (int r = 0; r < B200S_NRT_MAX_RES; ++r) {
    cfg.w[r] = c->nrt_w[r];
    cfg.res_flags[r] = c->nrt_res_flags[r];
  }
  // A batch goes through the batched path of nrt2.cu when it applies (score tables per distinct request vector,
  // 32-bit scaled arithmetic, coalesced natural-order rows); single cycles, LeastNUMANodes and shapes / ranges
  // outside it keep the direct per-(pod, node) kernel below.
  const int batched = nrt2_prepare(c);
  if (batched < 0) return batched;
  KernelTimer kt(c, B200S_PLUGIN_NRT);
  if (batched == 1) {
    B200S_TRY(nrt2_eval(c, dtype));
    o.valid = true;
    return B200S_OK;
  }
  nrt2_note_direct(c);
  int rc;
  if (c->nrt_Z <= 2 && c->nrt_R <= 4)
    rc = launch<2, 4>(c, dtype, nc, pc, cfg);
  else if (c->nrt_Z <= 4 && c->nrt_R <= 4)
    rc = launch<4, 4>(c, dtype, nc, pc, cfg);
  else
    rc = launch<8, 8>(c, dtype, nc, pc, cfg);
  if (rc != B200S_OK) return rc;
  o.valid = true;
  return B200S_OK;
}

}  // namespace b200s

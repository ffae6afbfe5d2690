// NodeResourcesAllocatable: Score + NormalizeScore for all pods x all nodes.
//
// Reference semantics (pkg/noderesources):
//   raw(n)   = ( sum_r sign * alloc_r(n) * w_r ) / sum_r w_r        allocatable.go:117-140
//              int64 wrapping, Go truncating division; sign = -1 (Least) / +1 (Most), 0 otherwise.
//              The scorer never reads `requested` (allocatable.go:122) => raw is POD-INDEPENDENT.
//   norm(p,n)= (raw(n) - lo_p) * 100 / (hi_p - lo_p)                allocatable.go:143-168
//              lo_p/hi_p = min/max of raw over the pod's FEASIBLE nodes; range 0 => 0.
//
// B200 design (not a translation):
//   * raw[] is computed once per (snapshot, args) and radix-sorted (CUB, snapshot time, off the
//     hot path).  Per pod, lo/hi are then the first/last FEASIBLE entries of the sorted order:
//     one warp ballots 32 sorted positions at a time from each end — O(1/density) probes per pod
//     instead of an O(N) reduction per pod.
//   * The P x N pass is a pure streaming-store kernel: each thread keeps its nodes' raw values
//     in registers, walks the pod tile, and divides with a per-pod 32-bit magic reciprocal
//     (range*100 < 2^32; anything else takes the exact generic int64 path).  The only HBM
//     traffic is the score matrix itself (8 B/eval for int64, 1 B/eval for u8), written with
//     128-bit evict-first stores.  Inputs (raw 4-8 B/node, 32 B/pod, 1 bit/eval mask) stay in L2.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "engine.h"

namespace b200s {

namespace {

struct AllocArgs {
  int64_t w[16];
  int R;
  int mode;
};

__global__ void alloc_raw_kernel(const int64_t* __restrict__ cols, int N, int Npad, AllocArgs a,
                                 int64_t* __restrict__ raw, int32_t* __restrict__ iota) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Npad) return;
  int64_t node_score = 0, weight_sum = 0;
  for (int r = 0; r < a.R; ++r) {
    int64_t cap = cols[(size_t)r * Npad + n];
    int64_t rs = a.mode == B200S_ALLOC_LEAST ? wrap_mul(-1, cap) : (a.mode == B200S_ALLOC_MOST ? cap : 0);
    node_score = wrap_add(node_score, wrap_mul(rs, a.w[r]));
    weight_sum = wrap_add(weight_sum, a.w[r]);
  }
  raw[n] = n < N ? go_div(node_score, weight_sum) : 0;
  if (n < N) iota[n] = n;
}

__device__ __forceinline__ NormParam make_norm_param(int64_t lo, int64_t hi);

// One warp per pod.  lo = raw of the first feasible node in ascending sorted order, hi = last.
// Single-GPU: also emits the pod's NormParam (no all-reduce in between, one launch less per step).
// The scan from each end is bounded: behind a chain of filters a pod's feasible set can be sparse or EMPTY (a
// Guaranteed pod that fits no NUMA zone anywhere), and walking the whole sorted order with dependent gathers costs
// milliseconds per pod.  After kScanRounds x 32 entries the warp switches to one coalesced pass over the pod's
// feasibility words and takes the exact min / max of the raw scores of the set bits -- the same two values.
constexpr int kScanRounds = 16;
__global__ void alloc_minmax_kernel(const int64_t* __restrict__ sorted_raw, const int32_t* __restrict__ order,
                                    const int64_t* __restrict__ raw, const uint64_t* __restrict__ feasible, int words, int N,
                                    int P, int64_t* __restrict__ lo, int64_t* __restrict__ hi,
                                    NormParam* __restrict__ params_out) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= P) return;
  const uint64_t* row = feasible ? feasible + (size_t)warp * words : nullptr;
  int64_t vlo = INT64_MAX, vhi = -INT64_MAX;  // allocatable.go:145-146
  if (N > 0) {
    if (!row) {
      vlo = sorted_raw[0];
      vhi = sorted_raw[N - 1];
    } else {
      bool found_lo = false, found_hi = false;
      const int span = min(N, 32 * kScanRounds);
      for (int base = 0; base < span; base += 32) {
        int i = base + lane;
        bool f = false;
        if (i < N) {
          int node = order[i];
          f = (row[node >> 6] >> (node & 63)) & 1ull;
        }
        unsigned b = __ballot_sync(0xffffffffu, f);
        if (b) {
          vlo = sorted_raw[base + __ffs(b) - 1];
          found_lo = true;
          break;
        }
      }
      if (found_lo) {  // an empty or sparse row goes straight to the full pass
        for (int top = N - 1; top > N - 1 - span; top -= 32) {
          int i = top - lane;
          bool f = false;
          if (i >= 0) {
            int node = order[i];
            f = (row[node >> 6] >> (node & 63)) & 1ull;
          }
          unsigned b = __ballot_sync(0xffffffffu, f);
          if (b) {
            vhi = sorted_raw[top - (__ffs(b) - 1)];
            found_hi = true;
            break;
          }
        }
      }
      if (!found_lo || !found_hi) {
        int64_t mn = INT64_MAX, mx = -INT64_MAX;
        for (int w = lane; w < words; w += 32) {
          uint64_t bits = row[w];
          while (bits) {
            const int n = w * 64 + __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (n < N) {
              const int64_t r = raw[n];
              mn = r < mn ? r : mn;
              mx = r > mx ? r : mx;
            }
          }
        }
        for (int o = 16; o; o >>= 1) {
          const int64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
          mn = a < mn ? a : mn;
          mx = b > mx ? b : mx;
        }
        vlo = mn;
        vhi = mx;
      }
    }
  }
  if (lane == 0) {
    lo[warp] = vlo;
    hi[warp] = vhi;
    if (params_out) params_out[warp] = make_norm_param(vlo, vhi);
  }
}

__device__ __forceinline__ NormParam make_norm_param(int64_t lo, int64_t hi) {
  NormParam q;
  q.lo = lo;
  q.range = wrap_sub(hi, lo);
  q.magic = 0;
  q.shift = 0;
  q.pad = 0;
  if (lo > hi || q.range == 0) {
    q.mode = 0;  // empty feasible set, or oldRange == 0 => MinNodeScore (allocatable.go:158-160)
  } else if (q.range > 0 && q.range <= (int64_t)(0xffffffffu / 100u)) {
    // (raw-lo)*100 < 2^32 for every feasible node: exact 32-bit reciprocal division with one fix-up.
    uint32_t r = (uint32_t)q.range;
    uint32_t s = 31 - __clz(r);
    uint64_t m = ((1ull << 32) << s) / r;  // floor(2^(32+s)/r) in (2^31, 2^32]
    q.magic = m > 0xffffffffull ? 0xffffffffu : (uint32_t)m;
    q.shift = s;
    q.mode = 1;
  } else {
    q.mode = 2;  // generic wrapping int64 path (Go semantics incl. overflow)
  }
  return q;
}

__global__ void norm_params_kernel(const int64_t* __restrict__ lo, const int64_t* __restrict__ hi, int P,
                                   NormParam* __restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) out[p] = make_norm_param(lo[p], hi[p]);
}

// Tile: (256 threads x NPT nodes) x PT pods.  Thread t owns nodes n0 + t*NPT .. +NPT-1.
template <class OutT, int NPT, int PT>
__global__ void __launch_bounds__(256)
alloc_norm_kernel(const int64_t* __restrict__ raw, const NormParam* __restrict__ params,
                  const uint64_t* __restrict__ feasible, int words, int N, int Npad, int P,
                  OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  __shared__ NormParam sp[PT];
  __shared__ uint64_t sm[PT][CHUNK / 64];
  const int n0 = blockIdx.x * CHUNK;
  const int p0 = blockIdx.y * PT;
  const int t = threadIdx.x;
  const int nb = n0 + t * NPT;

  for (int i = t; i < PT; i += 256)
    if (p0 + i < P) sp[i] = params[p0 + i];
  if (feasible) {
    for (int i = t; i < PT * (CHUNK / 64); i += 256) {
      int pp = i / (CHUNK / 64), w = i % (CHUNK / 64);
      int gw = n0 / 64 + w;
      sm[pp][w] = (p0 + pp < P && gw < words) ? feasible[(size_t)(p0 + pp) * words + gw] : 0ull;
    }
  }
  uint32_t r32[NPT];
  uint32_t valid = 0;
  if (nb < Npad) {
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      r32[j] = (uint32_t)(uint64_t)raw[nb + j];
      if (nb + j < N) valid |= 1u << j;
    }
  }
  __syncthreads();
  if (nb >= Npad) return;
  const int wi = (t * NPT) / 64, sh = (t * NPT) % 64;
  const int pend = min(PT, P - p0);
  OutT* orow = out + (size_t)p0 * Npad + nb;
  for (int pp = 0; pp < pend; ++pp, orow += Npad) {
    const NormParam np = sp[pp];
    uint32_t bits = valid;
    if (feasible) bits &= (uint32_t)(sm[pp][wi] >> sh);
    if (np.mode == 1) {
      const uint32_t lo32 = (uint32_t)(uint64_t)np.lo, rg = (uint32_t)np.range;
      uint32_t q[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        uint32_t n100 = (r32[j] - lo32) * 100u;
        uint32_t q0 = __umulhi(n100, np.magic) >> np.shift;
        uint32_t rem = n100 - q0 * rg;
        q0 += rem >= rg ? 1u : 0u;
        q[j] = ((bits >> j) & 1u) ? q0 : 0u;
      }
      Store<OutT, NPT>::put32(orow, q);
    } else if (np.mode == 0) {
      uint32_t q[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) q[j] = 0;
      Store<OutT, NPT>::put32(orow, q);
    } else {
      int64_t q[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        int64_t v = go_div(wrap_mul(wrap_sub(raw[nb + j], np.lo), 100), np.range);
        q[j] = ((bits >> j) & 1u) ? v : 0;
      }
      Store<OutT, NPT>::put64(orow, q);
    }
  }
}

// ---- TMA-store variant of the int64 pass (experiment, selected with B200S_ALLOC_TMA=1) -------------------
// Same arithmetic; the scores of PB pods x 512 nodes are staged in shared memory (registers -> STS.128) and
// one elected thread hands each 4 KB row segment to the TMA engine (cp.async.bulk.global.shared::cta), NSTAGE
// staging buffers deep, so the LSU issues shared-memory stores only and global stores are bulk, asynchronous.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int PT, int PB, int NSTAGE>
__global__ void __launch_bounds__(256)
alloc_norm_tma_kernel(const int64_t* __restrict__ raw, const NormParam* __restrict__ params,
                      const uint64_t* __restrict__ feasible, int words, int N, int Npad, int P,
                      int64_t* __restrict__ out) {
  constexpr int NPT = 2, CHUNK = 256 * NPT;
  extern __shared__ __align__(128) unsigned char dyn[];
  int64_t* stage = reinterpret_cast<int64_t*>(dyn);  // [NSTAGE][PB][CHUNK]
  __shared__ NormParam sp[PT];
  __shared__ uint64_t sm[PT][CHUNK / 64];
  const int n0 = blockIdx.x * CHUNK, p0 = blockIdx.y * PT, t = threadIdx.x, nb = n0 + t * NPT;
  for (int i = t; i < PT; i += 256)
    if (p0 + i < P) sp[i] = params[p0 + i];
  if (feasible)
    for (int i = t; i < PT * (CHUNK / 64); i += 256) {
      int pp = i / (CHUNK / 64), w = i % (CHUNK / 64), gw = n0 / 64 + w;
      sm[pp][w] = (p0 + pp < P && gw < words) ? feasible[(size_t)(p0 + pp) * words + gw] : 0ull;
    }
  uint32_t r32[NPT];
  uint32_t valid = 0;
  const bool in = nb < Npad;
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    r32[j] = in ? (uint32_t)(uint64_t)raw[nb + j] : 0u;
    if (in && nb + j < N) valid |= 1u << j;
  }
  __syncthreads();
  const int wi = (t * NPT) / 64, sh = (t * NPT) % 64;
  const int pend = min(PT, P - p0);
  const int cols = min(CHUNK, Npad - n0);  // columns of this CTA that exist (multiple of 128)
  int g = 0;
  for (int pg = 0; pg < pend; pg += PB, ++g) {
    const int s = g % NSTAGE;
    if (g >= NSTAGE) {  // the buffer's previous bulk stores must have finished READING shared memory
      if (t == 0) bulk_wait_read<NSTAGE - 1>();
      __syncthreads();
    }
    int64_t* buf = stage + (size_t)s * PB * CHUNK;
#pragma unroll
    for (int b = 0; b < PB; ++b) {
      const int pp = pg + b;
      if (pp >= pend) break;
      const NormParam np = sp[pp];
      uint32_t bits = valid;
      if (feasible) bits &= (uint32_t)(sm[pp][wi] >> sh);
      int64_t q[NPT];
      if (np.mode == 1) {
        const uint32_t lo32 = (uint32_t)(uint64_t)np.lo, rg = (uint32_t)np.range;
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          uint32_t n100 = (r32[j] - lo32) * 100u;
          uint32_t q0 = __umulhi(n100, np.magic) >> np.shift;
          uint32_t rem = n100 - q0 * rg;
          q0 += rem >= rg ? 1u : 0u;
          q[j] = ((bits >> j) & 1u) ? (int64_t)q0 : 0;
        }
      } else if (np.mode == 0) {
#pragma unroll
        for (int j = 0; j < NPT; ++j) q[j] = 0;
      } else {
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
          int64_t v = in ? go_div(wrap_mul(wrap_sub(raw[nb + j], np.lo), 100), np.range) : 0;
          q[j] = ((bits >> j) & 1u) ? v : 0;
        }
      }
      *reinterpret_cast<longlong2*>(buf + (size_t)b * CHUNK + t * NPT) = make_longlong2(q[0], q[1]);
    }
    fence_async_smem();  // generic-proxy writes -> visible to the async (TMA) proxy
    __syncthreads();
    if (t == 0) {
      for (int b = 0; b < PB && pg + b < pend; ++b)
        bulk_store_s2g(out + (size_t)(p0 + pg + b) * Npad + n0, buf + (size_t)b * CHUNK, (uint32_t)cols * 8u);
      bulk_commit();
    }
  }
  if (t == 0) bulk_wait_read<0>();  // shared memory must stay alive until the engine has read it
  __syncthreads();
}

}  // namespace

int build_norm_params(b200s_ctx* c, int P) {
  B200S_CUDA_TRY(c, c->norm_params.ensure((size_t)P * sizeof(NormParam)));
  if (P == 0) return B200S_OK;
  norm_params_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(c->pod_lo.as<int64_t>(), c->pod_lo.as<int64_t>() + P, P,
                                                              c->norm_params.as<NormParam>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

int alloc_prepare(b200s_ctx* c) {
  uint64_t key = c->snap_serial * 1000003ull + c->alloc_cfg_gen;
  if (c->alloc_prepared_key == key) return B200S_OK;
  const int N = c->N, Npad = c->Npad;
  B200S_CUDA_TRY(c, c->alloc_raw.ensure((size_t)Npad * 8));
  B200S_CUDA_TRY(c, c->alloc_sorted_raw.ensure((size_t)Npad * 8));
  B200S_CUDA_TRY(c, c->alloc_order.ensure((size_t)Npad * 4));
  B200S_CUDA_TRY(c, c->alloc_iota.ensure((size_t)Npad * 4));
  AllocArgs a;
  a.R = c->alloc_R;
  a.mode = c->alloc_mode;
  for (int r = 0; r < 16; ++r) a.w[r] = r < a.R ? c->alloc_w[r] : 0;
  alloc_raw_kernel<<<(Npad + 255) / 256, 256, 0, c->stream>>>(c->alloc_cols.as<int64_t>(), N, Npad, a,
                                                              c->alloc_raw.as<int64_t>(), c->alloc_iota.as<int32_t>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  if (N > 0) {
    size_t tmp = 0;
    B200S_CUDA_TRY(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp, c->alloc_raw.as<int64_t>(),
                                                      c->alloc_sorted_raw.as<int64_t>(), c->alloc_iota.as<int32_t>(),
                                                      c->alloc_order.as<int32_t>(), N, 0, 64, c->stream));
    B200S_CUDA_TRY(c, c->sort_tmp.ensure(tmp));
    B200S_CUDA_TRY(c, cub::DeviceRadixSort::SortPairs(c->sort_tmp.p, tmp, c->alloc_raw.as<int64_t>(),
                                                      c->alloc_sorted_raw.as<int64_t>(), c->alloc_iota.as<int32_t>(),
                                                      c->alloc_order.as<int32_t>(), N, 0, 64, c->stream));
  }
  c->alloc_prepared_key = key;
  B200S_CUDA_TRY(c, cudaEventRecord(c->ev_inputs, c->stream));  // raw scores + sorted order queued up to here
  return B200S_OK;
}

int alloc_eval(b200s_ctx* c, int dtype) {
  if (!c->has_alloc) return c->set_err(B200S_ERR_STATE, "NodeResourcesAllocatable: snapshot has no allocatable columns");
  if (!c->alloc_cfg) return c->set_err(B200S_ERR_STATE, "NodeResourcesAllocatable: args not configured");
  if (c->alloc_cfg_R != c->alloc_R)
    return c->set_err(B200S_ERR_INVALID, "NodeResourcesAllocatable: weights do not match the snapshot's resource columns");
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_ALLOCATABLE, dtype, false, false));
  if (P == 0) {
    c->out[B200S_PLUGIN_ALLOCATABLE].valid = true;
    return B200S_OK;
  }
  B200S_TRY(alloc_prepare(c));
  B200S_CUDA_TRY(c, c->pod_lo.ensure((size_t)P * 16));  // per chunk [lo | hi] contiguous: one all-reduce when sharded
  const uint64_t* feas = c->upstream_mask();
  const bool sharded = comm_world(c) > 1;
  B200S_CUDA_TRY(c, c->norm_params.ensure((size_t)P * sizeof(NormParam)));
  PluginOut& o = c->out[B200S_PLUGIN_ALLOCATABLE];
  static const bool use_tma = getenv("B200S_ALLOC_TMA") && atoi(getenv("B200S_ALLOC_TMA")) != 0;
  // pods [a, a + n): lo / hi of the feasible set (first / last feasible entry of the sorted order)
  auto launch_minmax = [&](int a, int n) {
    const int threads = 128, warps_per_block = threads / 32;
    int64_t* lo = c->pod_lo.as<int64_t>() + 2 * (size_t)a;
    alloc_minmax_kernel<<<(n + warps_per_block - 1) / warps_per_block, threads, 0, c->stream>>>(
        c->alloc_sorted_raw.as<int64_t>(), c->alloc_order.as<int32_t>(), c->alloc_raw.as<int64_t>(),
        feas ? feas + (size_t)a * words : nullptr, words, N, n, lo, lo + n, sharded ? nullptr : c->norm_params.as<NormParam>() + a);
    c->launches++;
  };
  auto launch_norm = [&](int a, int n) {
    KernelTimer kt(c, B200S_PLUGIN_ALLOCATABLE);
    const NormParam* params = c->norm_params.as<NormParam>() + a;
    const uint64_t* fm = feas ? feas + (size_t)a * words : nullptr;
    if (dtype == B200S_OUT_I64 && use_tma) {
      constexpr int PT = 64, PB = 4, NSTAGE = 4;
      constexpr size_t smem = (size_t)NSTAGE * PB * 512 * 8;
      auto kern = alloc_norm_tma_kernel<PT, PB, NSTAGE>;
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      dim3 grid((Npad + 511) / 512, (n + PT - 1) / PT);
      kern<<<grid, 256, smem, c->stream>>>(c->alloc_raw.as<int64_t>(), params, fm, words, N, Npad, n,
                                           o.scores.as<int64_t>() + (size_t)a * Npad);
    } else if (dtype == B200S_OUT_I64) {
      constexpr int NPT = 2, PT = 64;
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (n + PT - 1) / PT);
      alloc_norm_kernel<int64_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->alloc_raw.as<int64_t>(), params, fm, words, N, Npad,
                                                                      n, o.scores.as<int64_t>() + (size_t)a * Npad);
    } else {
      constexpr int NPT = 8, PT = 64;
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (n + PT - 1) / PT);
      alloc_norm_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->alloc_raw.as<int64_t>(), params, fm, words, N, Npad,
                                                                      n, o.scores.as<uint8_t>() + (size_t)a * Npad);
    }
    c->launches++;
  };
  if (!sharded) {
    launch_minmax(0, P);
    launch_norm(0, P);
  } else if (comm_has_peers(c) && c->mask_override == nullptr) {
    // Sharded NormalizeScore needs the per-pod min/max over ALL shards before any score can be written.  With the
    // peer-memory exchange the whole pre-pass (local min/max, exchange, fold, parameters) is a handful of small
    // kernels that depend on the batch's INPUTS only -- not on the previous eval.  It runs on the high-priority
    // communication stream, so back-to-back evals overlap it with the previous eval's P x N kernel; the parameter
    // buffers alternate between two copies, and a copy is reused only after the P x N kernel that read it is done.
    B200S_TRY(comm_ensure_streams(c));
    const int par = (int)(c->alloc_eval_seq++ & 1);
    DevBuf& lo_b = par ? c->pod_lo_alt : c->pod_lo;
    DevBuf& np_b = par ? c->norm_params_alt : c->norm_params;
    B200S_CUDA_TRY(c, lo_b.ensure((size_t)P * 16));
    B200S_CUDA_TRY(c, np_b.ensure((size_t)P * sizeof(NormParam)));
    cudaStream_t cs = c->comm_stream;
    B200S_CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_inputs, 0));
    if (c->ev_norm_valid[par]) B200S_CUDA_TRY(c, cudaStreamWaitEvent(cs, c->ev_norm_done[par], 0));
    int64_t* lo = lo_b.as<int64_t>();
    {
      const int threads = 128, warps_per_block = threads / 32;
      alloc_minmax_kernel<<<(P + warps_per_block - 1) / warps_per_block, threads, 0, cs>>>(
          c->alloc_sorted_raw.as<int64_t>(), c->alloc_order.as<int32_t>(), c->alloc_raw.as<int64_t>(), feas, words, N, P, lo,
          lo + P, nullptr);
      c->launches++;
    }
    B200S_TRY(comm_allreduce_minmax_on(c, cs, lo, lo + P, P));
    norm_params_kernel<<<(P + 255) / 256, 256, 0, cs>>>(lo, lo + P, P, np_b.as<NormParam>());
    c->launches++;
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_params[par], cs));
    B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_params[par], 0));
    {
      KernelTimer kt(c, B200S_PLUGIN_ALLOCATABLE);
      const NormParam* params = np_b.as<NormParam>();
      if (dtype == B200S_OUT_I64) {
        constexpr int NPT = 2, PT = 64;
        dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
        alloc_norm_kernel<int64_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->alloc_raw.as<int64_t>(), params, feas, words, N,
                                                                        Npad, P, o.scores.as<int64_t>());
      } else {
        constexpr int NPT = 8, PT = 64;
        dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
        alloc_norm_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->alloc_raw.as<int64_t>(), params, feas, words, N,
                                                                        Npad, P, o.scores.as<uint8_t>());
      }
      c->launches++;
    }
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_norm_done[par], c->stream));
    c->ev_norm_valid[par] = true;
  } else {
    // NCCL exchange (no peer mappings, or a filter chain feeding the mask on the main stream): the batch goes in up
    // to 4 pod chunks; every chunk's local min/max is launched first, each chunk's all-reduce runs on the
    // communication stream as soon as its min/max is there, and the main stream starts the P x N pass of chunk i as
    // soon as all-reduce i has landed.
    B200S_TRY(comm_ensure_streams(c));
    const int nchunk = (P >= 2048 && !comm_has_peers(c)) ? 4 : 1;
    const int step = ((P + nchunk - 1) / nchunk + 63) / 64 * 64;
    int starts[5], k = 0;
    for (int a = 0; a < P; a += step) starts[k++] = a;
    starts[k] = P;
    for (int i = 0; i < k; ++i) {
      const int a = starts[i], n = starts[i + 1] - a;
      launch_minmax(a, n);
      int64_t* lo = c->pod_lo.as<int64_t>() + 2 * (size_t)a;
      if (k == 1) {
        B200S_TRY(comm_allreduce_minmax_on(c, c->stream, lo, lo + n, n));
        continue;
      }
      B200S_CUDA_TRY(c, cudaEventRecord(c->ev_chunk[i], c->stream));
      B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->comm_stream, c->ev_chunk[i], 0));
      B200S_TRY(comm_allreduce_minmax_on(c, c->comm_stream, lo, lo + n, n));
      B200S_CUDA_TRY(c, cudaEventRecord(c->ev_reduced[i], c->comm_stream));
    }
    for (int i = 0; i < k; ++i) {
      const int a = starts[i], n = starts[i + 1] - a;
      if (k > 1) B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_reduced[i], 0));
      const int64_t* lo = c->pod_lo.as<int64_t>() + 2 * (size_t)a;
      norm_params_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(lo, lo + n, n, c->norm_params.as<NormParam>() + a);
      c->launches++;
      launch_norm(a, n);
    }
  }
  B200S_CUDA_TRY(c, cudaGetLastError());
  o.valid = true;
  return B200S_OK;
}

}  // namespace b200s

// Combined profile: the upstream scheduling cycle over the engine's plugins, all pods at once.
//
// Restated from the upstream framework runtime (k8s.io/kubernetes pkg/scheduler, NOT in the
// reference tree; SURVEY App. B "Upstream combination"):
//   feasible(n) = AND over the enabled Filter plugins (NodeResourceTopologyMatch, NetworkOverhead)
//                 and the caller's upstream mask;
//   Score plugins run on the feasible nodes only, NormalizeScore over exactly that set;
//   total(n)    = sum_p weight_p * score_p(n);
//   selectHost  = arg-max; ties are random upstream — here deterministic: (total desc, node asc).
// The filters are chained (each plugin's "upstream" set is what the previous filters left) so
// NodeResourcesAllocatable / NetworkOverhead normalise over the final feasible set, as upstream.
// Per-plugin scores travel as u8 (0..100); the weighted sum, the optional int64 total matrix and
// the per-pod top-k are produced by one kernel.  Sharded: ONE ncclAllGather of the [P][k]
// winners, then a fold — every rank ends up with the global top-k.
#include <algorithm>

#include "engine.h"

namespace b200s {

namespace {

constexpr int K_MAX = 16;

struct Cand {
  int64_t score;
  int32_t node;
};
__device__ __forceinline__ bool better(int64_t s1, int32_t n1, int64_t s2, int32_t n2) {
  return s1 > s2 || (s1 == s2 && n1 < n2);
}

struct PluginPtrs {
  const uint8_t* s[B200S_PLUGIN_COUNT];
  int64_t w[B200S_PLUGIN_COUNT];
  int n;
};

// One CTA per (pod, node slice): blockIdx.y selects a `chunk`-node slice so that a small batch (P = 1 in the
// real scheduling cycle) still fills the 148 SMs; the per-slice winners [slice][P][k] are folded afterwards.
template <int K>
__global__ void __launch_bounds__(256)
combine_topk_kernel(PluginPtrs pl, const uint64_t* __restrict__ feas, int words, int N, int Npad, int node_off, int k,
                    int chunk, int P, int64_t* __restrict__ total, b200s_topk_entry* __restrict__ out) {
  const int p = blockIdx.x, t = threadIdx.x;
  const int n_begin = blockIdx.y * chunk, n_end = min(Npad, n_begin + chunk);
  int64_t bs[K];
  int32_t bn[K];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    bs[i] = INT64_MIN;
    bn[i] = INT32_MAX;
  }
  const size_t rowoff = (size_t)p * Npad;
  for (int n4 = n_begin + t * 4; n4 < n_end; n4 += 256 * 4) {
    uint32_t packed[B200S_PLUGIN_COUNT];
    for (int j = 0; j < pl.n; ++j) packed[j] = *reinterpret_cast<const uint32_t*>(pl.s[j] + rowoff + n4);
    const uint64_t fw = feas ? feas[(size_t)p * words + (n4 >> 6)] : ~0ull;
    int64_t tot[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = n4 + e;
      const bool f = n < N && ((fw >> (n & 63)) & 1ull);
      int64_t s = 0;
      for (int j = 0; j < pl.n; ++j) s = wrap_add(s, wrap_mul(pl.w[j], (int64_t)((packed[j] >> (8 * e)) & 255u)));
      tot[e] = f ? s : 0;
      if (f) {
        const int32_t g = node_off + n;
        if (better(s, g, bs[K - 1], bn[K - 1])) {  // insertion into the sorted per-thread list
          bs[K - 1] = s;
          bn[K - 1] = g;
#pragma unroll
          for (int i = K - 1; i > 0; --i) {
            if (better(bs[i], bn[i], bs[i - 1], bn[i - 1])) {
              int64_t ts = bs[i]; bs[i] = bs[i - 1]; bs[i - 1] = ts;
              int32_t tn = bn[i]; bn[i] = bn[i - 1]; bn[i - 1] = tn;
            }
          }
        }
      }
    }
    if (total) {
      st_stream_v2(total + rowoff + n4, tot[0], tot[1]);
      st_stream_v2(total + rowoff + n4 + 2, tot[2], tot[3]);
    }
  }
  // k rounds of block arg-max over the heads of the per-thread lists
  __shared__ int64_t ss[8];
  __shared__ int32_t sn[8];
  __shared__ int64_t win_s;
  __shared__ int32_t win_n;
  for (int round = 0; round < k; ++round) {
    int64_t s = bs[0];
    int32_t n = bn[0];
    for (int o = 16; o; o >>= 1) {
      const int64_t os = __shfl_xor_sync(0xffffffffu, s, o);
      const int32_t on = __shfl_xor_sync(0xffffffffu, n, o);
      if (better(os, on, s, n)) {
        s = os;
        n = on;
      }
    }
    if ((t & 31) == 0) {
      ss[t >> 5] = s;
      sn[t >> 5] = n;
    }
    __syncthreads();
    if (t == 0) {
      for (int i = 1; i < 8; ++i)
        if (better(ss[i], sn[i], s, n)) {
          s = ss[i];
          n = sn[i];
        }
      win_s = s;
      win_n = n;
      b200s_topk_entry e;
      e.score = n == INT32_MAX ? 0 : s;
      e.node = n == INT32_MAX ? -1 : n;
      e.pad = 0;
      out[((size_t)blockIdx.y * P + p) * k + round] = e;
    }
    __syncthreads();
    if (bn[0] == win_n && win_n != INT32_MAX) {  // the owner pops its head
#pragma unroll
      for (int i = 0; i < K - 1; ++i) {
        bs[i] = bs[i + 1];
        bn[i] = bn[i + 1];
      }
      bs[K - 1] = INT64_MIN;
      bn[K - 1] = INT32_MAX;
    }
    __syncthreads();
  }
}

// k = 1 without the total matrix and with weights that keep the sum in 32 bits (the usual profile): a pure read
// stream.  16 nodes per thread and step -- one 16-byte load per plugin matrix, all in flight together -- and a running
// (score, node) maximum instead of the sorted list; a thread visits its nodes in ascending order, so `>` keeps the
// lowest node among equals, and the block reduction uses the same (score desc, node asc) order as the general kernel.
__global__ void __launch_bounds__(256)
combine_top1_kernel(PluginPtrs pl, const uint64_t* __restrict__ feas, int words, int N, int Npad, int node_off, int chunk,
                    int P, b200s_topk_entry* __restrict__ out) {
  const int p = blockIdx.x, t = threadIdx.x;
  const int n_begin = blockIdx.y * chunk, n_end = min(Npad, n_begin + chunk);
  int32_t w[B200S_PLUGIN_COUNT];
#pragma unroll
  for (int j = 0; j < B200S_PLUGIN_COUNT; ++j) w[j] = j < pl.n ? (int32_t)pl.w[j] : 0;
  int32_t best = -1, bestn = INT32_MAX;  // sums are >= 0
  const size_t rowoff = (size_t)p * Npad;
  const uint16_t* f16 = reinterpret_cast<const uint16_t*>(feas + (size_t)p * words);
#pragma unroll 2
  for (int n16 = n_begin + t * 16; n16 < n_end; n16 += 256 * 16) {
    uint4 v[B200S_PLUGIN_COUNT];
#pragma unroll
    for (int j = 0; j < B200S_PLUGIN_COUNT; ++j)
      v[j] = j < pl.n ? __ldcs(reinterpret_cast<const uint4*>(pl.s[j] + rowoff + n16)) : make_uint4(0, 0, 0, 0);
    uint32_t fb = f16[n16 >> 4];
    if (n16 + 16 > N) fb &= N > n16 ? (1u << (N - n16)) - 1u : 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int e = q * 4 + b;
        int32_t sum = 0;
#pragma unroll
        for (int j = 0; j < B200S_PLUGIN_COUNT; ++j) {
          const uint32_t word = q == 0 ? v[j].x : (q == 1 ? v[j].y : (q == 2 ? v[j].z : v[j].w));
          sum += w[j] * (int32_t)((word >> (8 * b)) & 255u);
        }
        const bool take = ((fb >> e) & 1u) && sum > best;
        best = take ? sum : best;
        bestn = take ? n16 + e : bestn;
      }
    }
  }
  int64_t s = best < 0 ? INT64_MIN : (int64_t)best;
  int32_t n = best < 0 ? INT32_MAX : node_off + bestn;
  for (int o = 16; o; o >>= 1) {
    const int64_t os = __shfl_xor_sync(0xffffffffu, s, o);
    const int32_t on = __shfl_xor_sync(0xffffffffu, n, o);
    if (better(os, on, s, n)) {
      s = os;
      n = on;
    }
  }
  __shared__ int64_t ss[8];
  __shared__ int32_t sn[8];
  if ((t & 31) == 0) {
    ss[t >> 5] = s;
    sn[t >> 5] = n;
  }
  __syncthreads();
  if (t == 0) {
    for (int i = 1; i < 8; ++i)
      if (better(ss[i], sn[i], s, n)) {
        s = ss[i];
        n = sn[i];
      }
    b200s_topk_entry e;
    e.score = n == INT32_MAX ? 0 : s;
    e.node = n == INT32_MAX ? -1 : n;
    e.pad = 0;
    out[(size_t)blockIdx.y * P + p] = e;
  }
}

// Fold the gathered [world][P][k] winners: one thread per pod, k rounds of arg-max.
__global__ void fold_topk_kernel(const b200s_topk_entry* __restrict__ all, int world, int P, int k,
                                 b200s_topk_entry* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int64_t last_s = 0;
  int32_t last_n = -1;
  bool have_last = false;
  for (int round = 0; round < k; ++round) {
    int64_t bsx = INT64_MIN;
    int32_t bnx = INT32_MAX;
    for (int r = 0; r < world; ++r)
      for (int i = 0; i < k; ++i) {
        const b200s_topk_entry e = all[((size_t)r * P + p) * k + i];
        if (e.node < 0) continue;
        // strictly after the previous winner in (score desc, node asc) order
        if (have_last && !better(last_s, last_n, e.score, e.node)) continue;
        if (better(e.score, e.node, bsx, bnx)) {
          bsx = e.score;
          bnx = e.node;
        }
      }
    b200s_topk_entry o;
    o.score = bnx == INT32_MAX ? 0 : bsx;
    o.node = bnx == INT32_MAX ? -1 : bnx;
    o.pad = 0;
    out[(size_t)p * k + round] = o;
    if (bnx == INT32_MAX) {
      for (int j = round + 1; j < k; ++j) out[(size_t)p * k + j] = o;
      return;
    }
    last_s = bsx;
    last_n = bnx;
    have_last = true;
  }
}

__global__ void fill_feasible_kernel(uint64_t* __restrict__ dst, int words, int N, int P) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P * words) return;
  const int w = i % words;
  const int base = w * 64;
  uint64_t v = 0;
  if (base + 64 <= N) v = ~0ull;
  else if (base < N) v = (1ull << (N - base)) - 1ull;
  dst[i] = v;
}

}  // namespace

int combined_eval(b200s_ctx* c, uint32_t mask, const int64_t* weights, int k, int write_total) {
  if (k < 1 || k > K_MAX) return c->set_err(B200S_ERR_INVALID, "eval_combined: k must be 1..16");
  if (!weights) return c->set_err(B200S_ERR_INVALID, "eval_combined: null weights");
  if ((mask & ((1u << B200S_PLUGIN_COUNT) - 1)) == 0) return c->set_err(B200S_ERR_INVALID, "eval_combined: no plugin enabled");
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  c->total_valid = c->topk_valid = c->feas_valid = false;
  c->mask_override = nullptr;
  if (cycle_applies(c, mask, k, write_total)) {  // a handful of pods on one GPU: the whole cycle in two launches (cycle.cu)
    for (auto& o : c->out) o.valid = false;
    return cycle_eval(c, mask, weights, k);
  }
  const uint64_t* cur = c->upstream_mask();
  struct Reset {
    b200s_ctx* c;
    ~Reset() { c->mask_override = nullptr; }
  } reset{c};
  // filters first, chained
  if (mask & (1u << B200S_PLUGIN_NRT)) {
    c->mask_override = cur;
    B200S_TRY(nrt_eval(c, B200S_OUT_U8));
    cur = c->out[B200S_PLUGIN_NRT].feas.as<uint64_t>();
  }
  if (mask & (1u << B200S_PLUGIN_NETWORK_OVERHEAD)) {
    c->mask_override = cur;
    B200S_TRY(netoh_eval(c, B200S_OUT_U8));
    cur = c->out[B200S_PLUGIN_NETWORK_OVERHEAD].feas.as<uint64_t>();
  }
  if (mask & (1u << B200S_PLUGIN_ALLOCATABLE)) {
    c->mask_override = cur;
    B200S_TRY(alloc_eval(c, B200S_OUT_U8));
  }
  if (mask & (1u << B200S_PLUGIN_PEAKS)) {  // NormalizeScore over what the filters left, like Allocatable
    c->mask_override = cur;
    B200S_TRY(peaks_eval(c, B200S_OUT_U8));
  }
  c->mask_override = nullptr;
  if (mask & (1u << B200S_PLUGIN_LOW_RISK)) B200S_TRY(lowrisk_eval(c, B200S_OUT_U8));
  if (mask & (1u << B200S_PLUGIN_TLP)) B200S_TRY(tlp_eval(c, B200S_OUT_U8));
  if (mask & (1u << B200S_PLUGIN_LVRB)) B200S_TRY(lvrb_eval(c, B200S_OUT_U8));

  B200S_CUDA_TRY(c, c->total_feas.ensure((size_t)(P > 0 ? P : 1) * words * 8));
  B200S_CUDA_TRY(c, c->topk_local.ensure((size_t)(P > 0 ? P : 1) * k * sizeof(b200s_topk_entry)));
  B200S_CUDA_TRY(c, c->topk_final.ensure((size_t)(P > 0 ? P : 1) * k * sizeof(b200s_topk_entry)));
  if (write_total) B200S_CUDA_TRY(c, c->total.ensure((size_t)(P > 0 ? P : 1) * Npad * 8));
  c->topk_k = k;
  if (P == 0) {
    c->topk_valid = true;
    c->total_valid = write_total != 0;
    return B200S_OK;
  }
  if (cur)
    B200S_CUDA_TRY(c, cudaMemcpyAsync(c->total_feas.p, cur, (size_t)P * words * 8, cudaMemcpyDeviceToDevice, c->stream));
  else {
    fill_feasible_kernel<<<(P * words + 255) / 256, 256, 0, c->stream>>>(c->total_feas.as<uint64_t>(), words, N, P);
    c->launches++;
  }
  PluginPtrs pl;
  pl.n = 0;
  for (int j = 0; j < B200S_PLUGIN_COUNT; ++j)
    if (mask & (1u << j)) {
      pl.s[pl.n] = c->out[j].scores.as<uint8_t>();
      pl.w[pl.n] = weights[j];
      pl.n++;
    }
  const int world = comm_world(c);
  b200s_topk_entry* local = world > 1 ? c->topk_local.as<b200s_topk_entry>() : c->topk_final.as<b200s_topk_entry>();
  int64_t* tot = write_total ? c->total.as<int64_t>() : nullptr;
  // node slices per pod: enough CTAs for 4 waves of the 148 SMs, slices of >= 1024 nodes (multiple of 1024)
  int S = (4 * B200S_SM_COUNT + P - 1) / P;
  S = std::max(1, std::min(S, std::min(64, (Npad + 1023) / 1024)));
  const int chunk = ((Npad + S - 1) / S + 1023) / 1024 * 1024;
  S = (Npad + chunk - 1) / chunk;
  b200s_topk_entry* stage1 = local;
  if (S > 1) {
    B200S_CUDA_TRY(c, c->topk_slices.ensure((size_t)S * P * k * sizeof(b200s_topk_entry)));
    stage1 = c->topk_slices.as<b200s_topk_entry>();
  }
  dim3 grid(P, S);
  {
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_COMBINE);  // weighted sum + per-pod top-k + slice fold
  bool small_w = true;  // sum_j w_j * 255 stays below 2^31 and no weight is negative
  {
    int64_t wsum = 0;
    for (int j = 0; j < pl.n; ++j) {
      if (pl.w[j] < 0 || pl.w[j] > (1 << 20)) small_w = false;
      wsum += pl.w[j] < 0 ? 0 : pl.w[j];
    }
    if (wsum * 255 >= ((int64_t)1 << 31)) small_w = false;
  }
  if (k == 1 && !tot && small_w)
    combine_top1_kernel<<<grid, 256, 0, c->stream>>>(pl, c->total_feas.as<uint64_t>(), words, N, Npad, c->node_off, chunk, P, stage1);
  else if (k == 1)
    combine_topk_kernel<1><<<grid, 256, 0, c->stream>>>(pl, c->total_feas.as<uint64_t>(), words, N, Npad, c->node_off, k, chunk, P, tot, stage1);
  else if (k <= 4)
    combine_topk_kernel<4><<<grid, 256, 0, c->stream>>>(pl, c->total_feas.as<uint64_t>(), words, N, Npad, c->node_off, k, chunk, P, tot, stage1);
  else
    combine_topk_kernel<K_MAX><<<grid, 256, 0, c->stream>>>(pl, c->total_feas.as<uint64_t>(), words, N, Npad, c->node_off, k, chunk, P, tot, stage1);
  c->launches++;
  if (S > 1) {
    fold_topk_kernel<<<(P + 127) / 128, 128, 0, c->stream>>>(stage1, S, P, k, local);
    c->launches++;
  }
  }
  B200S_CUDA_TRY(c, cudaGetLastError());
  if (world > 1) {
    const size_t bytes = (size_t)P * k * sizeof(b200s_topk_entry);
    B200S_CUDA_TRY(c, c->topk_all.ensure(bytes * world));
    B200S_TRY(comm_allgather(c, c->topk_local.p, c->topk_all.p, bytes));
    fold_topk_kernel<<<(P + 127) / 128, 128, 0, c->stream>>>(c->topk_all.as<b200s_topk_entry>(), world, P, k,
                                                             c->topk_final.as<b200s_topk_entry>());
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  c->topk_valid = true;
  c->total_valid = write_total != 0;
  return B200S_OK;
}

}  // namespace b200s

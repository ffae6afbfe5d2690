// Trimaran: TargetLoadPacking.Score and LoadVariationRiskBalancing.Score, all pods x all nodes.
//
// Reference semantics: float64 end to end, never fused (Go/amd64), result int64 via math.Round
//   TLP   pkg/trimaran/targetloadpacking/targetloadpacking.go:107-187
//   LVRB  pkg/trimaran/loadvariationriskbalancing/loadvariationriskbalancing.go:84-122,
//         analysis.go:34-60, pkg/trimaran/resourcestats.go:45-86
// The translation unit is compiled with -fmad=false so that a*b+c stays two roundings, and uses
// IEEE division (nvcc default -prec-div=true) — scores are bit-identical to the Go path.
//
// B200 design: the node-only sub-expressions are hoisted out of the P x N loop and live in
// registers for the whole pod tile (TLP: util%/100*cap and float(missing); LVRB: clamped
// usedAvg, capacity and the complete sigma term incl. Pow — sigma does not depend on the pod).
// Per eval what remains is one (TLP) or two (LVRB) fp64 divisions plus compares; the only HBM
// traffic is the streaming store of the score matrix.
#include <math_constants.h>

#include <cmath>

#include "engine.h"
#include "trimaran_device.cuh"

namespace b200s {

namespace {

using namespace tridev;

__global__ void div_check_kernel(const double* __restrict__ x, const double* __restrict__ d, int n,
                                 unsigned long long* __restrict__ mismatches) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double want = x[i] / d[i];
  const double got = (div_eligible(d[i]) && num_in_range(x[i])) ? div_inv(x[i], d[i], 1.0 / d[i]) : x[i] / d[i];
  if (__double_as_longlong(want) != __double_as_longlong(got) && !(want != want && got != got)) atomicAdd(mismatches, 1ull);
}

template <class OutT, int NPT, int PT>
__global__ void __launch_bounds__(256)
tlp_kernel(const double* __restrict__ util, const int64_t* __restrict__ cap, const int64_t* __restrict__ missing,
           const uint8_t* __restrict__ flags, const int64_t* __restrict__ pod_cpu, int64_t target, int N, int Npad,
           int P, OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  __shared__ double s_pod[PT];
  const int nb = blockIdx.x * CHUNK + threadIdx.x * NPT;
  const int p0 = blockIdx.y * PT;
  for (int i = threadIdx.x; i < PT; i += 256)
    if (p0 + i < P) s_pod[i] = (double)pod_cpu[p0 + i];
  double ncap[NPT], rcap[NPT], base[NPT], miss[NPT];
  uint32_t ok = 0;
  bool fast = true;  // every division this thread does satisfies the div_inv conditions (see above)
  if (nb < Npad) {
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      int n = nb + j;
      ncap[j] = (double)cap[n];
      rcap[j] = 1.0 / ncap[j];
      base[j] = (util[n] / 100) * ncap[j];  // nodeCPUUtilMillis, :147
      miss[j] = (double)missing[n];
      uint8_t f = flags[n];
      if (n < N && (f & B200S_TLP_HAS_METRICS) && (f & B200S_TLP_CPU_FOUND)) ok |= 1u << j;
      if (ncap[j] != 0 && !(div_eligible(ncap[j]) && num_in_range(base[j]))) fast = false;
    }
  }
  __syncthreads();
  if (nb >= Npad) return;
  const double t = (double)target;
  const double hundred_minus_t = 100 - t;
  const double r_t = 1.0 / t, r_hmt = 1.0 / hundred_minus_t;
  // target 0 or 100 divides by zero in one branch: the IEEE form handles it as Go does
  if (!(div_eligible(t) && div_eligible(hundred_minus_t))) fast = false;
  const int pend = min(PT, P - p0);
  OutT* orow = out + (size_t)p0 * Npad + nb;
  if (fast) {
    for (int pp = 0; pp < pend; ++pp, orow += Npad) {
      const double pc = s_pod[pp];
      int64_t q[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        double predicted = 0;
        if (ncap[j] != 0) predicted = div_inv(100 * (base[j] + pc + miss[j]), ncap[j], rcap[j]);  // :170-173
        // both branches of :174-184 are a division by a launch invariant followed by math.Round: evaluate them as ONE
        // straight-line chain with the operands selected (same operations on the same values, no divergent paths)
        const bool over = predicted > t;
        const double num = over ? t * (100 - predicted) : hundred_minus_t * predicted;
        const double d = over ? hundred_minus_t : t, r = over ? r_hmt : r_t;
        double quo = div_inv(num, d, r);
        quo = over ? quo : quo + t;                       // :183
        double s = go_round(quo);
        s = (over && predicted > 100) ? 0.0 : s;          // :175-177
        q[j] = ((ok >> j) & 1u) ? go_f2i(s) : 0;
      }
      Store<OutT, NPT>::put64(orow, q);
    }
    return;
  }
  for (int pp = 0; pp < pend; ++pp, orow += Npad) {  // the same formulas with the IEEE division
    const double pc = s_pod[pp];
    int64_t q[NPT];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      double predicted = 0;
      if (ncap[j] != 0) predicted = 100 * (base[j] + pc + miss[j]) / ncap[j];
      const bool over = predicted > t;
      const double num = over ? t * (100 - predicted) : hundred_minus_t * predicted;
      double quo = num / (over ? hundred_minus_t : t);
      quo = over ? quo : quo + t;
      double s = go_round(quo);
      s = (over && predicted > 100) ? 0.0 : s;
      q[j] = ((ok >> j) & 1u) ? go_f2i(s) : 0;
    }
    Store<OutT, NPT>::put64(orow, q);
  }
}

template <class OutT, int NPT, int PT>
__global__ void __launch_bounds__(256)
lvrb_kernel(const double* __restrict__ f64, const int64_t* __restrict__ i64, const uint8_t* __restrict__ flags,
            const int64_t* __restrict__ req_cpu, const int64_t* __restrict__ req_mem, double margin, double sens,
            int N, int Npad, int P, OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  constexpr double MEGA = 1. / 1024. / 1024.;  // resourcestats.go:29
  __shared__ double s_cpu[PT], s_mem[PT];
  const int nb = blockIdx.x * CHUNK + threadIdx.x * NPT;
  const int p0 = blockIdx.y * PT;
  for (int i = threadIdx.x; i < PT; i += 256)
    if (p0 + i < P) {
      s_cpu[i] = go_max((double)req_cpu[p0 + i], 0);         // analysis.go:41 on resourcestats.go:60
      s_mem[i] = go_max((double)req_mem[p0 + i] * MEGA, 0);  // resourcestats.go:64
    }
  LvrbNode cpu[NPT], mem[NPT];
  uint32_t fl[NPT];
  bool plain = true;  // every resource this thread scores is finite: the straight-line form applies
  if (nb < Npad) {
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      int n = nb + j;
      cpu[j] = lvrb_node(f64[n], f64[(size_t)Npad + n], (double)i64[n], margin, sens);
      double mcap = (double)i64[(size_t)Npad + n];
      mcap *= MEGA;  // resourcestats.go:62-63
      mem[j] = lvrb_node(f64[2 * (size_t)Npad + n], f64[3 * (size_t)Npad + n], mcap, margin, sens);
      fl[j] = n < N ? flags[n] : 0;
      if ((fl[j] & B200S_LVRB_CPU_OK) && !lvrb_finite(cpu[j])) plain = false;
      if ((fl[j] & B200S_LVRB_MEM_OK) && !lvrb_finite(mem[j])) plain = false;
    }
  }
  __syncthreads();
  if (nb >= Npad) return;
  const int pend = min(PT, P - p0);
  OutT* orow = out + (size_t)p0 * Npad + nb;
  if (plain) {
    for (int pp = 0; pp < pend; ++pp, orow += Npad) {
      const double rc = s_cpu[pp], rm = s_mem[pp];
      int64_t q[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        const bool cpu_ok = fl[j] & B200S_LVRB_CPU_OK, mem_ok = fl[j] & B200S_LVRB_MEM_OK;
        const double cs = cpu_ok ? lvrb_res_score_finite(cpu[j], rc) : 0.0;
        const double ms = mem_ok ? lvrb_res_score_finite(mem[j], rm) : 0.0;
        const double total = (mem_ok && cpu_ok) ? fmin(ms, cs) : fmax(ms, cs);  // :113-118, no NaN possible here
        q[j] = (fl[j] & B200S_LVRB_HAS_METRICS) ? go_f2i(go_round(total)) : 0;
      }
      Store<OutT, NPT>::put64(orow, q);
    }
    return;
  }
  for (int pp = 0; pp < pend; ++pp, orow += Npad) {
    const double rc = s_cpu[pp], rm = s_mem[pp];
    int64_t q[NPT];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const bool cpu_ok = fl[j] & B200S_LVRB_CPU_OK, mem_ok = fl[j] & B200S_LVRB_MEM_OK;
      double cs = cpu_ok ? lvrb_res_score(cpu[j], rc) : 0.0;
      double ms = mem_ok ? lvrb_res_score(mem[j], rm) : 0.0;
      double total = (mem_ok && cpu_ok) ? go_min(ms, cs) : go_max(ms, cs);  // :113-118
      q[j] = (fl[j] & B200S_LVRB_HAS_METRICS) ? go_f2i(go_round(total)) : 0;
    }
    Store<OutT, NPT>::put64(orow, q);
  }
}

}  // namespace

// Test hook: counts i with div_inv(x[i], d[i], 1/d[i]) != x[i]/d[i] (bit compare) on the device.
int debug_div_check(b200s_ctx* c, const double* x, const double* d, int n, uint64_t* mismatches) {
  DevBuf bx, bd, bm;
  B200S_CUDA_TRY(c, bx.ensure((size_t)n * 8));
  B200S_CUDA_TRY(c, bd.ensure((size_t)n * 8));
  B200S_CUDA_TRY(c, bm.ensure(8));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(bx.p, x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(bd.p, d, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemsetAsync(bm.p, 0, 8, c->stream));
  div_check_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(bx.as<double>(), bd.as<double>(), n,
                                                           bm.as<unsigned long long>());
  B200S_CUDA_TRY(c, cudaGetLastError());
  B200S_CUDA_TRY(c, cudaMemcpyAsync(mismatches, bm.p, 8, cudaMemcpyDeviceToHost, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  bx.release();
  bd.release();
  bm.release();
  return B200S_OK;
}

// Score-table path: pending pods share few distinct request keys (pods of one Deployment / Job are identical), and a
// Trimaran score depends on the pod only through that key.  The plugin's kernel runs once per DISTINCT key into a byte
// table T[key][node] (L2-resident: tens of rows), and this kernel expands it to the [P][Npad] matrix -- a streaming store
// at the HBM write rate instead of ~70-110 fp64-heavy instructions per (pod, node).  Results are the same bytes.
template <class OutT, int NPT, int PT>
__global__ void __launch_bounds__(256)
expand_rows_kernel(const uint8_t* __restrict__ T, const int32_t* __restrict__ row_of_pod, int Npad, int P, OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  __shared__ int32_t s_row[PT];
  const int nb = blockIdx.x * CHUNK + threadIdx.x * NPT, p0 = blockIdx.y * PT;
  for (int i = threadIdx.x; i < PT; i += 256)
    if (p0 + i < P) s_row[i] = row_of_pod[p0 + i];
  __syncthreads();
  if (nb >= Npad) return;
  const int pend = min(PT, P - p0);
  OutT* orow = out + (size_t)p0 * Npad + nb;
  for (int pp = 0; pp < pend; ++pp, orow += Npad) {
    const uint8_t* src = T + (size_t)s_row[pp] * Npad + nb;
    if constexpr (sizeof(OutT) == 8) {
      static_assert(sizeof(OutT) != 8 || NPT == 2, "int64 rows: two nodes per thread, one 16-byte store");
      const uint32_t v = *reinterpret_cast<const uint16_t*>(src);
      st_stream_v2(reinterpret_cast<int64_t*>(orow), (int64_t)(v & 255u), (int64_t)(v >> 8));
    } else {
      static_assert(sizeof(OutT) == 8 || NPT == 16, "u8 rows: sixteen nodes per thread, one 16-byte load and store");
      const uint4 v = *reinterpret_cast<const uint4*>(src);
      __stcs(reinterpret_cast<uint4*>(orow), v);
    }
  }
}

template <class OutT>
void launch_expand(b200s_ctx* c, const uint8_t* T, const int32_t* row, int P, OutT* out) {
  constexpr int PT = 64, NPT = sizeof(OutT) == 8 ? 2 : 16;
  dim3 grid((c->Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
  expand_rows_kernel<OutT, NPT, PT><<<grid, 256, 0, c->stream>>>(T, row, c->Npad, P, out);
  c->launches++;
}

int tlp_eval(b200s_ctx* c, int dtype) {
  if (!c->has_tlp) return c->set_err(B200S_ERR_STATE, "TargetLoadPacking: snapshot has no TLP columns");
  if (!c->has_tlp_pods) return c->set_err(B200S_ERR_STATE, "TargetLoadPacking: pod batch has no tlp_pod_cpu_milli");
  const int P = c->P, N = c->N, Npad = c->Npad;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_TLP, dtype, false, false));
  PluginOut& o = c->out[B200S_PLUGIN_TLP];
  // few distinct predicted-CPU values among the pods: one table row per value + the expansion
  const bool table = P > 0 && c->tlp_U > 0 && c->tlp_U * 2 <= P && c->tlp_sane && c->tlp_target > 0 && c->tlp_target < 100;
  if (table) {
    constexpr int PT = 64, NPT = 2;
    KernelTimer kt(c, B200S_PLUGIN_TLP);
    const int U = c->tlp_U;
    B200S_CUDA_TRY(c, c->score_table.ensure((size_t)U * Npad));
    dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (U + PT - 1) / PT);
    tlp_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->tlp_util.as<double>(), c->tlp_cap.as<int64_t>(),
                                                             c->tlp_missing.as<int64_t>(), c->tlp_flags.as<uint8_t>(),
                                                             c->tlp_uniq.as<int64_t>(), c->tlp_target, N, Npad, U,
                                                             c->score_table.as<uint8_t>());
    c->launches++;
    if (dtype == B200S_OUT_I64)
      launch_expand<int64_t>(c, c->score_table.as<uint8_t>(), c->tlp_row.as<int32_t>(), P, o.scores.as<int64_t>());
    else
      launch_expand<uint8_t>(c, c->score_table.as<uint8_t>(), c->tlp_row.as<int32_t>(), P, o.scores.as<uint8_t>());
    B200S_CUDA_TRY(c, cudaGetLastError());
  } else if (P > 0) {
    constexpr int PT = 64;
    KernelTimer kt(c, B200S_PLUGIN_TLP);
    if (dtype == B200S_OUT_I64) {
      constexpr int NPT = 2;
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
      tlp_kernel<int64_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->tlp_util.as<double>(), c->tlp_cap.as<int64_t>(), c->tlp_missing.as<int64_t>(),
          c->tlp_flags.as<uint8_t>(), c->tlp_pod_cpu.as<int64_t>(), c->tlp_target, N, Npad, P,
          o.scores.as<int64_t>());
    } else {
      constexpr int NPT = 2;  // fp64-issue bound: the narrower store does not matter, the registers do
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
      tlp_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->tlp_util.as<double>(), c->tlp_cap.as<int64_t>(), c->tlp_missing.as<int64_t>(),
          c->tlp_flags.as<uint8_t>(), c->tlp_pod_cpu.as<int64_t>(), c->tlp_target, N, Npad, P,
          o.scores.as<uint8_t>());
    }
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  o.valid = true;
  return B200S_OK;
}

int lvrb_eval(b200s_ctx* c, int dtype) {
  if (!c->has_lvrb) return c->set_err(B200S_ERR_STATE, "LoadVariationRiskBalancing: snapshot has no LVRB columns");
  if (!c->has_lvrb_pods) return c->set_err(B200S_ERR_STATE, "LoadVariationRiskBalancing: pod batch has no lvrb_req_*");
  const int P = c->P, N = c->N, Npad = c->Npad;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_LVRB, dtype, false, false));
  PluginOut& o = c->out[B200S_PLUGIN_LVRB];
  // few distinct (cpu, memory) request pairs among the pods: one table row per pair + the expansion
  const bool table = P > 0 && c->lvrb_U > 0 && c->lvrb_U * 2 <= P && c->lvrb_sane && std::isfinite(c->lvrb_margin) &&
                     std::isfinite(c->lvrb_sens);
  if (table) {
    constexpr int PT = 64, NPT = 2;
    KernelTimer kt(c, B200S_PLUGIN_LVRB);
    const int U = c->lvrb_U;
    B200S_CUDA_TRY(c, c->score_table.ensure((size_t)U * Npad));
    dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (U + PT - 1) / PT);
    lvrb_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(c->lvrb_f64.as<double>(), c->lvrb_i64.as<int64_t>(),
                                                              c->lvrb_flags.as<uint8_t>(), c->lvrb_uniq_cpu.as<int64_t>(),
                                                              c->lvrb_uniq_mem.as<int64_t>(), c->lvrb_margin, c->lvrb_sens, N, Npad,
                                                              U, c->score_table.as<uint8_t>());
    c->launches++;
    if (dtype == B200S_OUT_I64)
      launch_expand<int64_t>(c, c->score_table.as<uint8_t>(), c->lvrb_row.as<int32_t>(), P, o.scores.as<int64_t>());
    else
      launch_expand<uint8_t>(c, c->score_table.as<uint8_t>(), c->lvrb_row.as<int32_t>(), P, o.scores.as<uint8_t>());
    B200S_CUDA_TRY(c, cudaGetLastError());
  } else if (P > 0) {
    constexpr int PT = 64;
    KernelTimer kt(c, B200S_PLUGIN_LVRB);
    if (dtype == B200S_OUT_I64) {
      constexpr int NPT = 2;
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
      lvrb_kernel<int64_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->lvrb_f64.as<double>(), c->lvrb_i64.as<int64_t>(), c->lvrb_flags.as<uint8_t>(),
          c->lvrb_req_cpu.as<int64_t>(), c->lvrb_req_mem.as<int64_t>(), c->lvrb_margin, c->lvrb_sens, N, Npad, P,
          o.scores.as<int64_t>());
    } else {
      constexpr int NPT = 2;  // fp64-issue bound: the narrower store does not matter, the registers do
      dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
      lvrb_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->lvrb_f64.as<double>(), c->lvrb_i64.as<int64_t>(), c->lvrb_flags.as<uint8_t>(),
          c->lvrb_req_cpu.as<int64_t>(), c->lvrb_req_mem.as<int64_t>(), c->lvrb_margin, c->lvrb_sens, N, Npad, P,
          o.scores.as<uint8_t>());
    }
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  o.valid = true;
  return B200S_OK;
}

}  // namespace b200s

// Trimaran, second pair: Peaks.Score/NormalizeScore and LowRiskOverCommitment.Score, all pods x all nodes.
//
//   Peaks    pkg/trimaran/peaks/peaks.go:103-199
//   LowRisk  pkg/trimaran/lowriskovercommitment/lowriskovercommitment.go:105-254, beta.go:84-191,
//            pkg/trimaran/resourcestats.go:45-107, 160-228
// float64 in the reference's operation order (-fmad=false, IEEE division), results int64.
//
// B200 design.
//   Peaks: exp(k2 * util) is a property of the node and is hoisted; per eval one exp remains (Go's portable
//   math.Exp restated below -- range reduction, degree-5 minimax, ldexp -- about 25 fp64 operations).  The
//   normalisation needs the per-pod min/max over the feasible set first: pass A evaluates and reduces (warp
//   shuffle, one atomic pair per warp and pod), pass B evaluates again and stores the normalised score.
//   Recomputing costs ~30 fp64 operations per eval; a raw matrix round trip would cost 16 bytes per eval of HBM.
//   LowRisk: the measured-overcommitment risk (the beta-distribution part: two regularised incomplete beta
//   evaluations per resource) depends on the node only -- NodeRequestMinusPod, NodeLimitMinusPod, capacity, mu and
//   sigma do not involve the pending pod.  It is computed once per snapshot by a per-node kernel (N threads) and
//   kept as two f64 columns; the P x N kernel is then integer adds, at most two divisions, a weighted sum and a
//   rounding -- its only HBM traffic is the score matrix.
//
// Parity: exp is the same algorithm as the oracle's (bit-exact between the two).  The incomplete beta function
// uses CUDA's log / pow / tgamma / lgamma where the oracle uses libm's and the reference gonum's: agreement to
// ~1e-12 on the risk, identical integer scores except at a rounding boundary (tests/test_gpu_trimaran2.py states
// the tolerance).
#include <math_constants.h>

#include "engine.h"

namespace b200s {

namespace {

__device__ __forceinline__ int64_t go_f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
__device__ __forceinline__ double go_round(double x) {
  double a = fabs(x);
  if (!(a < 4503599627370496.0)) return x;
  double f = floor(a);
  if (a - f >= 0.5) f += 1.0;
  return copysign(f, x);
}
__device__ __forceinline__ double go_min(double a, double b) {
  if (a != a || b != b) return CUDART_NAN;
  return a < b ? a : b;
}
__device__ __forceinline__ double go_max(double a, double b) {
  if (a != a || b != b) return CUDART_NAN;
  return a > b ? a : b;
}

// math.Exp, Go's portable implementation (src/math/exp.go, Copyright (c) 2009 The Go Authors, BSD-style licence),
// which is a simplified version of FreeBSD's lib/msun/src/e_exp.c; algorithm and constants came with this notice:
//   ====================================================
//   Copyright (C) 2004 by Sun Microsystems, Inc. All rights reserved.
//   Permission to use, copy, modify, and distribute this software is freely granted, provided that this notice
//   is preserved.
//   ====================================================
// Restated here (not copied); full upstream notices in NOTICE.md.
__device__ __forceinline__ double go_exp(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
               Log2e = 1.44269504088896338700e+00, Overflow = 7.09782712893383973096e+02,
               Underflow = -7.45133219101941108420e+02, NearZero = 1.0 / (1 << 28);
  if (x != x || x == CUDART_INF) return x;
  if (x == -CUDART_INF) return 0;
  if (x > Overflow) return CUDART_INF;
  if (x < Underflow) return 0;
  if (-NearZero < x && x < NearZero) return 1 + x;
  int k;
  if (x < 0)
    k = (int)(Log2e * x - 0.5);
  else
    k = (int)(Log2e * x + 0.5);
  const double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
  const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double r = hi - lo, t = r * r;
  const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  const double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
  return ldexp(y, k);
}

// ------------------------------------------------------------------------------------------- Peaks
// Peaks.NormalizeScore for one score (peaks.go:152-168), lo/hi = getMinMaxScores over the feasible list
__device__ __forceinline__ int64_t peaks_norm(int64_t s, int64_t lo, int64_t hi) {
  if (lo > hi || (lo == 0 && hi == 0)) return s;  // empty list / early return :154-156 (every score is 0 then)
  double norm;
  if (hi != lo)
    norm = 100.0 * (double)wrap_sub(s, lo) / (double)wrap_sub(hi, lo);
  else
    norm = (double)wrap_sub(s, lo);
  return wrap_sub(100, go_f2i(norm));
}

template <class OutT, int NPT, int PT, bool STORE>
__global__ void __launch_bounds__(256)
peaks_kernel(const double* __restrict__ util, const int64_t* __restrict__ cap, const uint8_t* __restrict__ flags,
             const double* __restrict__ kk, const int64_t* __restrict__ pod_cpu, const uint64_t* __restrict__ feas,
             int words, int N, int Npad, int P, int64_t* __restrict__ lo, int64_t* __restrict__ hi,
             OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  __shared__ double s_pod[PT];
  __shared__ int64_t s_lo[PT], s_hi[PT];
  const int nb = blockIdx.x * CHUNK + threadIdx.x * NPT;
  const int p0 = blockIdx.y * PT;
  for (int i = threadIdx.x; i < PT; i += 256)
    if (p0 + i < P) {
      s_pod[i] = (double)pod_cpu[p0 + i];
      if constexpr (STORE) {
        s_lo[i] = lo[p0 + i];
        s_hi[i] = hi[p0 + i];
      }
    }
  double ncap[NPT], base[NPT], eutil[NPT], k1[NPT], k2[NPT];
  uint32_t ok = 0, live = 0;
  if (nb < Npad) {
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = nb + j;
      ncap[j] = (double)cap[n];              // :131
      base[j] = (util[n] / 100) * ncap[j];   // nodeCPUUtilMillis :132
      k1[j] = kk[n];
      k2[j] = kk[(size_t)Npad + n];
      eutil[j] = go_exp(k2[j] * util[n]);    // :190 second term: a property of the node
      const uint8_t f = flags[n];
      if (n < N) live |= 1u << j;
      if (n < N && (f & B200S_TLP_HAS_METRICS) && (f & B200S_TLP_CPU_FOUND)) ok |= 1u << j;
    }
  }
  __syncthreads();
  const int pend = min(PT, P - p0);
  for (int pp = 0; pp < pend; ++pp) {
    const int p = p0 + pp;
    const double pc = s_pod[pp];
    int64_t q[NPT];
    int64_t mn = INT64_MAX, mx = INT64_MIN;
    if (nb < Npad) {
      const uint64_t fw = feas ? feas[(size_t)p * words + (nb >> 6)] : ~0ull;
#pragma unroll
      for (int j = 0; j < NPT; ++j) {
        int64_t raw = 0;
        if ((ok >> j) & 1u) {
          double predicted = 0;
          if (ncap[j] != 0) predicted = 100 * (base[j] + pc) / ncap[j];  // :134-137
          if (!(predicted > 100))                                          // :138-139
            raw = go_f2i(k1[j] * (go_exp(k2[j] * predicted) - eutil[j]) * 1e15);  // :141-143, Pow(10, 15) == 1e15
        }
        const bool feasible = ((live >> j) & 1u) && ((fw >> ((nb + j) & 63)) & 1ull);
        if constexpr (STORE) {
          q[j] = feasible ? peaks_norm(raw, s_lo[pp], s_hi[pp]) : 0;
        } else if (feasible) {
          mn = raw < mn ? raw : mn;
          mx = raw > mx ? raw : mx;
        }
      }
      if constexpr (STORE) Store<OutT, NPT>::put64(out + (size_t)p * Npad + nb, q);
    }
    if constexpr (!STORE) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int64_t a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
      }
      if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(reinterpret_cast<long long*>(lo + p), (long long)mn);
        atomicMax(reinterpret_cast<long long*>(hi + p), (long long)mx);
      }
    }
  }
}

__global__ void fill_lo_hi_kernel(int64_t* lo, int64_t* hi, int P) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P) {
    lo[p] = INT64_MAX;
    hi[p] = INT64_MIN;
  }
}

// ------------------------------------------------------------------------------------- LowRisk: beta
constexpr double MACHEP = 1.11022302462515654042e-16, MAXLOG = 7.09782712893383996843e2,
                 MINLOG = -7.08396418532264106224e2, MAXGAM = 171.624376956302725, BIG = 4.503599627370496e15,
                 BIGINV = 2.22044604925031308085e-16;

__device__ double gamma_ratio(double a, double b) { return tgamma(a + b) / (tgamma(a) * tgamma(b)); }
__device__ double lbeta_neg(double a, double b) { return lgamma(a + b) - lgamma(a) - lgamma(b); }

// Cephes incbet pieces (gonum mathext.RegIncBeta): power series, the two continued fractions.  Restated from the
// published algorithm -- Cephes Math Library Release 2.3, Copyright 1984, 1995 by Stephen L. Moshier; gonum's port
// Copyright (c) 2016 The Gonum Authors, BSD-style licence.  Full upstream notices in NOTICE.md.
__device__ double incb_pseries(double a, double b, double x) {
  const double ai = 1.0 / a;
  double u = (1.0 - b) * x, v = u / (a + 1.0);
  const double t1 = v;
  double t = u, n = 2.0, s = 0.0;
  const double z = MACHEP * ai;
  while (fabs(v) > z) {
    u = (n - b) * x / n;
    t *= u;
    v = t / (a + n);
    s += v;
    n += 1.0;
  }
  s += t1;
  s += ai;
  u = a * log(x);
  if ((a + b) < MAXGAM && fabs(u) < MAXLOG) {
    t = gamma_ratio(a, b);
    s = s * t * pow(x, a);
  } else {
    t = lbeta_neg(a, b) + u + log(s);
    s = t < MINLOG ? 0.0 : exp(t);
  }
  return s;
}

__device__ double incb_cf(double a, double b, double x, bool second) {
  double k1 = a, k2 = second ? b - 1.0 : a + b, k3 = a, k4 = a + 1.0, k5 = 1.0, k6 = second ? a + b : b - 1.0,
         k7 = a + 1.0, k8 = a + 2.0;
  double pkm2 = 0.0, qkm2 = 1.0, pkm1 = 1.0, qkm1 = 1.0, ans = 1.0, r = 1.0, t;
  const double z = second ? x / (1.0 - x) : x, thresh = 3.0 * MACHEP;
  for (int n = 0; n < 300; ++n) {
    double xk = -(z * k1 * k2) / (k3 * k4);
    double pk = pkm1 + pkm2 * xk, qk = qkm1 + qkm2 * xk;
    pkm2 = pkm1, pkm1 = pk, qkm2 = qkm1, qkm1 = qk;
    xk = (z * k5 * k6) / (k7 * k8);
    pk = pkm1 + pkm2 * xk, qk = qkm1 + qkm2 * xk;
    pkm2 = pkm1, pkm1 = pk, qkm2 = qkm1, qkm1 = qk;
    if (qk != 0) r = pk / qk;
    if (r != 0) {
      t = fabs((ans - r) / r);
      ans = r;
    } else {
      t = 1.0;
    }
    if (t < thresh) return ans;
    k1 += 1.0, k3 += 2.0, k4 += 2.0, k5 += 1.0, k7 += 2.0, k8 += 2.0;
    if (second)
      k2 -= 1.0, k6 += 1.0;
    else
      k2 += 1.0, k6 -= 1.0;
    if (fabs(qk) + fabs(pk) > BIG) pkm2 *= BIGINV, pkm1 *= BIGINV, qkm2 *= BIGINV, qkm1 *= BIGINV;
    if (fabs(qk) < BIGINV || fabs(pk) < BIGINV) pkm2 *= BIG, pkm1 *= BIG, qkm2 *= BIG, qkm1 *= BIG;
  }
  return ans;
}

__device__ double incbet(double aa, double bb, double xx) {
  if (xx != xx || aa != aa || bb != bb) return CUDART_NAN;
  if (xx <= 0) return 0;
  if (xx >= 1) return 1;
  if (bb * xx <= 1.0 && xx <= 0.95) return incb_pseries(aa, bb, xx);
  bool flag = false;
  double w = 1.0 - xx, a, b, xc, x, t;
  if (xx > aa / (aa + bb)) {  // reverse a and b if x is greater than the mean
    flag = true, a = bb, b = aa, xc = xx, x = w;
  } else {
    a = aa, b = bb, xc = w, x = xx;
  }
  if (flag && b * x <= 1.0 && x <= 0.95) {
    t = incb_pseries(a, b, x);
  } else {
    double y = x * (a + b - 2.0) - (a - 1.0);  // the expansion that converges better
    if (y < 0.0)
      w = incb_cf(a, b, x, false);
    else
      w = incb_cf(a, b, x, true) / xc;
    // times x^a (1-x)^b Gamma(a+b) / (a Gamma(a) Gamma(b))
    y = a * log(x);
    t = b * log(xc);
    if ((a + b) < MAXGAM && fabs(y) < MAXLOG && fabs(t) < MAXLOG) {
      t = pow(xc, b);
      t *= pow(x, a);
      t /= a;
      t *= w;
      t *= gamma_ratio(a, b);
    } else {
      y += t + lbeta_neg(a, b);
      y += log(w / a);
      t = y < MINLOG ? 0.0 : exp(y);
    }
  }
  if (flag) t = t <= MACHEP ? 1.0 - MACHEP : 1.0 - t;
  return t;
}

struct BetaDist {
  bool valid;
  double alpha, beta;
};

// BetaDistribution.DistributionFunction (beta.go:84-90) over RegularizedIncomplete (:158-170)
__device__ double beta_cdf(const BetaDist& d, double x) {
  double p;
  if (d.alpha <= 0 || d.beta <= 0 || x < 0 || x > 1)
    p = CUDART_NAN;
  else if (x == 0)
    p = 0;
  else if (x == 1)
    p = 1;
  else
    p = incbet(d.alpha, d.beta, x);
  if (p != p || p < 0 || p > 1) p = 0;
  return p;
}

// ComputeProbability (beta.go:173-191) with MatchMoments (:105-116)
__device__ double compute_probability(double mu, double sigma, double threshold, BetaDist& d) {
  d.valid = false;
  if (mu == 0 || (sigma == 0 && mu <= threshold)) return 1;
  if (sigma == 0 && mu > threshold) return 0;
  const double m1 = mu, m2 = (sigma * sigma) + (mu * mu);
  const double variance = m2 - m1 * m1;
  if (m1 < 0 || m1 > 1 || variance < 0 || variance >= m1 * (1 - m1)) return 0;
  double temp = (m1 * (1 - m1) / variance) - 1;
  temp = go_max(temp, 4.9406564584124654e-324);  // math.SmallestNonzeroFloat64
  d.alpha = m1 * temp;
  d.beta = (1 - m1) * temp;
  d.valid = true;
  const double below = beta_cdf(d, threshold);
  return below != below ? 1 : below;
}

// riskLoad of one resource (lowriskovercommitment.go:213-249): a function of the node alone
__device__ double risk_load(bool stats_ok, double util, double sd, double capacity_f, int64_t capacity,
                            int64_t req_minus_pod, int64_t lim_minus_pod, int64_t window) {
  if (!stats_ok) return 0;
  const double used_avg = util * capacity_f / 100, used_std = sd * capacity_f / 100;  // resourcestats.go:69-70
  double mu = 0, sigma = 0;
  if (capacity_f > 0) {  // GetMuSigma :77-87 with a zero request
    mu = go_max(go_min((used_avg + 0.0) / capacity_f, 1), 0);
    sigma = go_max(go_min(used_std / capacity_f, 1), 0);
  }
  sigma *= sqrt((double)window);  // math.Pow(w, 0.5) == Sqrt(w)
  const double maxvar = (mu > 0 && mu < 1) ? mu * (1 - mu) : 0;
  sigma = go_min(sigma, sqrt(maxvar * 0.99));
  double alloc_threshold = (double)req_minus_pod / (double)capacity;
  alloc_threshold = go_min(go_max(alloc_threshold, 0), 1);
  BetaDist d;
  double alloc_prob = compute_probability(mu, sigma, alloc_threshold, d);
  if (lim_minus_pod < capacity && req_minus_pod <= lim_minus_pod) {
    const double limit_threshold = (double)lim_minus_pod / (double)capacity;
    if (limit_threshold == 0) {
      alloc_prob = 1;
    } else if (d.valid) {
      const double limit_prob = beta_cdf(d, limit_threshold);
      if (limit_prob > 0) {
        alloc_prob /= limit_prob;
        alloc_prob = go_min(go_max(alloc_prob, 0), 1);
      }
    }
  }
  return 1 - alloc_prob;
}

// f64 [4][Npad] cpuAvg cpuStd memAvg memStd; i64 [6][Npad] allocCpu allocMem nodeReqCpu nodeReqMem nodeLimCpu nodeLimMem
__global__ void lowrisk_node_kernel(const double* __restrict__ f, const int64_t* __restrict__ iv,
                                    const uint8_t* __restrict__ flags, int N, int Npad, int64_t window,
                                    double* __restrict__ load) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Npad) return;
  double lc = 0, lm = 0;
  if (n < N) {
    const size_t np = (size_t)Npad;
    const uint8_t fl = flags[n];
    const int64_t cap_cpu = iv[n], cap_mem = iv[np + n];
    const int64_t req_cpu = min(iv[2 * np + n], cap_cpu), req_mem = min(iv[3 * np + n], cap_mem);  // setMin :214-215
    const double mega = 1.0 / 1024.0 / 1024.0;
    lc = risk_load((fl & B200S_LVRB_CPU_OK) != 0, f[n], f[np + n], (double)cap_cpu, cap_cpu, req_cpu, iv[4 * np + n], window);
    lm = risk_load((fl & B200S_LVRB_MEM_OK) != 0, f[2 * np + n], f[3 * np + n], (double)cap_mem * mega, cap_mem, req_mem,
                   iv[5 * np + n], window);
  }
  load[n] = lc;
  load[(size_t)Npad + n] = lm;
}

// computeRisk minus the node-only half: riskLimit + the weighted sum (:200-206, :251-254)
// FINITE: load and w are not NaN (checked once per thread), so no NaN can arise and Go's NaN-propagating min/max
// are the hardware ones
template <bool FINITE>
__device__ __forceinline__ double total_risk(int64_t cap, int64_t node_req, int64_t node_lim, int64_t pod_req,
                                             int64_t pod_lim, double load, double w) {
  int64_t request = wrap_add(node_req, pod_req);
  const int64_t limit = wrap_add(node_lim, pod_lim);
  if (request > cap) request = cap;
  double risk_limit = 0;
  if (limit > cap) risk_limit = (double)wrap_sub(limit, cap) / (double)wrap_sub(limit, request);
  const double total = w * risk_limit + (1 - w) * load;
  if (FINITE) return fmin(fmax(total, 0.0), 1.0);
  return go_min(go_max(total, 0), 1);
}

template <class OutT, int NPT, int PT>
__global__ void __launch_bounds__(256)
lowrisk_kernel(const int64_t* __restrict__ iv, const uint8_t* __restrict__ flags, const double* __restrict__ load,
               const int64_t* __restrict__ pod /* [4][P] */, double w_cpu, double w_mem, int N, int Npad, int P,
               OutT* __restrict__ out) {
  constexpr int CHUNK = 256 * NPT;
  __shared__ int64_t s_pod[PT][4];
  const int nb = blockIdx.x * CHUNK + threadIdx.x * NPT;
  const int p0 = blockIdx.y * PT;
  for (int i = threadIdx.x; i < PT * 4; i += 256) {
    const int pp = i >> 2, k = i & 3;
    if (p0 + pp < P) s_pod[pp][k] = pod[(size_t)k * P + p0 + pp];
  }
  int64_t cap_c[NPT], cap_m[NPT], req_c[NPT], req_m[NPT], lim_c[NPT], lim_m[NPT];
  double load_c[NPT], load_m[NPT];
  uint32_t ok = 0;
  if (nb < Npad) {
    const size_t np = (size_t)Npad;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      const int n = nb + j;
      cap_c[j] = iv[n], cap_m[j] = iv[np + n];
      req_c[j] = iv[2 * np + n], req_m[j] = iv[3 * np + n];
      lim_c[j] = iv[4 * np + n], lim_m[j] = iv[5 * np + n];
      load_c[j] = load[n], load_m[j] = load[np + n];
      if (n < N && (flags[n] & B200S_LVRB_HAS_METRICS)) ok |= 1u << j;  // metrics == nil -> MinNodeScore :129-133
    }
  }
  __syncthreads();
  if (nb >= Npad) return;
  const int pend = min(PT, P - p0);
  OutT* orow = out + (size_t)p0 * Npad + nb;
  bool plain = w_cpu == w_cpu && w_mem == w_mem;
#pragma unroll
  for (int j = 0; j < NPT; ++j) plain = plain && load_c[j] == load_c[j] && load_m[j] == load_m[j];
  for (int pp = 0; pp < pend; ++pp, orow += Npad) {
    const int64_t pr_c = s_pod[pp][0], pr_m = s_pod[pp][1], pl_c = s_pod[pp][2], pl_m = s_pod[pp][3];
    const bool best_effort = pr_c == 0 && pr_m == 0 && pl_c == 0 && pl_m == 0;  // :122-127
    int64_t q[NPT];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
      int64_t s = 0;
      if (!best_effort && ((ok >> j) & 1u)) {
        double rc, rm, worst;
        if (plain) {
          rc = total_risk<true>(cap_c[j], req_c[j], lim_c[j], pr_c, pl_c, load_c[j], w_cpu);
          rm = total_risk<true>(cap_m[j], req_m[j], lim_m[j], pr_m, pl_m, load_m[j], w_mem);
          worst = fmax(rc, rm);
        } else {
          rc = total_risk<false>(cap_c[j], req_c[j], lim_c[j], pr_c, pl_c, load_c[j], w_cpu);
          rm = total_risk<false>(cap_m[j], req_m[j], lim_m[j], pr_m, pl_m, load_m[j], w_mem);
          worst = go_max(rc, rm);
        }
        const double rank = 1 - worst;   // computeRank :163
        s = go_f2i(go_round(rank * 100.0));       // :138-139
      }
      q[j] = s;
    }
    Store<OutT, NPT>::put64(orow, q);
  }
}

}  // namespace

int peaks_eval(b200s_ctx* c, int dtype) {
  if (!c->has_peaks) return c->set_err(B200S_ERR_STATE, "Peaks: snapshot has no Peaks columns");
  if (!c->has_peaks_pods) return c->set_err(B200S_ERR_STATE, "Peaks: pod batch has no peaks_pod_cpu_milli");
  const int P = c->P, N = c->N, Npad = c->Npad, words = Npad / 64;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_PEAKS, dtype, false, false));
  PluginOut& o = c->out[B200S_PLUGIN_PEAKS];
  if (P > 0) {
    constexpr int PT = 32, NPT = 2;
    B200S_CUDA_TRY(c, c->pod_lo.ensure((size_t)P * 16));  // [lo | hi] contiguous: one all-reduce when sharded
    int64_t* const lo = c->pod_lo.as<int64_t>();
    int64_t* const hi = lo + P;
    fill_lo_hi_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(lo, hi, P);
    dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
    const uint64_t* up = c->upstream_mask();
    {
      KernelTimer kt(c, B200S_PLUGIN_PEAKS);
      peaks_kernel<int64_t, NPT, PT, false><<<grid, 256, 0, c->stream>>>(
          c->peaks_util.as<double>(), c->peaks_cap.as<int64_t>(), c->peaks_flags.as<uint8_t>(), c->peaks_k.as<double>(),
          c->peaks_pod_cpu.as<int64_t>(), up, words, N, Npad, P, lo, hi, nullptr);
    }
    c->launches += 2;
    B200S_CUDA_TRY(c, cudaGetLastError());
    B200S_TRY(comm_allreduce_minmax(c, lo, hi, P));
    {
      KernelTimer kt(c, B200S_PLUGIN_PEAKS);
      if (dtype == B200S_OUT_I64)
        peaks_kernel<int64_t, NPT, PT, true><<<grid, 256, 0, c->stream>>>(
            c->peaks_util.as<double>(), c->peaks_cap.as<int64_t>(), c->peaks_flags.as<uint8_t>(), c->peaks_k.as<double>(),
            c->peaks_pod_cpu.as<int64_t>(), up, words, N, Npad, P, lo, hi, o.scores.as<int64_t>());
      else
        peaks_kernel<uint8_t, NPT, PT, true><<<grid, 256, 0, c->stream>>>(
            c->peaks_util.as<double>(), c->peaks_cap.as<int64_t>(), c->peaks_flags.as<uint8_t>(), c->peaks_k.as<double>(),
            c->peaks_pod_cpu.as<int64_t>(), up, words, N, Npad, P, lo, hi, o.scores.as<uint8_t>());
    }
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  o.valid = true;
  return B200S_OK;
}

static int lowrisk_prepare(b200s_ctx* c) {
  const uint64_t key = c->snap_serial * 1000003ull + c->lowrisk_cfg_gen;
  if (c->lowrisk_prepared_key == key) return B200S_OK;
  B200S_CUDA_TRY(c, c->lowrisk_load.ensure((size_t)c->Npad * 2 * 8));
  lowrisk_node_kernel<<<(c->Npad + 127) / 128, 128, 0, c->stream>>>(
      c->lowrisk_f64.as<double>(), c->lowrisk_i64.as<int64_t>(), c->lowrisk_flags.as<uint8_t>(), c->N, c->Npad,
      c->lowrisk_window, c->lowrisk_load.as<double>());
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  c->lowrisk_prepared_key = key;
  return B200S_OK;
}

int lowrisk_eval(b200s_ctx* c, int dtype) {
  if (!c->has_lowrisk) return c->set_err(B200S_ERR_STATE, "LowRiskOverCommitment: snapshot has no LowRisk columns");
  if (!c->has_lowrisk_pods) return c->set_err(B200S_ERR_STATE, "LowRiskOverCommitment: pod batch has no low_risk_pod");
  const int P = c->P, N = c->N, Npad = c->Npad;
  B200S_TRY(ensure_out(c, B200S_PLUGIN_LOW_RISK, dtype, false, false));
  PluginOut& o = c->out[B200S_PLUGIN_LOW_RISK];
  if (P > 0) {
    B200S_TRY(lowrisk_prepare(c));
    constexpr int PT = 64, NPT = 2;
    dim3 grid((Npad + 256 * NPT - 1) / (256 * NPT), (P + PT - 1) / PT);
    KernelTimer kt(c, B200S_PLUGIN_LOW_RISK);
    if (dtype == B200S_OUT_I64)
      lowrisk_kernel<int64_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->lowrisk_i64.as<int64_t>(), c->lowrisk_flags.as<uint8_t>(), c->lowrisk_load.as<double>(),
          c->lowrisk_pod.as<int64_t>(), c->lowrisk_w_cpu, c->lowrisk_w_mem, N, Npad, P, o.scores.as<int64_t>());
    else
      lowrisk_kernel<uint8_t, NPT, PT><<<grid, 256, 0, c->stream>>>(
          c->lowrisk_i64.as<int64_t>(), c->lowrisk_flags.as<uint8_t>(), c->lowrisk_load.as<double>(),
          c->lowrisk_pod.as<int64_t>(), c->lowrisk_w_cpu, c->lowrisk_w_mem, N, Npad, P, o.scores.as<uint8_t>());
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
  }
  o.valid = true;
  return B200S_OK;
}

}  // namespace b200s

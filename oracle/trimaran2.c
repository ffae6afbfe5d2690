/* Oracle: Trimaran Peaks + LowRiskOverCommitment.
 * TEST INFRASTRUCTURE — see oracle.h.  float64 in the reference's operation order, -ffp-contract=off.
 *
 * PARITY NOTE.  Both plugins go through transcendental functions of Go's runtime / gonum that are not in the
 * reference tree:
 *   - math.Exp (Peaks).  Restated below from Go's portable implementation (src/math/exp.go, the FreeBSD
 *     e_exp.c algorithm: k = round(x/ln2), r = hi - lo, degree-5 minimax in r^2, Ldexp).  Go's amd64 and arm64
 *     builds use an assembly variant that may differ in the last ulp; Peaks magnifies exp by 1e15, so a raw
 *     score may differ from a Go build by a few units.  The normalised 0..100 score moves only at a truncation
 *     boundary.
 *   - gonum mathext.RegIncBeta / Beta [gonum.org/v1/gonum v0.12.0, go.mod] (LowRiskOverCommitment): the Cephes
 *     incbet algorithm (power series / two continued fractions / gamma prefactor), restated below from the
 *     published Cephes description; log, pow, tgamma and lgamma come from libm.  Pinned against the reference's
 *     unit-test vectors (tests/golden/lowrisk.json: the sigma = 0 paths and beta(2,2)) and cross-checked against
 *     scipy.special.betainc to 1e-12 -- "parity partial": bit-identity with gonum cannot be claimed.
 *
 * UPSTREAM NOTICES (restatements, not copies; full texts in NOTICE.md).
 *   math.Exp: Copyright (c) 2009 The Go Authors (BSD-style licence); the algorithm and constants are FreeBSD's
 *   lib/msun/src/e_exp.c, which came with this notice:
 *     ====================================================
 *     Copyright (C) 2004 by Sun Microsystems, Inc. All rights reserved.
 *     Permission to use, copy, modify, and distribute this software is freely granted, provided that this notice
 *     is preserved.
 *     ====================================================
 *   incbet: Cephes Math Library Release 2.3, Copyright 1984, 1995 by Stephen L. Moshier; gonum's port
 *   Copyright (c) 2016 The Gonum Authors (BSD-style licence).
 */
#include <math.h>

#include "oracle.h"

static int64_t go_f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
static double go_min(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? a : b;
  return a < b ? a : b;
}
static double go_max(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  if (a == 0 && b == 0) return signbit(a) ? b : a;
  return a > b ? a : b;
}
static double go_round(double x) { return round(x); } /* math.Round: half away from zero == C round() */

/* ---- math.Exp, portable Go implementation ---- */
double orc_go_exp(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01, Ln2Lo = 1.90821492927058770002e-10,
               Log2e = 1.44269504088896338700e+00, Overflow = 7.09782712893383973096e+02,
               Underflow = -7.45133219101941108420e+02, NearZero = 1.0 / (1 << 28);
  if (isnan(x) || x == INFINITY) return x;
  if (x == -INFINITY) return 0;
  if (x > Overflow) return INFINITY;
  if (x < Underflow) return 0;
  if (-NearZero < x && x < NearZero) return 1 + x;
  int k;
  if (x < 0)
    k = (int)(Log2e * x - 0.5);
  else
    k = (int)(Log2e * x + 0.5);
  const double hi = x - (double)k * Ln2Hi, lo = (double)k * Ln2Lo;
  /* expmulti */
  const double P1 = 1.66666666666666657415e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  const double r = hi - lo, t = r * r;
  const double c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  const double y = 1 - ((lo - (r * c) / (2 - c)) - hi);
  return ldexp(y, k);
}

/* ---- Peaks ---- */
/* Peaks.Score: pkg/trimaran/peaks/peaks.go:103-146.  flags bit0 = the node has metrics (:107-111),
 * bit1 = a CPU metric with operator Average|Latest exists; util_pct = the FIRST such entry (:117-126, `break`). */
int64_t orc_peaks_score(double util_pct, int64_t cap_milli, uint8_t flags, double k1, double k2, int64_t pod_cpu_milli) {
  if (!(flags & 1)) return 0; /* metrics == nil -> MinNodeScore */
  if (!(flags & 2)) return 0; /* :127-130 */
  const double cap = (double)cap_milli;             /* :131 Status.Capacity cpu */
  const double util_millis = (util_pct / 100) * cap; /* :132 */
  double predicted = 0;
  if (cap != 0) predicted = 100 * (util_millis + (double)pod_cpu_milli) / cap; /* :134-137 */
  if (predicted > 100) return 0;                                                /* :138-139 */
  const double jump = k1 * (orc_go_exp(k2 * predicted) - orc_go_exp(k2 * util_pct)); /* :189-191 */
  return go_f2i(jump * 1e15); /* :143  math.Pow(10, 15) == 1e15 exactly */
}

/* Peaks.NormalizeScore over one pod's list, in place: peaks.go:152-168 with getMinMaxScores :170-186. */
void orc_peaks_normalize(int64_t* scores, int n) {
  int64_t mx = INT64_MIN, mn = INT64_MAX;
  for (int i = 0; i < n; ++i) {
    if (scores[i] > mx) mx = scores[i];
    if (scores[i] < mn) mn = scores[i];
  }
  if (mn == 0 && mx == 0) return;
  for (int i = 0; i < n; ++i) {
    double norm;
    if (mx != mn)
      norm = 100.0 * (double)orc_wrap_sub(scores[i], mn) / (double)orc_wrap_sub(mx, mn);
    else
      norm = (double)orc_wrap_sub(scores[i], mn);
    scores[i] = orc_wrap_sub(100, go_f2i(norm));
  }
}

void orc_peaks_batch(const double* util, const int64_t* cap, const uint8_t* flags, const double* k1, const double* k2,
                     int N, const int64_t* pod_cpu, int P, const uint64_t* feasible, int words, int64_t* out, int pitch) {
  int64_t* list = (int64_t*)__builtin_malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
  int* idx = (int*)__builtin_malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  for (int p = 0; p < P; ++p) {
    int m = 0;
    for (int n = 0; n < N; ++n) {
      out[(size_t)p * pitch + n] = 0;
      if (feasible && !((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull)) continue;
      list[m] = orc_peaks_score(util[n], cap[n], flags[n], k1[n], k2[n], pod_cpu[p]);
      idx[m++] = n;
    }
    orc_peaks_normalize(list, m);
    for (int i = 0; i < m; ++i) out[(size_t)p * pitch + idx[i]] = list[i];
  }
  __builtin_free(list);
  __builtin_free(idx);
}

/* ---- regularised incomplete beta (Cephes incbet, as gonum mathext.RegIncBeta) ---- */
#define MACHEP 1.11022302462515654042e-16
#define MAXLOG 7.09782712893383996843e2
#define MINLOG (-7.08396418532264106224e2)
#define MAXGAM 171.624376956302725
#define BIG 4.503599627370496e15
#define BIGINV 2.22044604925031308085e-16

static double gamma_ratio(double a, double b) { return tgamma(a + b) / (tgamma(a) * tgamma(b)); }
static double lbeta_neg(double a, double b) { return lgamma(a + b) - lgamma(a) - lgamma(b); }

static double pseries(double a, double b, double x) {
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

/* continued fraction #1 (incbcf) and #2 (incbd) */
static double incb_cf(double a, double b, double x, int second) {
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

/* I_x(a, b) for a, b > 0 and 0 < x < 1 */
double orc_incbet(double aa, double bb, double xx) {
  if (isnan(xx) || isnan(aa) || isnan(bb)) return NAN;
  if (xx <= 0) return 0;
  if (xx >= 1) return 1;
  int flag = 0;
  double t;
  if (bb * xx <= 1.0 && xx <= 0.95) return pseries(aa, bb, xx);
  double w = 1.0 - xx, a, b, xc, x;
  if (xx > aa / (aa + bb)) { /* reverse a and b if x is greater than the mean */
    flag = 1, a = bb, b = aa, xc = xx, x = w;
  } else {
    a = aa, b = bb, xc = w, x = xx;
  }
  if (flag == 1 && b * x <= 1.0 && x <= 0.95) {
    t = pseries(a, b, x);
    goto done;
  }
  {
    double y = x * (a + b - 2.0) - (a - 1.0); /* choose the expansion that converges better */
    if (y < 0.0)
      w = incb_cf(a, b, x, 0);
    else
      w = incb_cf(a, b, x, 1) / xc;
    /* multiply by x^a (1-x)^b Gamma(a+b) / (a Gamma(a) Gamma(b)) */
    y = a * log(x);
    t = b * log(xc);
    if ((a + b) < MAXGAM && fabs(y) < MAXLOG && fabs(t) < MAXLOG) {
      t = pow(xc, b);
      t *= pow(x, a);
      t /= a;
      t *= w;
      t *= gamma_ratio(a, b);
      goto done;
    }
    y += t + lbeta_neg(a, b);
    y += log(w / a);
    t = y < MINLOG ? 0.0 : exp(y);
  }
done:
  if (flag == 1) t = t <= MACHEP ? 1.0 - MACHEP : 1.0 - t;
  return t;
}

/* ---- LowRiskOverCommitment ---- */
typedef struct {
  int valid;
  double alpha, beta;
} beta_dist;

/* BetaDistribution.DistributionFunction: beta.go:84-90 with RegularizedIncomplete :158-170 */
static double beta_cdf(const beta_dist* d, double x) {
  double p;
  if (d->alpha <= 0 || d->beta <= 0 || x < 0 || x > 1)
    p = NAN;
  else if (x == 0)
    p = 0;
  else if (x == 1)
    p = 1;
  else
    p = orc_incbet(d->alpha, d->beta, x);
  if (isnan(p) || p < 0 || p > 1) p = 0;
  return p;
}

/* ComputeProbability: beta.go:173-191 (NewBetaDistribution(1,1) then MatchMoments :105-116) */
static double compute_probability(double mu, double sigma, double threshold, beta_dist* d) {
  d->valid = 0;
  if (mu == 0 || (sigma == 0 && mu <= threshold)) return 1;
  if (sigma == 0 && mu > threshold) return 0;
  const double m1 = mu, m2 = (sigma * sigma) + (mu * mu);
  const double variance = m2 - m1 * m1;
  if (m1 < 0 || m1 > 1 || variance < 0 || variance >= m1 * (1 - m1)) return 0; /* MatchMoments false */
  double temp = (m1 * (1 - m1) / variance) - 1;
  temp = go_max(temp, 4.9406564584124654e-324); /* math.SmallestNonzeroFloat64 */
  d->alpha = m1 * temp;
  d->beta = (1 - m1) * temp;
  d->valid = 1;
  const double below = beta_cdf(d, threshold);
  if (isnan(below)) return 1;
  return below;
}

/* The measured-overcommitment half of computeRisk (lowriskovercommitment.go:213-249): depends on the node only.
 * util/std = GetResourceData of the node's metrics; capacity_f = CreateResourceStats' Capacity (cpu: allocatable
 * milli; memory: allocatable bytes * MegaFactor, resourcestats.go:58-66); capacity = NodeRequestsAndLimits.Nodecapacity
 * (cpu milli / memory bytes); req_minus_pod is ALREADY capped by capacity (resourcestats.go:214-215). */
double orc_lowrisk_risk_load(int stats_ok, double util, double std, double capacity_f, int64_t capacity,
                             int64_t req_minus_pod, int64_t lim_minus_pod, int64_t window) {
  if (!stats_ok) return 0; /* riskLoad keeps its zero value, :213-216 */
  /* CreateResourceStats with a zero pod request, then GetMuSigma (resourcestats.go:69-71, 77-87) */
  const double used_avg = util * capacity_f / 100, used_std = std * capacity_f / 100;
  double mu = 0, sigma = 0;
  if (capacity_f > 0) {
    mu = go_max(go_min((used_avg + 0.0) / capacity_f, 1), 0);
    sigma = go_max(go_min(used_std / capacity_f, 1), 0);
  }
  sigma *= sqrt((double)window); /* math.Pow(w, 0.5) is Sqrt(w) in Go */
  const double maxvar = (mu > 0 && mu < 1) ? mu * (1 - mu) : 0; /* GetMaxVariance beta.go:119-124 */
  sigma = go_min(sigma, sqrt(maxvar * 0.99));
  double alloc_threshold = (double)req_minus_pod / (double)capacity;
  alloc_threshold = go_min(go_max(alloc_threshold, 0), 1);
  beta_dist d;
  double alloc_prob = compute_probability(mu, sigma, alloc_threshold, &d);
  if (lim_minus_pod < capacity && req_minus_pod <= lim_minus_pod) { /* :231-243 */
    const double limit_threshold = (double)lim_minus_pod / (double)capacity;
    if (limit_threshold == 0) {
      alloc_prob = 1;
    } else if (d.valid) {
      const double limit_prob = beta_cdf(&d, limit_threshold);
      if (limit_prob > 0) {
        alloc_prob /= limit_prob;
        alloc_prob = go_min(go_max(alloc_prob, 0), 1);
      }
    }
  }
  return 1 - alloc_prob;
}

/* computeRisk for one resource: lowriskovercommitment.go:172-254.  node_req / node_lim = sums over the pods already
 * on the node (SetMaxLimits applied per pod, resourcestats.go:186-200); pod_req / pod_lim = the pending pod's. */
double orc_lowrisk_compute_risk(int stats_ok, double util, double std, double capacity_f, int64_t capacity,
                                int64_t node_req, int64_t node_lim, int64_t pod_req, int64_t pod_lim, int64_t window,
                                double weight) {
  int64_t request = orc_wrap_add(node_req, pod_req), limit = orc_wrap_add(node_lim, pod_lim);
  int64_t req_minus_pod = node_req;
  if (request > capacity) request = capacity;             /* setMin, :212-215 */
  if (req_minus_pod > capacity) req_minus_pod = capacity;
  double risk_limit = 0;
  if (limit > capacity) risk_limit = (double)orc_wrap_sub(limit, capacity) / (double)orc_wrap_sub(limit, request);
  const double risk_load = orc_lowrisk_risk_load(stats_ok, util, std, capacity_f, capacity, req_minus_pod, node_lim, window);
  double total = weight * risk_limit + (1 - weight) * risk_load;
  return go_min(go_max(total, 0), 1);
}

/* LowRiskOverCommitment.Score: lowriskovercommitment.go:105-141 + computeRank :155-168.
 * flags: bit0 node has metrics, bit1 CPU data valid, bit2 memory data valid (GetResourceData). */
int64_t orc_lowrisk_score(double cpu_avg, double cpu_std, double mem_avg, double mem_std, int64_t alloc_cpu_milli,
                          int64_t alloc_mem_bytes, uint8_t flags, int64_t node_req_cpu, int64_t node_req_mem,
                          int64_t node_lim_cpu, int64_t node_lim_mem, int64_t pod_req_cpu, int64_t pod_req_mem,
                          int64_t pod_lim_cpu, int64_t pod_lim_mem, int64_t window, double w_cpu, double w_mem) {
  if (pod_req_cpu == 0 && pod_req_mem == 0 && pod_lim_cpu == 0 && pod_lim_mem == 0) return 0; /* best effort, :122-127 */
  if (!(flags & 1)) return 0;                                                                   /* :129-133 */
  const double mega = 1.0 / 1024.0 / 1024.0;
  const double risk_cpu = orc_lowrisk_compute_risk((flags & 2) != 0, cpu_avg, cpu_std, (double)alloc_cpu_milli,
                                                   alloc_cpu_milli, node_req_cpu, node_lim_cpu, pod_req_cpu,
                                                   pod_lim_cpu, window, w_cpu);
  const double risk_mem = orc_lowrisk_compute_risk((flags & 4) != 0, mem_avg, mem_std, (double)alloc_mem_bytes * mega,
                                                   alloc_mem_bytes, node_req_mem, node_lim_mem, pod_req_mem,
                                                   pod_lim_mem, window, w_mem);
  const double rank = 1 - go_max(risk_cpu, risk_mem);
  return go_f2i(go_round(rank * 100.0));
}

void orc_lowrisk_batch(const double* cpu_avg, const double* cpu_std, const double* mem_avg, const double* mem_std,
                       const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags,
                       const int64_t* node_req_cpu, const int64_t* node_req_mem, const int64_t* node_lim_cpu,
                       const int64_t* node_lim_mem, int N, const int64_t* pod_req_cpu, const int64_t* pod_req_mem,
                       const int64_t* pod_lim_cpu, const int64_t* pod_lim_mem, int P, int64_t window, double w_cpu,
                       double w_mem, int64_t* out, int pitch) {
  for (int p = 0; p < P; ++p)
    for (int n = 0; n < N; ++n)
      out[(size_t)p * pitch + n] =
          orc_lowrisk_score(cpu_avg[n], cpu_std[n], mem_avg[n], mem_std[n], alloc_cpu[n], alloc_mem[n], flags[n],
                            node_req_cpu[n], node_req_mem[n], node_lim_cpu[n], node_lim_mem[n], pod_req_cpu[p],
                            pod_req_mem[p], pod_lim_cpu[p], pod_lim_mem[p], window, w_cpu, w_mem);
}

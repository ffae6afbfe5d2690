/* Oracle: Trimaran TargetLoadPacking + LoadVariationRiskBalancing.
 * TEST INFRASTRUCTURE — see oracle.h.  All arithmetic is float64 in the reference's operation
 * order; compile with -ffp-contract=off (Go on amd64 never emits FMA for a*b+c). */
#include <math.h>

#include "oracle.h"

/* Go's float64 -> int64 conversion on amd64 (CVTTSD2SI): truncates; NaN and out-of-range give
 * the "integer indefinite" value MinInt64. */
static int64_t go_f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}

/* Go builtin min/max on float64 (Go 1.21+): NaN-propagating, -0 < +0. */
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

/* math.Pow special cases that Go resolves without its general algorithm [Go 1.25 runtime,
 * not in the reference tree]: y==0|x==1 -> 1, y==1 -> x, y==0.5 -> Sqrt, y==-0.5 -> 1/Sqrt,
 * y==+-Inf.  Everything else falls to libm pow(): PARITY UNPINNED for those exponents
 * (SURVEY §8c rule ii: equal or off by one only at a rounding boundary). */
static double go_pow(double x, double y) {
  if (y == 0 || x == 1) return 1;
  if (y == 1) return x;
  if (isnan(x) || isnan(y)) return NAN;
  if (isinf(y) && x != 0) {
    if (x == -1) return 1;
    if ((fabs(x) < 1) == (y > 0)) return 0;
    return INFINITY;
  }
  if (y == 0.5 && !isinf(x) && x != 0) return sqrt(x);
  if (y == -0.5 && !isinf(x) && x != 0) return 1 / sqrt(x);
  return pow(x, y);
}

/* TargetLoadPacking.Score: targetloadpacking.go:107-187.
 * flags bit0 = metrics present (line 114), bit1 = CPU metric (Average|Latest) found (131-145). */
int64_t orc_tlp_score(double cpu_util_pct, int64_t cap_milli, int64_t missing_milli, uint8_t flags,
                      int64_t pod_cpu_milli, int64_t target_pct) {
  if (!(flags & 1)) return 0; /* metrics == nil -> MinNodeScore (114-120) */
  if (!(flags & 2)) return 0; /* !cpuMetricFound (142-145) */
  double t = (double)target_pct;
  double node_cap = (double)cap_milli;                  /* 146: Status.Capacity, not Allocatable */
  double node_util_millis = (cpu_util_pct / 100) * node_cap; /* 147 */
  double predicted = 0;
  if (node_cap != 0) /* 170-173 */
    predicted = 100 * (node_util_millis + (double)pod_cpu_milli + (double)missing_milli) / node_cap;
  if (predicted > t) { /* 174 */
    if (predicted > 100) return 0; /* 175-177 */
    return go_f2i(round(t * (100 - predicted) / (100 - t))); /* 178 */
  }
  return go_f2i(round((100 - t) * predicted / t + t)); /* 183-184 */
}

void orc_tlp_batch(const double* util, const int64_t* cap, const int64_t* missing, const uint8_t* flags, int N,
                   const int64_t* pod_cpu, int P, int64_t target, int64_t* out, int pitch) {
  for (int p = 0; p < P; ++p) {
    for (int n = 0; n < N; ++n)
      out[(size_t)p * pitch + n] = orc_tlp_score(util[n], cap[n], missing[n], flags[n], pod_cpu[p], target);
    for (int n = N; n < pitch; ++n) out[(size_t)p * pitch + n] = 0;
  }
}

/* GetMuSigma: resourcestats.go:77-86. */
void orc_lvrb_mu_sigma(double used_avg, double used_std, double req, double capacity, double* mu_out,
                       double* sigma_out) {
  if (capacity <= 0) {
    *mu_out = 0;
    *sigma_out = 0;
    return;
  }
  double mu = (used_avg + req) / capacity;
  mu = go_max(go_min(mu, 1), 0);
  double sigma = used_std / capacity;
  sigma = go_max(go_min(sigma, 1), 0);
  *mu_out = mu;
  *sigma_out = sigma;
}

/* computeScore: analysis.go:34-60. */
double orc_lvrb_compute_score(double used_avg, double used_std, double req, double capacity, double margin,
                              double sensitivity) {
  if (capacity <= 0) return 0; /* 35-38 */
  req = go_max(req, 0);                                   /* 41 */
  used_avg = go_max(go_min(used_avg, capacity), 0);       /* 42 */
  used_std = go_max(go_min(used_std, capacity), 0);       /* 43 */
  double mu, sigma;
  orc_lvrb_mu_sigma(used_avg, used_std, req, capacity, &mu, &sigma); /* 46 */
  if (sensitivity >= 0) sigma = go_pow(sigma, 1 / sensitivity); /* 49-51 */
  sigma *= margin;                                              /* 53 */
  sigma = go_max(go_min(sigma, 1), 0);                          /* 54 */
  double risk = (mu + sigma) / 2;                               /* 57 */
  return (1. - risk) * 100.0;                                   /* 59: float64(fwk.MaxNodeScore) */
}

#define ORC_MEGA (1. / 1024. / 1024.) /* resourcestats.go:29 */

/* LoadVariationRiskBalancing.Score: loadvariationriskbalancing.go:84-122 with
 * CreateResourceStats resourcestats.go:45-74.  flags bit0 metrics present, bit1 cpuOK, bit2 memOK. */
int64_t orc_lvrb_score(double cpu_avg, double cpu_std, double mem_avg, double mem_std, int64_t alloc_cpu_milli,
                       int64_t alloc_mem_bytes, uint8_t flags, int64_t req_cpu_milli, int64_t req_mem_bytes,
                       double margin, double sensitivity) {
  if (!(flags & 1)) return 0; /* 91-94 */
  int cpu_ok = (flags & 2) != 0, mem_ok = (flags & 4) != 0;
  double cpu_score = 0, mem_score = 0;
  if (cpu_ok) {
    double cap = (double)alloc_cpu_milli; /* resourcestats.go:58-60 */
    double req = (double)req_cpu_milli;
    double used_avg = cpu_avg * cap / 100; /* 68 */
    double used_std = cpu_std * cap / 100; /* 69 */
    cpu_score = orc_lvrb_compute_score(used_avg, used_std, req, cap, margin, sensitivity);
  }
  if (mem_ok) {
    double cap = (double)alloc_mem_bytes; /* 62 */
    cap *= ORC_MEGA;                      /* 63 */
    double req = (double)req_mem_bytes * ORC_MEGA; /* 64 */
    double used_avg = mem_avg * cap / 100;
    double used_std = mem_std * cap / 100;
    mem_score = orc_lvrb_compute_score(used_avg, used_std, req, cap, margin, sensitivity);
  }
  double total;
  if (mem_ok && cpu_ok) /* 113-118 */
    total = go_min(mem_score, cpu_score);
  else
    total = go_max(mem_score, cpu_score);
  return go_f2i(round(total)); /* 119 */
}

void orc_lvrb_batch(const double* cpu_avg, const double* cpu_std, const double* mem_avg, const double* mem_std,
                    const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags, int N,
                    const int64_t* req_cpu, const int64_t* req_mem, int P, double margin, double sens,
                    int64_t* out, int pitch) {
  for (int p = 0; p < P; ++p) {
    for (int n = 0; n < N; ++n)
      out[(size_t)p * pitch + n] = orc_lvrb_score(cpu_avg[n], cpu_std[n], mem_avg[n], mem_std[n], alloc_cpu[n],
                                                  alloc_mem[n], flags[n], req_cpu[p], req_mem[p], margin, sens);
    for (int n = N; n < pitch; ++n) out[(size_t)p * pitch + n] = 0;
  }
}

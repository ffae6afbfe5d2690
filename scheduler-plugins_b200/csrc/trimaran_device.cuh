// Device helpers of the Trimaran plugins (Go float64 semantics, the reciprocal division with its hoisted eligibility
// checks, the node-only part of LoadVariationRiskBalancing's computeScore).  Shared by trimaran.cu and the fused
// single-cycle kernel (cycle.cu); moved here verbatim from trimaran.cu.  Include only from translation units built
// with -fmad=false.
#pragma once
#include <math_constants.h>

#include "engine.h"

namespace b200s {
namespace tridev {

// Go float64 -> int64 (CVTTSD2SI): NaN / out of range -> MinInt64.
__device__ __forceinline__ int64_t go_f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}
// math.Round: half away from zero.  floor + exact remainder test (Sterbenz) instead of libdevice round().
__device__ __forceinline__ double go_round(double x) {
  double a = fabs(x);
  if (!(a < 4503599627370496.0)) return x;  // already integral, Inf or NaN
  double f = floor(a);
  if (a - f >= 0.5) f += 1.0;
  return copysign(f, x);
}
// Go builtin min/max: NaN-propagating (signed zeros cannot change a rounded score).
__device__ __forceinline__ double go_min(double a, double b) {
  if (a != a || b != b) return CUDART_NAN;
  return a < b ? a : b;
}
__device__ __forceinline__ double go_max(double a, double b) {
  if (a != a || b != b) return CUDART_NAN;
  return a > b ? a : b;
}
// math.Pow as Go resolves it for the exponents 1/sensitivity (special cases first; see oracle).
static __device__ double go_pow(double x, double y) {
  if (y == 0 || x == 1) return 1;
  if (y == 1) return x;
  if (x != x || y != y) return CUDART_NAN;
  if (isinf(y) && x != 0) {
    if (x == -1) return 1;
    if ((fabs(x) < 1) == (y > 0)) return 0;
    return CUDART_INF;
  }
  if (y == 0.5 && !isinf(x) && x != 0) return sqrt(x);
  if (y == -0.5 && !isinf(x) && x != 0) return 1 / sqrt(x);
  return pow(x, y);  // general exponent: <= 2 ulp from Go's software Pow; tolerance rule SURVEY §8c(ii)
}

// x / d for a divisor whose correctly rounded reciprocal r = RN(1/d) is hoisted out of the pod loop:
//   q0 = RN(x*r);  rem = x - q0*d (exact in one FMA);  q1 = RN(q0 + rem*r).
// When is q1 the correctly rounded quotient?  Markstein's theorem needs q0 FAITHFUL (one of the two neighbours of
// x/d), and RN(x * RN(1/d)) can be 1.5 ulp off, so it does not hold for every divisor.  Brisebarre, Muller & Raina
// ("Accelerating correctly rounded floating-point division when the divisor is known in advance", IEEE TC 53(8),
// 2004, Theorem 4, on exactly this 1 multiplication + 2 FMA sequence) give sufficient conditions on the DIVISOR alone;
// the first: the last bit of d's significand is 0.  Every divisor on this path is an integer below 2^52 converted to
// float64 (capacity in milli-cores, targetloadpacking.go:146; the target utilisation and its complement, :174-184;
// allocatable milli-cores, resourcestats.go:55) or such an integer times 2^-20 (memory in MiB, resourcestats.go:60-64):
// all have a zero last significand bit.  The theorem also assumes that nothing overflows or underflows.  Both
// conditions are checked ONCE per divisor / per node, hoisted out of the pod loop (div_eligible, num_in_range): a
// thread whose nodes pass runs the straight-line div_inv form, any other thread runs the same formulas with the
// IEEE division -- so an infinite or denormal metric, a capacity with 53 significant bits, or a target of 0 or 100
// behave exactly as Go's x/d.  b200s_debug_div_check compares the two on the device (eligible divisors through
// div_inv, the rest through the division itself), on random operands and on operands constructed at rounding
// boundaries of the quotient (tests/test_gpu_divcheck.py).
__device__ __forceinline__ bool div_eligible(double d) {
  const long long b = __double_as_longlong(d);
  const unsigned e = (unsigned)(b >> 52) & 0x7ffu;
  return !(b & 1) && e - 923u <= 200u;  // last significand bit 0, 2^-100 <= |d| < 2^101
}
// a numerator term that is 0 or of moderate magnitude: sums / products with the int64-derived pod terms then stay
// either exactly 0 or inside [2^-300, 2^300], where neither the quotient nor the residual can underflow or overflow
__device__ __forceinline__ bool num_in_range(double x) {
  const unsigned e = ((unsigned)__double2hiint(x) >> 20) & 0x7ffu;
  // +0 (the FMA steps would turn -0 / d into +0), or 2^-200 <= |x| < 2^201 (excludes NaN, Inf, denormals)
  return __double_as_longlong(x) == 0 || e - 823u <= 400u;
}
__device__ __forceinline__ double div_inv(double x, double d, double r) {
  const double q0 = x * r;
  const double rem = __fma_rn(-q0, d, x);
  return __fma_rn(rem, r, q0);
}

struct LvrbNode {
  double avg, cap, rcap, sigma;  // clamped usedAvg, capacity, RN(1/capacity), final sigma (after Pow, margin, clamp)
};

// Node-only part of computeScore (analysis.go:34-54) for one resource.
__device__ __forceinline__ LvrbNode lvrb_node(double util_avg, double util_std, double cap, double margin,
                                              double sens) {
  LvrbNode r;
  r.cap = cap;
  r.rcap = 1.0 / cap;
  double used_avg = util_avg * cap / 100;  // resourcestats.go:68
  double used_std = util_std * cap / 100;  // :69
  r.avg = go_max(go_min(used_avg, cap), 0);
  used_std = go_max(go_min(used_std, cap), 0);
  double sigma = 0;
  if (cap > 0) {
    sigma = used_std / cap;
    sigma = go_max(go_min(sigma, 1), 0);
    if (sens >= 0) sigma = go_pow(sigma, 1 / sens);
    sigma *= margin;
    sigma = go_max(go_min(sigma, 1), 0);
  }
  r.sigma = sigma;
  return r;
}

__device__ __forceinline__ double lvrb_res_score(const LvrbNode& nd, double req) {
  if (nd.cap <= 0) return 0;  // analysis.go:35-38
  double mu = (nd.avg + req) / nd.cap;
  mu = go_max(go_min(mu, 1), 0);
  double risk = (mu + nd.sigma) / 2;
  return (1. - risk) * 100.0;
}
// Same, for a node whose avg / sigma / capacity are finite and satisfy the div_inv conditions (checked once per
// node): mu cannot be NaN, so the NaN-propagating clamps of Go's builtin min/max reduce to the hardware min/max.
__device__ __forceinline__ double lvrb_res_score_finite(const LvrbNode& nd, double req) {
  const double mu = fmax(fmin(div_inv(nd.avg + req, nd.cap, nd.rcap), 1.0), 0.0);
  return (1. - (mu + nd.sigma) / 2) * 100.0;
}
__device__ __forceinline__ bool lvrb_finite(const LvrbNode& nd) {
  return nd.cap > 0 && div_eligible(nd.cap) && num_in_range(nd.avg) && nd.sigma == nd.sigma;
}

}  // namespace tridev
}  // namespace b200s

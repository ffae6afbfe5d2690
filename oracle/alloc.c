/* Oracle: NodeResourcesAllocatable.  TEST INFRASTRUCTURE — see oracle.h. */
#include "oracle.h"

#include <stdlib.h>

/* Go int64 arithmetic wraps (two's complement); C signed overflow is undefined, so go through
 * uint64_t.  Go `/` truncates toward zero; MinInt64 / -1 == MinInt64 in Go (no trap). */
int64_t orc_wrap_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
int64_t orc_wrap_sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
int64_t orc_wrap_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
int64_t orc_go_div(int64_t x, int64_t y) {
  if (y == 0) return 0; /* Go panics; unreachable for validated args (weights > 0) */
  if (y == -1) return (int64_t)(0ull - (uint64_t)x);
  return x / y;
}

/* score(): allocatable.go:130-140 — Least: -1 * capacity, Most: capacity, otherwise 0. */
static int64_t mode_score(int64_t capacity, int mode) {
  if (mode == 0) return orc_wrap_mul(-1, capacity);
  if (mode == 1) return capacity;
  return 0;
}

/* resourceScorer closure: allocatable.go:117-128.  `requested` is ignored by the reference
 * (line 122 reads only allocable[resource]); map iteration order is irrelevant because
 * wrapping addition commutes. */
int64_t orc_alloc_score(const int64_t* alloc, const int64_t* w, int R, int mode) {
  int64_t node_score = 0, weight_sum = 0;
  for (int r = 0; r < R; ++r) {
    int64_t resource_score = mode_score(alloc[r], mode);
    node_score = orc_wrap_add(node_score, orc_wrap_mul(resource_score, w[r]));
    weight_sum = orc_wrap_add(weight_sum, w[r]);
  }
  return orc_go_div(node_score, weight_sum);
}

/* NormalizeScore: allocatable.go:143-168. */
void orc_alloc_normalize(int64_t* scores, int n) {
  int64_t highest = -INT64_MAX; /* -math.MaxInt64, line 145 */
  int64_t lowest = INT64_MAX;   /* line 146 */
  for (int i = 0; i < n; ++i) {
    if (scores[i] > highest) highest = scores[i];
    if (scores[i] < lowest) lowest = scores[i];
  }
  int64_t old_range = orc_wrap_sub(highest, lowest);
  const int64_t new_range = 100 - 0; /* fwk.MaxNodeScore - fwk.MinNodeScore */
  for (int i = 0; i < n; ++i) {
    if (old_range == 0)
      scores[i] = 0;
    else
      scores[i] = orc_wrap_add(orc_go_div(orc_wrap_mul(orc_wrap_sub(scores[i], lowest), new_range), old_range), 0);
  }
}

/* The upstream cycle for one pod: Score on every feasible node (allocatable.go:63 ->
 * resource_allocation.go:49-76), then NormalizeScore on that list. */
void orc_alloc_batch(const int64_t* const* cols, int R, int N, const int64_t* w, int mode, int P,
                     const uint64_t* feasible, int words, int64_t* out, int pitch) {
  int64_t* list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
  int* idx = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  int64_t vals[64];
  for (int p = 0; p < P; ++p) {
    int m = 0;
    for (int n = 0; n < N; ++n) {
      out[(size_t)p * pitch + n] = 0;
      if (feasible && !((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull)) continue;
      for (int r = 0; r < R; ++r) vals[r] = cols[r][n];
      list[m] = orc_alloc_score(vals, w, R, mode);
      idx[m++] = n;
    }
    orc_alloc_normalize(list, m);
    for (int i = 0; i < m; ++i) out[(size_t)p * pitch + idx[i]] = list[i];
    for (int n = N; n < pitch; ++n) out[(size_t)p * pitch + n] = 0;
  }
  free(list);
  free(idx);
}

/* Plain-C use of libb200sched.so: one scheduling cycle of NodeResourcesAllocatable (mode Most) for one pod over four
 * nodes -- the call sequence a cgo binding makes.  Build and run on a B200 box:
 *   gcc -std=c11 -Iinclude examples/cycle.c -Lscheduler-plugins_b200/lib -lb200sched \
 *       -Wl,-rpath,$PWD/scheduler-plugins_b200/lib -o /tmp/cycle && /tmp/cycle
 * tests/test_abi.py compiles this file with -fsyntax-only -pedantic -Werror: the header must stay plain C. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200sched.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != B200S_OK) {                                                   \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc_, b200s_last_error(ctx)); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(void) {
  b200s_ctx* ctx = NULL;
  if (b200s_init(0, &ctx) != B200S_OK) {
    fprintf(stderr, "no usable GPU: %s\n", b200s_last_error(NULL));
    return 2; /* there is no CPU fallback */
  }
  /* snapshot: allocatable cpu (milli) and memory (bytes) of four nodes */
  const int64_t cpu[4] = {4000, 8000, 16000, 2000};
  const int64_t mem[4] = {8ll << 30, 16ll << 30, 64ll << 30, 4ll << 30};
  const int64_t* cols[2] = {cpu, mem};
  const int64_t weights[2] = {1 << 20, 1}; /* defaults of the plugin: cpu 1<<20, memory 1 */
  CHECK(b200s_snapshot_begin(ctx, 1, 4, 0, 4));
  CHECK(b200s_snapshot_allocatable(ctx, 2, cols));
  CHECK(b200s_snapshot_commit(ctx));
  CHECK(b200s_config_allocatable(ctx, B200S_ALLOC_MOST, 2, weights));
  /* one pending pod; upstream's filters left nodes 0, 1 and 2 */
  const int32_t npad = b200s_npad(ctx);
  uint64_t* feasible = (uint64_t*)calloc((size_t)npad / 64, sizeof(uint64_t));
  uint8_t* scores = (uint8_t*)calloc((size_t)npad, 1);
  if (!feasible || !scores) return 3;
  feasible[0] = 0x7;
  b200s_pod_batch batch;
  memset(&batch, 0, sizeof(batch));
  batch.n_pods = 1;
  batch.feasible = feasible;
  CHECK(b200s_score_batch(ctx, B200S_PLUGIN_ALLOCATABLE, &batch, B200S_OUT_U8, scores, NULL, NULL));
  for (int n = 0; n < 4; ++n) printf("node %d: %d\n", n, scores[n]); /* 0, 17, 100, 0 (node 3 is infeasible) */
  /* a node event: node 0 grows; only that row travels */
  const int32_t idx[1] = {0};
  const int64_t cpu0[1] = {32000}, mem0[1] = {128ll << 30};
  const int64_t* row[2] = {cpu0, mem0};
  CHECK(b200s_snapshot_patch_begin(ctx, 2));
  CHECK(b200s_snapshot_patch_allocatable(ctx, 1, idx, 2, row));
  CHECK(b200s_snapshot_commit(ctx));
  CHECK(b200s_score_batch(ctx, B200S_PLUGIN_ALLOCATABLE, &batch, B200S_OUT_U8, scores, NULL, NULL));
  for (int n = 0; n < 4; ++n) printf("node %d: %d\n", n, scores[n]); /* 100, 0, 41, 0 */
  /* an engine-only profile does not need the row at all: upload, the whole cycle (two kernels, one graph launch), the
   * winner back -- one call, one synchronisation */
  int64_t profile[B200S_PLUGIN_COUNT];
  memset(profile, 0, sizeof(profile));
  profile[B200S_PLUGIN_ALLOCATABLE] = 1;
  b200s_topk_entry winner;
  CHECK(b200s_schedule_batch(ctx, &batch, 1u << B200S_PLUGIN_ALLOCATABLE, profile, 1, &winner));
  printf("winner: node %d, score %lld\n", (int)winner.node, (long long)winner.score); /* node 0, score 100 */
  free(feasible);
  free(scores);
  b200s_shutdown(ctx);
  return 0;
}

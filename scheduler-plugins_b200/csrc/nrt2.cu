// NodeResourceTopologyMatch, batched path: Filter + Score (Least/Most/BalancedAllocation) for P pods x N nodes.
//
// Reference semantics (pkg/noderesourcetopology): Filter filter.go:176-225 with the handlers :39-78 / :162-173 and
// resourcesAvailableInAnyNUMANodes :90-160; Score score.go:62-165 with least_allocated.go:25-55,
// most_allocated.go:25-54, balanced_allocation.go:27-54.  nrt.cu evaluates every (pod, node) pair from scratch
// (the shape of the reference); this file exploits what a BATCH of pending pods has in common:
//
//  * score_each_numa (score.go:110-124) depends on ONE request vector and the node only -- container scope averages it
//    over the pod's containers without subtraction (score.go:152-165), pod scope takes it on the pod-effective
//    request (:142-150).  Pending pods share few distinct request vectors (c4: 10 504 container instances of the
//    Guaranteed pods, 1 216 distinct), so a byte table T[vector][node] is built once (nrt2_table_kernel) and the
//    P x N pass gathers from it.  The pod-scope Filter (filter.go:162-173) is a function of (effective vector, node)
//    as well and shares the table entry (value >= 128 = rejected).
//  * only the container-scope Filter (first-fit with subtraction, filter.go:39-78) is a per-(pod, node) state
//    machine; it runs on the container-scope nodes only, in 32-bit arithmetic on gcd-scaled quantities.
//  * nodes are split by control-flow class into two ORDER-PRESERVING compact lists (container-scope / pod-scope
//    nodes that reach a handler); every other node is decided by its flags.  A CTA of the P x N kernel owns a tile of
//    consecutive NATURAL node indices cut (on the host, at 32-node boundaries) so that it holds at most 256
//    container-scope nodes -- one per thread for the state machine, a contiguous slot range; their results are
//    staged in shared memory, and the expansion to the [P][Npad] score / reason / feasibility outputs is written in
//    natural order with fully coalesced rows (round 1's permuted kernel scattered them: 30.9 sectors per request).
//
// Exact 32-bit arithmetic: per resource slot every quantity that enters a comparison or subtraction (zone
// Available, requests; milli-units) is divided by their common gcd gm[r]; every quantity that enters a score ratio
// (Quantity.Value() of capacity and request) by gv[r].  Comparisons, differences and the ratios
// (cv - rv) * 100 / cv are invariant under the common factor, so results are bit-identical to the 64-bit path as long as
// the scaled values fit (checked on the host; otherwise nrt_eval keeps the direct 64-bit kernel of nrt.cu).
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <type_traits>

#include "engine.h"

namespace b200s {

namespace {

constexpr int C_MAX = B200S_NRT_MAX_CONT;
// tuning knobs (defaults = the measured best; -D overrides for A/B builds, tools/ab_build.sh)
#ifndef B200S_NRT2_SPAN
#define B200S_NRT2_SPAN 1024
#endif
#ifndef B200S_NRT2_PT
#define B200S_NRT2_PT 32
#endif
#ifndef B200S_NRT2_SKIP
#define B200S_NRT2_SKIP 1
#endif
constexpr int TILE = 256;                   // threads per CTA of the P x N kernel = most container-scope slots of one tile
constexpr int TILE_SPAN = B200S_NRT2_SPAN;  // most natural node indices of one tile
constexpr int PT = B200S_NRT2_PT;           // pods per CTA
constexpr int UT = 32;     // request vectors per CTA of the table kernel
constexpr int32_t S_MIN = INT32_MIN, S_MAX = INT32_MAX;
constexpr int64_t LIM_MILLI = (int64_t)1 << 30;  // scaled milli quantities stay below (sentinels are +-2^31)
constexpr int64_t LIM_VALUE = 42000000;          // scaled Value() quantities: x100 stays below 2^32 (32-bit ratio path)
constexpr int64_t LIM_VALUE_WIDE = (int64_t)1 << 31;  // beyond that, up to here: 64-bit numerator (WIDE table kernel)

// One distinct request vector of the batch, scaled (device record, 96 bytes).
struct alignas(16) VecRec {
  int32_t eff[4];   // Filter threshold per resource: S_MIN = does not constrain, S_MIN+1 = any LISTING zone (QoS-exempt
                    // NUMA-affine resource, numaresources.go:137-142), else the scaled quantity
  int32_t sub[4];   // what an app container takes from its zone (0 where exempt / zero / not requested)
  int32_t rq[4];    // scaled request (milli domain) for the `req > cap` test of the strategies
  int32_t rv[4];    // scaled Quantity.Value() of the request
  int32_t w[4];     // weight of the resource if requested (the strategies iterate the requested NAMES), else 0
  uint32_t wsum;    // sum of w
  uint32_t wmagic;  // floor(2^32 / wsum) (2^32-1 for 1)
  uint8_t need;     // bit r: requested with a non-zero quantity (filter.go:101-105)
  uint8_t guar;     // Guaranteed QoS
  uint8_t nreq;     // number of requested resource names
  uint8_t mask;     // bit r: resource name requested
  uint32_t pad;
};
static_assert(sizeof(VecRec) == 96, "VecRec layout");

struct HostVec {
  int64_t req[4];
  uint8_t mask, guar;
};

inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
inline uint64_t hash_vec(const HostVec& v) {
  uint64_t h = 0x9e3779b97f4a7c15ull ^ ((uint64_t)v.mask << 8 | v.guar);
  for (int r = 0; r < 4; ++r) h = mix64(h ^ (uint64_t)v.req[r]);
  return h;
}
inline bool same_vec(const HostVec& a, const HostVec& b) {
  return a.mask == b.mask && a.guar == b.guar && a.req[0] == b.req[0] && a.req[1] == b.req[1] && a.req[2] == b.req[2] &&
         a.req[3] == b.req[3];
}
inline uint64_t gcd_u64(uint64_t a, uint64_t b) { return std::gcd(a, b); }
inline int64_t qty_value_h(int64_t milli) { return (milli + 999) / 1000; }  // Quantity.Value() of a non-negative quantity

}  // namespace

// Host + device state of the batched path (owned by the ctx, opaque to engine.cu).
struct Nrt2 {
  // ---- snapshot side
  uint64_t gm[4] = {0, 0, 0, 0}, gv[4] = {0, 0, 0, 0};  // gcd of listed zone Available (milli / Value()); 0 = none
  int64_t maxm[4] = {0, 0, 0, 0};
  bool neg = false;          // a listed Available is negative: the scaled encoding does not apply
  uint64_t stats_serial = ~0ull;  // snapshot serial the four lines above were reduced at
  DevBuf stats_part;              // per-CTA partials of the reduction
  bool lists_valid = false;  // class lists match the flags mirror
  int Nc = 0, Np = 0, Sc = 0, Sp = 0;  // container-scope / pod-scope handler nodes; padded to 128
  DevBuf slot, list, tile_n0, tile_c0;  // [Npad] int32, [Sc + Sp] int32, [tiles + 1] first node / first container slot
  int n_tiles = 0;
  DevBuf filt, capv, magic, nzs, nrm;  // [Z][R][S] int32 x3, [S] u8 x2 (node columns of the two lists, scaled)
  uint64_t node_prep_serial = ~0ull;
  uint64_t node_prep_gm[4] = {0, 0, 0, 0}, node_prep_gv[4] = {0, 0, 0, 0};
  bool wide = false, node_prep_wide = false;  // a scaled Value() needs the 64-bit numerator form of the ratio
  // ---- pod side (host dictionary, built at pods_upload)
  bool pods_ok = false;  // the batch fits the path (shape, signs, QoS consistency)
  const char* pods_note = "no NodeResourceTopologyMatch pod columns";
  std::vector<HostVec> vecs;
  std::vector<int32_t> pod_vec;  // [P][C_MAX + 1] vector id or -1
  std::vector<int32_t> tc_of, tp_of;      // [U] table row of the vector or -1
  std::vector<int32_t> tc_list, tp_list;  // row -> vector id
  uint64_t pgm[4] = {0, 0, 0, 0}, pgv[4] = {0, 0, 0, 0};
  int64_t pmaxm[4] = {0, 0, 0, 0};
  bool balanced_ok = true;  // every scored vector names >= 2 resources (else the variance is NaN, balanced_allocation.go:41)
  uint64_t pods_serial = 0;
  // ---- prepared batch (device), keyed
  uint64_t prep_pods_serial = ~0ull, prep_cfg = ~0ull;
  uint64_t prep_gm[4] = {0, 0, 0, 0}, prep_gv[4] = {0, 0, 0, 0};
  bool prep_ok = false;
  DevBuf vecrec, d_pod_vec, d_pod_tc, d_pod_tp, d_tc_list, d_tp_list;
  PinStage stage;  // pinned staging of the per-batch uploads below
  // quotient-table form of the table kernels (Least / MostAllocated): distinct (resource, request value) keys
  bool use_q = false;
  int n_qkeys = 0;
  DevBuf qkeys, rowq_c, rowq_p, Q;  // [K] QKey, [rows] RowQ x2, [K][S] u32 (one byte per zone)
  DevBuf Tc, Tp;
  int last_path = 0;  // 1 direct, 2 table (what the last nrt_eval ran)
  const char* note = "";  // why the batched path was declined last time (static string)
  int force = 0;      // 0 auto, 1 direct, 2 table when applicable
  ~Nrt2() {
    for (DevBuf* b : {&stats_part, &slot, &list, &tile_n0, &tile_c0, &filt, &capv, &magic, &nzs, &nrm, &vecrec, &d_pod_vec, &d_pod_tc, &d_pod_tp,
                      &d_tc_list, &d_tp_list, &Tc, &Tp, &qkeys, &rowq_c, &rowq_p, &Q})
      b->release();
    stage.release();
  }
};

Nrt2* nrt2_get(b200s_ctx* c) {
  if (!c->nrt2) c->nrt2 = new Nrt2();
  return c->nrt2;
}
void nrt2_destroy(b200s_ctx* c) {
  delete c->nrt2;
  c->nrt2 = nullptr;
}
void nrt2_set_force(b200s_ctx* c, int path) { nrt2_get(c)->force = path; }
int nrt2_last_path(b200s_ctx* c) { return c->nrt2 ? c->nrt2->last_path : 0; }
const char* nrt2_note(b200s_ctx* c) { return c->nrt2 ? c->nrt2->note : ""; }
void nrt2_note_direct(b200s_ctx* c) { nrt2_get(c)->last_path = 1; }

// ---- snapshot-side statistics ---------------------------------------------------------------------------------------
// The units (gcd of the listed zone Available per resource, in milli-units and as Quantity.Value()), the largest
// value and the presence of a negative value are reduced ON THE DEVICE from the resident columns, once per snapshot
// serial and only when a batch asks for the batched path: full uploads, row patches, OverReserve deductions and the
// on-device assume all just bump the serial -- no host-side bookkeeping that could drift from the columns.
void nrt2_on_snapshot_full(b200s_ctx* c, const b200s_nrt_nodes*) {
  Nrt2* s = nrt2_get(c);
  s->lists_valid = false;
  s->stats_serial = ~0ull;
}
void nrt2_on_patch_rows(b200s_ctx* c, int, const b200s_nrt_nodes*) { nrt2_get(c)->stats_serial = ~0ull; }
void nrt2_on_deduct(b200s_ctx* c, int, const int64_t*) { nrt2_get(c)->stats_serial = ~0ull; }
void nrt2_on_class_change(b200s_ctx* c) {
  if (c->nrt2) c->nrt2->lists_valid = false;
}

// ---- pod-side dictionary (called by pods_upload with the caller's host columns) ----------------------------------
void nrt2_on_pods(b200s_ctx* c, const b200s_nrt_pods* q, int P) {
  Nrt2* s = nrt2_get(c);
  s->pods_serial++;
  s->pods_ok = false;
  s->vecs.clear();
  s->tc_of.clear();
  s->tp_of.clear();
  s->tc_list.clear();
  s->tp_list.clear();
  for (int r = 0; r < 4; ++r) s->pgm[r] = s->pgv[r] = 0, s->pmaxm[r] = 0;
  s->balanced_ok = true;
  const int R = c->nrt_R;
  s->pods_note = "no NodeResourceTopologyMatch pod columns / more than 4 zones or resource slots";
  if (!q || P <= 0 || R > 4 || c->nrt_Z > 4) return;
  if (P < 32 && s->force != 2) {  // a scheduling cycle (P = 1) never takes the batched path: skip the dictionary
    s->pods_note = "fewer than 32 pods";
    return;
  }
  s->pods_note = "";
  s->pod_vec.assign((size_t)P * (C_MAX + 1), -1);
  size_t cap = 64;
  while (cap < (size_t)P * 6) cap <<= 1;
  std::vector<int32_t> table(cap, -1);
  const uint32_t rmask = (1u << R) - 1u;
  bool ok = true;
  auto intern = [&](int p, int slot, bool guar) -> int32_t {
    HostVec v;
    v.mask = (uint8_t)(q->req_mask[(size_t)p * (C_MAX + 1) + slot] & rmask);
    v.guar = guar ? 1 : 0;
    const int64_t* rq = q->req + ((size_t)p * (C_MAX + 1) + slot) * R;
    for (int r = 0; r < 4; ++r) {
      int64_t x = (r < R && ((v.mask >> r) & 1u)) ? rq[r] : 0;
      if (x < 0) ok = false, s->pods_note = "negative request";
      // a non-Guaranteed pod's NUMA-affine requests only matter as zero / non-zero (numaresources.go:137-142)
      if (!guar && r < R && (c->nrt_res_flags[r] & B200S_NRT_RES_AFFINE)) x = x != 0 ? 1 : 0;
      v.req[r] = x;
    }
    size_t h = (size_t)hash_vec(v) & (cap - 1);
    while (table[h] >= 0) {
      if (same_vec(s->vecs[(size_t)table[h]], v)) return table[h];
      h = (h + 1) & (cap - 1);
    }
    const int32_t id = (int32_t)s->vecs.size();
    table[h] = id;
    s->vecs.push_back(v);
    s->tc_of.push_back(-1);
    s->tp_of.push_back(-1);
    return id;
  };
  for (int p = 0; p < P; ++p) {
    const uint8_t fl = q->flags[p];
    const bool guar = q->qos[p] == B200S_QOS_GUARANTEED;
    if (fl & B200S_NRT_POD_UNSUPPORTED) continue;  // decided by the flag alone
    if (fl & B200S_NRT_POD_FILTER_BYPASS) {
      if (guar) ok = false, s->pods_note = "Filter bypass on a Guaranteed pod";  // Filter bypass is BestEffort-only (filter.go:180-183); keep the direct kernel for anything else
      continue;
    }
    const int nc = (int)q->n_init[p] + (int)q->n_app[p];
    for (int cidx = 0; cidx < nc; ++cidx) {
      const int32_t id = intern(p, cidx, guar);
      s->pod_vec[(size_t)p * (C_MAX + 1) + cidx] = id;
      if (s->tc_of[(size_t)id] < 0) {  // one row of the container table per distinct container vector
        s->tc_of[(size_t)id] = (int32_t)s->tc_list.size();
        s->tc_list.push_back(id);
      }
    }
    if (guar && nc == 0) ok = false, s->pods_note = "Guaranteed pod without containers";  // mean over zero containers (score.go:162): NaN -> direct kernel
    const int32_t id = intern(p, C_MAX, guar);
    s->pod_vec[(size_t)p * (C_MAX + 1) + C_MAX] = id;
    if (s->tp_of[(size_t)id] < 0) {
      s->tp_of[(size_t)id] = (int32_t)s->tp_list.size();
      s->tp_list.push_back(id);
    }
  }
  // statistics over the distinct vectors
  for (size_t u = 0; u < s->vecs.size(); ++u) {
    const HostVec& v = s->vecs[u];
    int nreq = 0;
    for (int r = 0; r < R; ++r) {
      if (!((v.mask >> r) & 1u)) continue;
      ++nreq;
      const bool exempt = !v.guar && (c->nrt_res_flags[r] & B200S_NRT_RES_AFFINE);
      if (exempt) continue;
      s->pgm[r] = gcd_u64(s->pgm[r], (uint64_t)v.req[r]);
      s->pmaxm[r] = std::max(s->pmaxm[r], v.req[r]);
      if (v.guar) s->pgv[r] = gcd_u64(s->pgv[r], (uint64_t)qty_value_h(v.req[r]));
    }
    if (v.guar && nreq < 2) s->balanced_ok = false;
  }
  s->pods_ok = ok;
}

namespace {

// ---- device code ----------------------------------------------------------------------------------------------------

// per-resource {gcd of milli values, gcd of Value()s, max, any negative} over the listed cells: thread-local
// accumulation over a grid-stride loop, warp shuffle + shared-memory fold (gcd is associative and commutative),
// one partial per CTA; the host folds the <= 64 partials.
struct StatsPart {
  unsigned long long gm[4], gv[4];
  long long mx[4];
  unsigned int neg, pad;
};
__device__ __forceinline__ unsigned long long dgcd(unsigned long long a, unsigned long long b) {
  while (b) {
    const unsigned long long t = a % b;
    a = b;
    b = t;
  }
  return a;
}
__global__ void __launch_bounds__(256) nrt2_stats_kernel(const uint8_t* __restrict__ nz, const uint8_t* __restrict__ zmask,
                                                         const int64_t* __restrict__ avail, int Zs, int Rs, int N, int Npad,
                                                         StatsPart* __restrict__ out) {
  __shared__ StatsPart sp[8];
  StatsPart a;
  for (int r = 0; r < 4; ++r) a.gm[r] = a.gv[r] = 0, a.mx[r] = 0;
  a.neg = 0;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < N; n += gridDim.x * 256) {
    const int nzn = min((int)nz[n], Zs);
    for (int z = 0; z < nzn; ++z) {
      const uint32_t m = zmask[(size_t)z * Npad + n];
      for (int r = 0; r < Rs && r < 4; ++r) {
        if (!((m >> r) & 1u)) continue;
        const long long v = avail[((size_t)z * Rs + r) * Npad + n];
        if (v < 0) {
          a.neg = 1;
          continue;
        }
        if (a.gm[r] != 1) a.gm[r] = dgcd(a.gm[r], (unsigned long long)v);
        if (a.gv[r] != 1) a.gv[r] = dgcd(a.gv[r], (unsigned long long)((v + 999) / 1000));
        a.mx[r] = max(a.mx[r], v);
      }
    }
  }
  for (int o = 16; o; o >>= 1) {
    for (int r = 0; r < 4; ++r) {
      a.gm[r] = dgcd(a.gm[r], __shfl_xor_sync(0xffffffffu, a.gm[r], o));
      a.gv[r] = dgcd(a.gv[r], __shfl_xor_sync(0xffffffffu, a.gv[r], o));
      a.mx[r] = max(a.mx[r], __shfl_xor_sync(0xffffffffu, a.mx[r], o));
    }
    a.neg |= __shfl_xor_sync(0xffffffffu, a.neg, o);
  }
  if ((threadIdx.x & 31) == 0) sp[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      for (int r = 0; r < 4; ++r) {
        a.gm[r] = dgcd(a.gm[r], sp[w].gm[r]);
        a.gv[r] = dgcd(a.gv[r], sp[w].gv[r]);
        a.mx[r] = max(a.mx[r], sp[w].mx[r]);
      }
      a.neg |= sp[w].neg;
    }
    out[blockIdx.x] = a;
  }
}

struct NodePrepArgs {
  const uint8_t* node_flags;
  const uint8_t* nz;
  const uint8_t* node_res_mask;
  const uint8_t* zone_res_mask;  // [Zs][Npad]
  const int64_t* avail;          // [Zs][Rs][Npad]
  const int32_t* list;           // [S] node of slot, -1 = padding
  int Zs, Rs, Npad, S;
  uint64_t gm[4], gv[4];
  uint8_t res_flags[4];
  int wide;  // magic column holds the fp32 reciprocal bits instead of floor(2^32 / capv)
};

// One thread per slot of the two class lists: the node's zones x resources block in the scaled encodings.
//   filt  Filter operand (see nrt.cu nrt_filter): scaled Available where the zone lists r; S_MIN where it does not but
//         another zone does, or no zone does and r is NUMA-bound; S_MAX where no zone lists a host-level r (:139-142)
//   capv  scaled Value() of the capacity the strategies see (0 = zone lacks the resource or has none left)
//   magic floor(2^32 / capv) for the exact reciprocal division
template <int Z, int R>
__global__ void nrt2_node_prep_kernel(NodePrepArgs a, int32_t* __restrict__ filt, int32_t* __restrict__ capv,
                                      uint32_t* __restrict__ magic, uint8_t* __restrict__ nzs, uint8_t* __restrict__ nrm) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= a.S) return;
  const int n = a.list[s];
  int nz = 0;
  uint32_t zm[Z];
#pragma unroll
  for (int z = 0; z < Z; ++z) zm[z] = 0;
  if (n >= 0) {
    nz = min((int)a.nz[n], Z);
#pragma unroll
    for (int z = 0; z < Z; ++z)
      if (z < a.Zs && z < nz) zm[z] = a.zone_res_mask[(size_t)z * a.Npad + n];
  }
  nzs[s] = (uint8_t)nz;
  nrm[s] = n >= 0 ? a.node_res_mask[n] : 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    uint32_t any = 0;
#pragma unroll
    for (int z = 0; z < Z; ++z) any |= (zm[z] >> r) & 1u;
    const int32_t none = (!any && r < a.Rs && (a.res_flags[r] & B200S_NRT_RES_HOST_LEVEL)) ? S_MAX : S_MIN;
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      int32_t f = none, cv = 0;
      if (r < a.Rs && ((zm[z] >> r) & 1u)) {
        const int64_t av = a.avail[((size_t)z * a.Rs + r) * a.Npad + n];
        f = (int32_t)((uint64_t)av / a.gm[r]);
        cv = (int32_t)((uint64_t)((av + 999) / 1000) / a.gv[r]);
      }
      const size_t o = ((size_t)z * R + r) * a.S + s;
      filt[o] = f;
      capv[o] = cv;
      magic[o] = a.wide ? __float_as_uint(1.0f / (float)(cv == 0 ? 1 : cv))
                        : (cv <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / (uint32_t)cv));
    }
  }
}

// floor(num / d) for num < 2^32 with m = floor(2^32 / d) (2^32 - 1 for d == 1): the estimate umulhi(num, m) is the
// quotient or one below it (m >= 2^32/d - 1 gives num*m/2^32 > num/d - 1), so one correction step is exact.
__device__ __forceinline__ uint32_t div_magic(uint32_t num, uint32_t d, uint32_t m) {
  uint32_t q = __umulhi(num, m);
  const uint32_t rem = num - q * d;
  return rem >= d ? q + 1 : q;
}
// floor(a * 100 / cv) for 0 <= a <= cv < 2^31 with rcv ~ 1/cv in fp32: the fp32 estimate of a quotient <= 100 is
// off by far less than 1 (relative error < 2^-20), so after truncation it is the quotient or a neighbour; the exact
// 64-bit residual decides.
__device__ __forceinline__ uint32_t div100_wide(uint32_t a, uint32_t cv, float rcv) {
  const uint64_t num = (uint64_t)a * 100u;
  uint32_t q = (uint32_t)((float)a * 100.0f * rcv);
  const int64_t rem = (int64_t)num - (int64_t)((uint64_t)q * cv);
  q = rem < 0 ? q - 1 : (rem >= (int64_t)cv ? q + 1 : q);
  return q;
}
__device__ __forceinline__ int64_t f2i(double x) {
  if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}

struct TableArgs {
  const int32_t* filt;   // + class base already applied: [Z][R][S] with stride S, element (z, r, slot)
  const int32_t* capv;
  const uint32_t* magic;
  const uint8_t* nzs;
  const uint8_t* nrm;
  const VecRec* vecs;
  const int32_t* rows;  // row -> vector id
  int nrows;
  int S;      // stride of the node columns
  int count;  // slots of this class (padded to 128)
  int most;
};

// T[row][slot] for UT request vectors x 128 slots per CTA: the thread keeps its node's cells in registers, the
// vectors come from shared memory (warp-uniform).  POD: the byte table of the pod-scope nodes -- entry = score (100 for
// a non-Guaranteed vector, score.go:72-75) if resourcesAvailableInAnyNUMANodes finds a zone (filter.go:162-173), else
// 128 + B200S_REASON_NRT_ALIGN_POD.  Otherwise the 16-bit table of the container-scope nodes: the vector's score in
// the low byte and, in bits 8.., the zones it fits on the node as the snapshot has it -- the Filter verdict of every
// init container and of the first app container, which see the zones before any subtraction (filter.go:43-66).
template <int Z, int R, int SC, bool POD, bool WIDE>
__global__ void __launch_bounds__(128) nrt2_table_kernel(TableArgs a, void* __restrict__ Tout) {
  __shared__ VecRec sv[UT];
  const int row0 = blockIdx.y * UT, nrow = min(UT, a.nrows - row0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.vecs);
    uint4* dst = reinterpret_cast<uint4*>(sv);
    for (int i = threadIdx.x; i < nrow * 6; i += 128) dst[i] = src[(size_t)a.rows[row0 + i / 6] * 6 + i % 6];
  }
  const int slot = blockIdx.x * 128 + threadIdx.x;
  int32_t filt[Z][R], cv[Z][R];
  uint32_t mg[Z][R];
#pragma unroll
  for (int z = 0; z < Z; ++z)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const size_t o = ((size_t)z * R + r) * a.S + slot;
      filt[z][r] = a.filt[o];
      cv[z][r] = a.capv[o];
      mg[z][r] = a.magic[o];
    }
  const int nz = a.nzs[slot];
  const uint32_t nrm = a.nrm[slot];
  __syncthreads();
  for (int j = 0; j < nrow; ++j) {
    const VecRec& v = sv[j];
    // resourcesAvailableInAnyNUMANodes (filter.go:90-160) on the node as the snapshot has it: which zones fit
    uint32_t ok = 0;
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      bool fits = true;
#pragma unroll
      for (int r = 0; r < R; ++r) fits &= filt[z][r] >= v.eff[r];
      ok |= (fits ? 1u : 0u) << z;
    }
    if (v.need & ~nrm) ok = 0;  // :107-113: a non-zero request must be reported at node level
    int32_t score = POD ? 100 : 0;
    if (v.guar) {  // warp-uniform
      int32_t min_score = 0;
#pragma unroll
      for (int z = 0; z < Z; ++z) {
        int32_t s;
        if constexpr (SC == 1) {
          // balanced_allocation.go:27-54 -- same operation order as nrt.cu strategy_score (fractions of scaled
          // Value()s: the quotient of the same rational, correctly rounded, is the same double)
          double fr[R];
          bool over = false;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            fr[r] = 0;
            if (!((v.mask >> r) & 1u)) continue;
            const int32_t c = cv[z][r];
            const double q = (double)v.rv[r] / (double)(c == 0 ? 1 : c);
            const double f = c == 0 ? 1.0 : q;
            over |= f > 1;
            fr[r] = f;
          }
          const int n = v.nreq;
          double sum = 0;
#pragma unroll
          for (int r = 0; r < R; ++r)
            if ((v.mask >> r) & 1u) sum += fr[r];
          const double mean = sum / (double)n;
          double ss = 0, comp = 0;
#pragma unroll
          for (int r = 0; r < R; ++r)
            if ((v.mask >> r) & 1u) {
              const double d = fr[r] - mean;
              ss += d * d;
              comp += d;
            }
          const double variance = (ss - comp * comp / (double)n) / ((double)n - 1);
          s = over ? 0 : (int32_t)f2i((1 - variance) * 100.0);
        } else {
          // least_allocated.go:25-55 / most_allocated.go:25-54: sum_r w_r * s_r / sum_r w_r over the requested names
          uint32_t acc = 0;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int32_t c = cv[z][r];
            const bool zero = c == 0 || v.rq[r] > filt[z][r];  // capacity 0 (or resource missing) or request > capacity
            const uint32_t av = zero ? 0u : (uint32_t)(a.most ? v.rv[r] : c - v.rv[r]);
            const uint32_t q = WIDE ? div100_wide(av, (uint32_t)(c == 0 ? 1 : c), __uint_as_float(mg[z][r]))
                                    : div_magic(av * 100u, (uint32_t)(c == 0 ? 1 : c), mg[z][r]);
            acc += q * (uint32_t)v.w[r];
          }
          s = v.wsum == 0 ? 0 : (int32_t)div_magic(acc, v.wsum, v.wmagic);
        }
        // scoreForEachNUMANode (score.go:110-124): minimum of the non-zero zone scores
        const bool take = z < nz && (min_score == 0 || (s != 0 && s < min_score));
        min_score = take ? s : min_score;
      }
      score = min_score;
    }
    if constexpr (POD)
      static_cast<uint8_t*>(Tout)[(size_t)(row0 + j) * a.count + slot] =
          (uint8_t)(ok != 0 ? (uint32_t)score : 128u + B200S_REASON_NRT_ALIGN_POD);
    else  // container table: score of the vector (Guaranteed only) | zones that fit before any subtraction << 8
      static_cast<uint16_t*>(Tout)[(size_t)(row0 + j) * a.count + slot] = (uint16_t)((uint32_t)score | (ok << 8));
  }
}

// ---- quotient tables -----------------------------------------------------------------------------------------------------
// A table entry is a function of (request vector, node), but each of its 16 divisions is a function of ONE request VALUE
// and one (zone, resource) cell: floor((cv - rv) * 100 / cv) (least_allocated.go:57-66; rv * 100 / cv for
// most_allocated.go:56-63) and the Filter comparison available >= request (filter.go:127-135).  A batch names few distinct
// values per resource (c4: ~30 against 3 700 vectors), so the quotients are computed once per (value, cell) -- Q[key][slot],
// one byte per zone: quotient | fits << 7 -- and a table entry is four coalesced word loads, a 16-bit-lane weighted sum,
// four exact divisions by the weight sum and the minimum over the zones.
struct QKey {
  int32_t eff, rq, rv, r;  // Filter threshold, scaled request (milli), scaled Value(), resource slot
};
struct alignas(16) RowQ {   // one table row = one distinct request vector
  uint16_t j[4];            // key of each resource slot
  uint16_t w[4];            // weight if the resource is named, else 0 (sum <= 655: 100 x sum fits a 16-bit lane)
  uint32_t wsum, wmagic;
  uint8_t need, guar;
  uint16_t pad;
  uint32_t pad2;
};
static_assert(sizeof(RowQ) == 32, "RowQ layout");
constexpr int QT = 16;    // keys per CTA of the quotient kernel
constexpr int RT = 64;    // rows per CTA of the quotient-table kernel

struct QArgs {
  const int32_t* filt;  // [Z][R][S]
  const int32_t* capv;
  const uint32_t* magic;
  const QKey* keys;
  int K, S, most;
};

template <int Z, int R, bool WIDE>
__global__ void __launch_bounds__(128) nrt2_q_kernel(QArgs a, uint32_t* __restrict__ Q) {
  __shared__ QKey sk[QT];
  const int k0 = blockIdx.y * QT, nk = min(QT, a.K - k0);
  if ((int)threadIdx.x < nk) sk[threadIdx.x] = a.keys[k0 + threadIdx.x];
  __syncthreads();
  const int slot = blockIdx.x * 128 + threadIdx.x;
  if (slot >= a.S) return;
  for (int j = 0; j < nk; ++j) {
    const QKey key = sk[j];  // warp-uniform
    uint32_t word = 0;
#pragma unroll
    for (int z = 0; z < Z; ++z) {
      const size_t o = ((size_t)z * R + key.r) * a.S + slot;
      const int32_t f = a.filt[o], c = a.capv[o];
      const uint32_t mg = a.magic[o];
      const bool fits = f >= key.eff;
      const bool zero = c == 0 || key.rq > f;  // capacity 0 (or resource missing) or request > capacity
      const uint32_t av = zero ? 0u : (uint32_t)(a.most ? key.rv : c - key.rv);
      const uint32_t q = WIDE ? div100_wide(av, (uint32_t)(c == 0 ? 1 : c), __uint_as_float(mg))
                              : div_magic(av * 100u, (uint32_t)(c == 0 ? 1 : c), mg);
      word |= (min(q, 127u) | (fits ? 128u : 0u)) << (8 * z);
    }
    Q[(size_t)(k0 + j) * a.S + slot] = word;
  }
}

struct TableQArgs {
  const uint32_t* Q;   // + class base already applied; stride S
  const uint8_t* nzs;  // + class base
  const uint8_t* nrm;
  const RowQ* rows;
  int nrows, S, count;
};

// Same entries as nrt2_table_kernel<4, 4, 0, POD, *> (bit for bit: the per-cell quotients and comparisons are those of
// nrt2_q_kernel, the combination below is the rest of that kernel).
template <bool POD>
__global__ void __launch_bounds__(128) nrt2_tableq_kernel(TableQArgs a, void* __restrict__ Tout) {
  __shared__ RowQ sr[RT];
  const int row0 = blockIdx.y * RT, nrow = min(RT, a.nrows - row0);
  {
    const uint4* src = reinterpret_cast<const uint4*>(a.rows + row0);
    uint4* dst = reinterpret_cast<uint4*>(sr);
    for (int i = threadIdx.x; i < nrow * 2; i += 128) dst[i] = src[i];
  }
  const int slot = blockIdx.x * 128 + threadIdx.x;
  const int nz = a.nzs[slot];
  const uint32_t nrm = a.nrm[slot];
  const uint32_t* q = a.Q + slot;
  __syncthreads();
#pragma unroll 2
  for (int j = 0; j < nrow; ++j) {
    const RowQ v = sr[j];
    const uint32_t w0 = q[(size_t)v.j[0] * a.S], w1 = q[(size_t)v.j[1] * a.S], w2 = q[(size_t)v.j[2] * a.S],
                   w3 = q[(size_t)v.j[3] * a.S];
    // bit 7 of byte z of every word: zone z fits resource r; gather the four bits of the AND into a nibble
    uint32_t ok = ((((w0 & w1 & w2 & w3) >> 7) & 0x01010101u) * 0x01020408u) >> 24;
    if (v.need & ~nrm) ok = 0;  // filter.go:107-113: a non-zero request must be reported at node level
    uint32_t score = POD ? 100u : 0u;
    if (v.guar) {  // warp-uniform
      // zones 0 / 2 in the 16-bit lanes of A, zones 1 / 3 in those of B
      const uint32_t A = (w0 & 0x007f007fu) * v.w[0] + (w1 & 0x007f007fu) * v.w[1] + (w2 & 0x007f007fu) * v.w[2] +
                         (w3 & 0x007f007fu) * v.w[3];
      const uint32_t B = ((w0 >> 8) & 0x007f007fu) * v.w[0] + ((w1 >> 8) & 0x007f007fu) * v.w[1] +
                         ((w2 >> 8) & 0x007f007fu) * v.w[2] + ((w3 >> 8) & 0x007f007fu) * v.w[3];
      const uint32_t acc[4] = {A & 0xffffu, B & 0xffffu, A >> 16, B >> 16};
      uint32_t min_score = 0;
#pragma unroll
      for (int z = 0; z < 4; ++z) {
        const uint32_t sz = v.wsum == 0 ? 0u : div_magic(acc[z], v.wsum, v.wmagic);
        // scoreForEachNUMANode (score.go:110-124): minimum of the non-zero zone scores
        const bool take = z < nz && (min_score == 0 || (sz != 0 && sz < min_score));
        min_score = take ? sz : min_score;
      }
      score = min_score;
    }
    if constexpr (POD)
      static_cast<uint8_t*>(Tout)[(size_t)(row0 + j) * a.count + slot] =
          (uint8_t)(ok != 0 ? score : 128u + B200S_REASON_NRT_ALIGN_POD);
    else
      static_cast<uint16_t*>(Tout)[(size_t)(row0 + j) * a.count + slot] = (uint16_t)(score | (ok << 8));
  }
}

struct ExpandArgs {
  // node side
  const uint8_t* node_flags;  // [Npad]
  const int32_t* slot;        // [Npad] class-local slot
  const int32_t* tile_n0;     // [tiles + 1] first natural node index of the tile (multiple of 32)
  const int32_t* tile_c0;     // [tiles + 1] first container-scope slot of the tile
  const int32_t* filt;        // container-scope slots, stride S
  const uint8_t* nrm;
  int S;
  int Sc, Sp;
  // pod side
  const uint8_t* qos;
  const uint8_t* flags;
  const uint8_t* n_init;
  const uint8_t* n_app;
  const uint8_t* kind;     // [P][C_MAX]
  const int32_t* pod_vec;  // [P][C_MAX + 1]
  const int32_t* pod_tc;   // [P][C_MAX] row of Tc or -1
  const int32_t* pod_tp;   // [P] row of Tp or -1
  const VecRec* vecs;
  const uint16_t* Tc;  // [Uc][Sc] score | fresh-node fit mask << 8
  const uint8_t* Tp;  // [Ue][Sp]
  const uint64_t* upstream;
  int words, N, Npad, P;
};

struct alignas(16) PodMeta {
  uint8_t flags;   // bit 0 Filter bypass, bit 1 unsupported pod, bit 2 Guaranteed
  uint8_t n_init, steps, first;  // init containers, init + app containers, containers that see the unmodified zones
  uint32_t tpoff;      // row of Tp x Sp
  uint32_t inv_steps;  // ceil(65536 / steps): (sum * inv_steps) >> 16 == sum / steps for sum <= 800
  uint32_t pad;
};

__device__ __forceinline__ int32_t mad_lo(int32_t a, int32_t b, int32_t c) {  // a * b + c on the FMA pipe (IMAD)
  int32_t d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// The P x N pass.  Phase 1: the container-scope nodes of this tile (a contiguous slot range, one per thread) run
// singleNUMAContainerLevelHandler (filter.go:39-78) per pod -- table look-ups for the containers that see the
// unmodified zones, the first-fit state machine with subtraction for the app containers after the first -- and
// gather the container-scope score (score.go:152-165) from the same table entries; one byte per (pod, slot) goes
// to shared memory (< 128: feasible with that score, >= 128: 128 + reason).  Phase 2: the threads sweep the tile's
// NATURAL node indices and expand pod by pod -- flag-decided nodes, pod-scope nodes (one byte of Tp),
// container-scope nodes (the staged byte) -- into coalesced rows.  All pod data is warp-uniform (shared-memory
// broadcasts, uniform branches).
template <int Z, int R, class OutT>
__global__ void __launch_bounds__(TILE, 4) nrt2_expand_kernel(ExpandArgs a, OutT* __restrict__ out,
                                                           uint32_t* __restrict__ feas32, uint8_t* __restrict__ reasons) {
  static_assert(R == 4, "the request vectors travel as int4");
  __shared__ int4 s_eff[PT][C_MAX];
  __shared__ int4 s_sub[PT][C_MAX];
  __shared__ uint32_t s_tcoff[PT][C_MAX];  // row of Tc x Sc (0 for unused container slots)
  __shared__ uint8_t s_code[PT][C_MAX];
  __shared__ PodMeta s_meta[PT];
  __shared__ uint8_t s_v[PT][TILE];
  const int tid = threadIdx.x;
  const int p0 = blockIdx.y * PT, pend = min(PT, a.P - p0);
  for (int i = tid; i < pend * C_MAX; i += TILE) {
    const int pp = i / C_MAX, c = i % C_MAX;
    const int u = a.pod_vec[(size_t)(p0 + pp) * (C_MAX + 1) + c];
    int4 e = make_int4(S_MIN, S_MIN, S_MIN, S_MIN), sb = make_int4(0, 0, 0, 0);
    if (u >= 0) {
      const int4* v = reinterpret_cast<const int4*>(a.vecs + u);  // VecRec: eff[4], sub[4] lead the record
      e = v[0];
      sb = v[1];
    }
    s_eff[pp][c] = e;
    s_sub[pp][c] = sb;
    // reason code if this container cannot be aligned: logging.go:68-73 names init containers with RestartPolicy
    // Always "sidecar"
    const uint8_t kind = a.kind[(size_t)(p0 + pp) * C_MAX + c];
    s_code[pp][c] = c >= a.n_init[p0 + pp] ? B200S_REASON_NRT_ALIGN_CONTAINER
                                           : (kind == B200S_CONT_SIDECAR ? B200S_REASON_NRT_ALIGN_SIDECAR : B200S_REASON_NRT_ALIGN_INIT);
    s_tcoff[pp][c] = (uint32_t)max(a.pod_tc[(size_t)(p0 + pp) * C_MAX + c], 0) * (uint32_t)a.Sc;
  }
  for (int i = tid; i < pend; i += TILE) {
    const int p = p0 + i;
    PodMeta m;
    const int n_init = a.n_init[p], steps = n_init + a.n_app[p];
    m.flags = (uint8_t)((a.flags[p] & 3u) | (a.qos[p] == B200S_QOS_GUARANTEED ? 4u : 0u));
    m.n_init = (uint8_t)n_init;
    m.steps = (uint8_t)steps;
    // containers that see the zones as the snapshot has them: the init containers, the first app container, and every
    // further app container as long as its predecessors took nothing (a Burstable / BestEffort pod's NUMA-affine
    // requests are not subtracted, numaresources.go:137-142 -- for most of them the state machine never starts)
    int first = min(n_init + 1, steps);
    while (first < steps) {
      const int32_t u = a.pod_vec[(size_t)p * (C_MAX + 1) + first - 1];
      if (u >= 0) {
        const int4 sb = reinterpret_cast<const int4*>(a.vecs + u)[1];
        if ((sb.x | sb.y | sb.z | sb.w) != 0) break;
      }
      ++first;
    }
    m.first = (uint8_t)first;
    m.tpoff = (uint32_t)max(a.pod_tp[p], 0) * (uint32_t)a.Sp;
    m.inv_steps = (65536u + (uint32_t)max(steps, 1) - 1u) / (uint32_t)max(steps, 1);
    m.pad = 0;
    s_meta[i] = m;
  }
  const int n0 = a.tile_n0[blockIdx.x], n1 = a.tile_n0[blockIdx.x + 1];
  const int c0 = a.tile_c0[blockIdx.x], ncs = a.tile_c0[blockIdx.x + 1] - c0;
  __syncthreads();
  if (tid < ncs) {
    const int slot = c0 + tid;
    const uint32_t nrm = a.nrm[slot];
    int32_t node[Z][R];
    bool nolist[R];  // no zone lists the (host-level) resource: nothing to subtract from (filter.go:139-142)
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int z = 0; z < Z; ++z) node[z][r] = a.filt[((size_t)z * R + r) * a.S + slot];
      nolist[r] = node[0][r] == S_MAX;
      // a resource the node does not report at node level rejects every non-zero request for it (:107-113):
      // S_MIN fails `>= eff` for every needed resource and passes the unconstrained ones (eff == S_MIN)
      if (!((nrm >> r) & 1u)) {
#pragma unroll
        for (int z = 0; z < Z; ++z) node[z][r] = S_MIN;
      }
    }
    const uint16_t* tc = a.Tc + slot;
    for (int pp = 0; pp < pend; ++pp) {
      const PodMeta m = s_meta[pp];
      if (m.flags & 3u) continue;  // Filter bypass / unsupported pod: decided in phase 2
      const int steps = m.steps, first = m.first;
      // every container's table entry at once (independent loads): score in the low byte, zones it fits on the
      // node as the snapshot has it in bits 8..
      // init containers and the FIRST app container see the unmodified zones (init containers do not subtract,
      // :43-55): their verdict is the stored fit mask
      uint32_t sum = 0, ok = 0, fail = 0;
      auto gather = [&](auto cm) {  // most pods have <= 4 containers: half the predicated work (steps is CTA-uniform)
        constexpr int CM = decltype(cm)::value;
        uint32_t e[CM];
#pragma unroll
        for (int s = 0; s < CM; ++s) e[s] = s < steps ? (uint32_t)tc[s_tcoff[pp][s]] : 0u;
#pragma unroll
        for (int s = 0; s < CM; ++s) {
          sum += e[s] & 0xFFu;
          const bool look = s < first;
          ok = look ? e[s] >> 8 : ok;
          fail |= (look && e[s] < 256u) ? 1u << s : 0u;
        }
      };
      if (steps <= 4)
        gather(std::integral_constant<int, 4>());
      else
        gather(std::integral_constant<int, C_MAX>());
      uint32_t reason = fail ? s_code[pp][__ffs(fail) - 1] : 0u;
      if (steps > first) {  // further app containers: the first-fit state machine with subtraction (:57-76)
        int32_t zs[Z][R];
#pragma unroll
        for (int z = 0; z < Z; ++z)
#pragma unroll
          for (int r = 0; r < R; ++r) zs[z][r] = node[z][r];
        for (int s = first; s < steps; ++s) {
          // the previous app container takes its request from the lowest fitting zone
          // (subtractResourcesFromNUMANodeList, numaresources.go:145-182); the last one's subtraction is never read
          const int4 sb = s_sub[pp][s - 1], ef = s_eff[pp][s];
          const int32_t q[4] = {nolist[0] ? 0 : sb.x, nolist[1] ? 0 : sb.y, nolist[2] ? 0 : sb.z, nolist[3] ? 0 : sb.w};
          const int32_t ee[4] = {ef.x, ef.y, ef.z, ef.w};
          const uint32_t low = ok & (0u - ok);
          ok = 0;
#pragma unroll
          for (int z = 0; z < Z; ++z) {
            const int32_t take = -(int32_t)((low >> z) & 1u);  // -1 for the chosen zone
            bool fit = true;
#pragma unroll
            for (int r = 0; r < R; ++r) {
              zs[z][r] = mad_lo(q[r], take, zs[z][r]);
              fit &= zs[z][r] >= ee[r];
            }
            ok |= (fit ? 1u : 0u) << z;
          }
          reason = (reason == 0 && ok == 0) ? (uint32_t)B200S_REASON_NRT_ALIGN_CONTAINER : reason;
        }
      }
      // containerScopeScore (score.go:152-165): mean over init + app containers, truncated; non-Guaranteed pods
      // score 100 (:72-75)
      const uint32_t score = (m.flags & 4u) ? (sum * m.inv_steps) >> 16 : 100u;
      s_v[pp][tid] = (uint8_t)(reason ? 128u + reason : score);
    }
  }
  __syncthreads();
  // n0 and n1 are multiples of 32: a warp is entirely inside or outside the tile (the ballot below needs full warps)
  const uint32_t* up32 = reinterpret_cast<const uint32_t*>(a.upstream);
  for (int n = n0 + tid; n < n1; n += TILE) {
    // what decides this node (the reference's gates in order, filter.go:194-209 / score.go:79-94):
    //   1 stale NRT -> "invalid node topology data"; 2 no NRT object / policy not single-numa-node -> pass, score 0
    //   (100 for a non-Guaranteed pod); 3 shape outside the dense encoding; 4 pod-scope table; 5 container-scope byte
    const uint32_t nfl = n < a.N ? a.node_flags[n] : 0u;
    const int slot = a.slot[n];
    int kind;
    if (n >= a.N) kind = 0;
    else if (!(nfl & B200S_NRT_NODE_FRESH)) kind = 1;
    else if (!(nfl & B200S_NRT_NODE_HAS_NRT) || !(nfl & B200S_NRT_NODE_SINGLE_NUMA)) kind = 2;
    else if (slot < 0) kind = 3;
    else kind = (nfl & B200S_NRT_NODE_SCOPE_POD) ? 4 : 5;
    const uint32_t v_guar = kind == 1 ? 128u + B200S_REASON_NRT_INVALID_TOPOLOGY : (kind == 3 ? 128u + B200S_REASON_UNSUPPORTED : 0u);
    const uint32_t v_other = kind == 2 ? 100u : v_guar;
    const uint8_t* tp = a.Tp + (kind == 4 ? slot : 0);
    const uint8_t* sv = &s_v[0][kind == 5 ? slot - c0 : 0];
    size_t o = (size_t)p0 * a.Npad + n;
    const size_t w32 = (size_t)(a.words * 2);
    const uint32_t* upw = up32 ? up32 + (size_t)p0 * w32 + (n >> 5) : nullptr;
#pragma unroll 4
    for (int pp = 0; pp < pend; ++pp, o += a.Npad) {
      const PodMeta m = s_meta[pp];
      uint32_t v = (m.flags & 4u) ? v_guar : v_other;
      if (kind == 4) v = tp[m.tpoff];
      if (kind == 5) v = sv[pp * TILE];
      if ((m.flags & 2u) && kind >= 3) v = 128u + B200S_REASON_UNSUPPORTED;
      if (m.flags & 1u) v = 100u;  // filter.go:180-183
      uint32_t reason = v >= 128u ? v - 128u : 0u;
      bool up = true;
      if (upw) up = (upw[(size_t)pp * w32] >> (n & 31)) & 1u;
      const bool feasible = reason == 0 && up && kind != 0;
      reason = (reason == 0 && !up) ? (uint32_t)B200S_REASON_UPSTREAM : reason;
      out[o] = (OutT)(feasible ? v : 0u);
      reasons[o] = (uint8_t)(kind != 0 ? reason : 0u);
      const uint32_t w = __ballot_sync(0xffffffffu, feasible);
      if ((tid & 31) == 0) feas32[o >> 5] = w;
    }
  }
}

int build_lists(b200s_ctx* c, Nrt2* s) {
  const int N = c->N, Npad = c->Npad;
  std::vector<int32_t> slot((size_t)Npad, -1), clist, plist;
  std::vector<uint8_t> is_c((size_t)Npad, 0);
  clist.reserve((size_t)N);
  plist.reserve((size_t)N);
  constexpr uint32_t need = B200S_NRT_NODE_FRESH | B200S_NRT_NODE_HAS_NRT | B200S_NRT_NODE_SINGLE_NUMA;
  for (int n = 0; n < N; ++n) {
    const uint32_t fl = (uint32_t)((c->nrt_key_h[(size_t)n] >> 48) & 0xFF);
    if ((fl & need) != need || (fl & B200S_NRT_NODE_UNSUPPORTED)) continue;
    if (fl & B200S_NRT_NODE_SCOPE_POD) {
      slot[(size_t)n] = (int32_t)plist.size();
      plist.push_back(n);
    } else {
      slot[(size_t)n] = (int32_t)clist.size();
      clist.push_back(n);
      is_c[(size_t)n] = 1;
    }
  }
  // tiles of the P x N kernel: consecutive 32-node groups, at most TILE container-scope nodes (one state machine per
  // thread) and at most TILE_SPAN nodes each
  std::vector<int32_t> tile_n0, tile_c0;
  int32_t cs = 0;
  for (int n = 0; n < Npad;) {
    tile_n0.push_back(n);
    tile_c0.push_back(cs);
    int in_tile = 0;
    const int start = n;
    while (n < Npad && n - start < TILE_SPAN) {
      int cnt = 0;
      for (int j = n; j < n + 32; ++j) cnt += is_c[(size_t)j];
      if (in_tile + cnt > TILE) break;
      in_tile += cnt;
      n += 32;
    }
    cs += in_tile;
  }
  tile_n0.push_back(Npad);
  tile_c0.push_back(cs);
  s->n_tiles = (int)tile_n0.size() - 1;
  s->Nc = (int)clist.size();
  s->Np = (int)plist.size();
  s->Sc = round_up(std::max(s->Nc, 1), 128);
  s->Sp = round_up(std::max(s->Np, 1), 128);
  std::vector<int32_t> list((size_t)(s->Sc + s->Sp), -1);
  std::copy(clist.begin(), clist.end(), list.begin());
  std::copy(plist.begin(), plist.end(), list.begin() + s->Sc);
  B200S_CUDA_TRY(c, s->slot.ensure((size_t)Npad * 4));
  B200S_CUDA_TRY(c, s->list.ensure(list.size() * 4));
  B200S_CUDA_TRY(c, s->tile_n0.ensure(tile_n0.size() * 4));
  B200S_CUDA_TRY(c, s->tile_c0.ensure(tile_c0.size() * 4));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(s->slot.p, slot.data(), (size_t)Npad * 4, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(s->list.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(s->tile_n0.p, tile_n0.data(), tile_n0.size() * 4, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(s->tile_c0.p, tile_c0.data(), tile_c0.size() * 4, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // pageable sources die at return
  s->lists_valid = true;
  s->node_prep_serial = ~0ull;
  return B200S_OK;
}

inline uint32_t magic_of(uint32_t d) { return d <= 1 ? 0xFFFFFFFFu : (uint32_t)(0x100000000ull / d); }

}  // namespace

static int nrt2_refresh_stats(b200s_ctx* c, Nrt2* s) {
  constexpr int kBlocks = 64;
  B200S_CUDA_TRY(c, s->stats_part.ensure(sizeof(StatsPart) * kBlocks));
  nrt2_stats_kernel<<<kBlocks, 256, 0, c->stream>>>(c->nrt_nz.as<uint8_t>(), c->nrt_zone_res_mask.as<uint8_t>(),
                                                   c->nrt_avail.as<int64_t>(), c->nrt_Z, c->nrt_R, c->N, c->Npad,
                                                   s->stats_part.as<StatsPart>());
  c->launches++;
  StatsPart h[kBlocks];
  B200S_CUDA_TRY(c, cudaMemcpyAsync(h, s->stats_part.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (int r = 0; r < 4; ++r) s->gm[r] = s->gv[r] = 0, s->maxm[r] = 0;
  s->neg = false;
  for (int b = 0; b < kBlocks; ++b) {
    for (int r = 0; r < 4; ++r) {
      s->gm[r] = gcd_u64(s->gm[r], h[b].gm[r]);
      s->gv[r] = gcd_u64(s->gv[r], h[b].gv[r]);
      s->maxm[r] = std::max<int64_t>(s->maxm[r], h[b].mx[r]);
    }
    s->neg = s->neg || h[b].neg;
  }
  s->stats_serial = c->snap_serial;
  return B200S_OK;
}

// Decides whether the batched path applies to (snapshot, pod batch, args) and prepares its device state.
// Returns 1 = applicable and prepared, 0 = keep the direct kernel, < 0 = error.
int nrt2_prepare(b200s_ctx* c) {
  Nrt2* s = nrt2_get(c);
  s->note = "";
  auto decline = [&](const char* why) { s->note = why; return 0; };
  if (s->force == 1) return decline("direct path forced");
  if (c->nrt_strategy == B200S_NRT_LEAST_NUMA_NODES) return decline("LeastNUMANodes");
  if (c->nrt_Z > 4 || c->nrt_R > 4 || c->N <= 0) return decline("more than 4 zones or 4 resource slots");
  if (s->force != 2 && c->P < 32) return decline("fewer than 32 pods");  // a P = 1 cycle is one pass of the direct kernel
  if (!s->pods_ok) return decline(s->pods_note);
  if (s->stats_serial != c->snap_serial) B200S_TRY(nrt2_refresh_stats(c, s));
  if (s->neg) return decline("negative zone availability");
  if (c->nrt_strategy == B200S_NRT_BALANCED_ALLOCATION && !s->balanced_ok)
    return decline("BalancedAllocation with a Guaranteed request naming < 2 resources");
  const int R = c->nrt_R;
  int64_t wtot = 0;
  for (int r = 0; r < R; ++r) {
    if (c->nrt_w[r] < 1 || c->nrt_w[r] > (1 << 20)) return decline("resource weight above 2^20");
    wtot += c->nrt_w[r];
  }
  if (wtot * 100 >= ((int64_t)1 << 31)) return decline("resource weights too large");
  uint64_t gm[4] = {1, 1, 1, 1}, gv[4] = {1, 1, 1, 1};
  bool wide = false;
  for (int r = 0; r < R; ++r) {
    gm[r] = gcd_u64(s->gm[r], s->pgm[r]);
    gv[r] = gcd_u64(s->gv[r], s->pgv[r]);
    if (gm[r] == 0) gm[r] = 1;
    if (gv[r] == 0) gv[r] = 1;
    const int64_t mx = std::max(s->maxm[r], s->pmaxm[r]);
    if (mx / (int64_t)gm[r] >= LIM_MILLI) return decline("a quantity does not fit 30 bits after gcd scaling");
    if (qty_value_h(mx) / (int64_t)gv[r] >= LIM_VALUE_WIDE) return decline("a Value() does not fit 31 bits after gcd scaling");
    if (qty_value_h(mx) / (int64_t)gv[r] >= LIM_VALUE) wide = true;
  }
  s->wide = wide;
  if (wide) s->note = "batched, 64-bit Value() ratios";
  if (!s->lists_valid) B200S_TRY(build_lists(c, s));
  const size_t S = (size_t)s->Sc + s->Sp;
  const size_t tc_bytes = (size_t)std::max<size_t>(s->tc_list.size(), 1) * s->Sc * 2;
  const size_t tp_bytes = (size_t)std::max<size_t>(s->tp_list.size(), 1) * s->Sp;
  if (tc_bytes + tp_bytes > ((size_t)16 << 30)) return decline("score tables above 16 GiB");
  if (tc_bytes / 2 >= ((size_t)1 << 32) || tp_bytes >= ((size_t)1 << 32)) return decline("score table above 2^32 entries");
  // node columns of the two lists in the scaled encodings (per snapshot and scale)
  if (s->node_prep_serial != c->snap_serial || s->node_prep_wide != wide || memcmp(s->node_prep_gm, gm, sizeof(gm)) != 0 ||
      memcmp(s->node_prep_gv, gv, sizeof(gv)) != 0) {
    B200S_CUDA_TRY(c, s->filt.ensure(16 * S * 4));
    B200S_CUDA_TRY(c, s->capv.ensure(16 * S * 4));
    B200S_CUDA_TRY(c, s->magic.ensure(16 * S * 4));
    B200S_CUDA_TRY(c, s->nzs.ensure(S));
    B200S_CUDA_TRY(c, s->nrm.ensure(S));
    NodePrepArgs a;
    a.node_flags = c->nrt_node_flags.as<uint8_t>();
    a.nz = c->nrt_nz.as<uint8_t>();
    a.node_res_mask = c->nrt_node_res_mask.as<uint8_t>();
    a.zone_res_mask = c->nrt_zone_res_mask.as<uint8_t>();
    a.avail = c->nrt_avail.as<int64_t>();
    a.list = s->list.as<int32_t>();
    a.Zs = c->nrt_Z;
    a.Rs = R;
    a.Npad = c->Npad;
    a.S = (int)S;
    a.wide = wide ? 1 : 0;
    for (int r = 0; r < 4; ++r) {
      a.gm[r] = gm[r];
      a.gv[r] = gv[r];
      a.res_flags[r] = r < R ? c->nrt_res_flags[r] : 0;
    }
    nrt2_node_prep_kernel<4, 4><<<(unsigned)((S + 127) / 128), 128, 0, c->stream>>>(
        a, s->filt.as<int32_t>(), s->capv.as<int32_t>(), s->magic.as<uint32_t>(), s->nzs.as<uint8_t>(),
        s->nrm.as<uint8_t>());
    c->launches += 1;
    B200S_CUDA_TRY(c, cudaGetLastError());
    s->node_prep_serial = c->snap_serial;
    s->node_prep_wide = wide;
    memcpy(s->node_prep_gm, gm, sizeof(gm));
    memcpy(s->node_prep_gv, gv, sizeof(gv));
  }
  // the batch's vectors in the scaled encoding (per pod batch, scale and plugin args)
  const uint64_t cfg = c->nrt_cfg_gen;
  if (!s->prep_ok || s->prep_pods_serial != s->pods_serial || s->prep_cfg != cfg || memcmp(s->prep_gm, gm, sizeof(gm)) != 0 ||
      memcmp(s->prep_gv, gv, sizeof(gv)) != 0) {
    s->prep_ok = false;
    const size_t U = s->vecs.size(), P = (size_t)c->P;
    std::vector<VecRec> recs(std::max<size_t>(U, 1));
    memset(recs.data(), 0, recs.size() * sizeof(VecRec));
    for (size_t u = 0; u < U; ++u) {
      const HostVec& v = s->vecs[u];
      VecRec& o = recs[u];
      uint32_t wsum = 0;
      for (int r = 0; r < 4; ++r) {
        const bool named = r < R && ((v.mask >> r) & 1u);
        const bool needed = named && v.req[r] != 0;
        const bool exempt = !v.guar && r < R && (c->nrt_res_flags[r] & B200S_NRT_RES_AFFINE);
        const int32_t q = (needed && !exempt) ? (int32_t)((uint64_t)v.req[r] / gm[r]) : 0;
        o.eff[r] = needed ? (exempt ? S_MIN + 1 : q) : S_MIN;
        o.sub[r] = q;
        o.rq[r] = (named && v.guar) ? (int32_t)((uint64_t)v.req[r] / gm[r]) : 0;
        o.rv[r] = (named && v.guar) ? (int32_t)((uint64_t)qty_value_h(v.req[r]) / gv[r]) : 0;
        o.w[r] = named ? (int32_t)c->nrt_w[r] : 0;
        wsum += (uint32_t)o.w[r];
        if (needed) o.need |= (uint8_t)(1u << r);
        if (named) o.nreq++;
      }
      o.wsum = wsum;
      o.wmagic = magic_of(wsum);
      o.guar = v.guar;
      o.mask = v.mask;
    }
    // quotient-table form: distinct (resource, eff, rq, rv) keys and the rows of both tables in terms of them
    std::vector<QKey> qkeys;
    std::vector<RowQ> rowq_c(std::max<size_t>(s->tc_list.size(), 1)), rowq_p(std::max<size_t>(s->tp_list.size(), 1));
    memset(rowq_c.data(), 0, rowq_c.size() * sizeof(RowQ));
    memset(rowq_p.data(), 0, rowq_p.size() * sizeof(RowQ));
    {
      static const bool q_enabled = []() { const char* e = getenv("B200S_NRT2_Q"); return !(e && e[0] == '0'); }();
      bool ok = q_enabled && c->nrt_strategy != B200S_NRT_BALANCED_ALLOCATION;
      std::map<std::array<int32_t, 4>, int> index;  // a few hundred keys at most (8192 cap)
      std::array<int32_t, 4> last_key[4] = {{-1, 0, 0, 0}, {-1, 0, 0, 0}, {-1, 0, 0, 0}, {-1, 0, 0, 0}};
      int last_id[4] = {0, 0, 0, 0};
      std::vector<std::array<uint16_t, 4>> vec_keys(std::max<size_t>(U, 1));
      for (size_t u = 0; u < U && ok; ++u) {
        const VecRec& o = recs[u];
        if (o.wsum > 655) ok = false;
        for (int r = 0; r < 4 && ok; ++r) {
          const std::array<int32_t, 4> key = {r, o.eff[r], o.rq[r], o.rv[r]};
          if (key == last_key[r]) {  // consecutive vectors often repeat a value
            vec_keys[u][r] = (uint16_t)last_id[r];
            continue;
          }
          auto it = index.find(key);
          if (it == index.end()) {
            if (qkeys.size() >= 8192) {
              ok = false;
              break;
            }
            it = index.emplace(key, (int)qkeys.size()).first;
            qkeys.push_back(QKey{o.eff[r], o.rq[r], o.rv[r], r});
          }
          vec_keys[u][r] = (uint16_t)it->second;
          last_key[r] = key;
          last_id[r] = it->second;
        }
      }
      if (ok && qkeys.size() * (size_t)(s->Sc + s->Sp) * 4 > ((size_t)1 << 30)) ok = false;
      auto fill = [&](std::vector<RowQ>& rows, const std::vector<int32_t>& list) {
        for (size_t i = 0; i < list.size(); ++i) {
          const VecRec& o = recs[(size_t)list[i]];
          RowQ& q = rows[i];
          for (int r = 0; r < 4; ++r) {
            q.j[r] = vec_keys[(size_t)list[i]][r];
            q.w[r] = (uint16_t)o.w[r];
          }
          q.wsum = o.wsum;
          q.wmagic = o.wmagic;
          q.need = o.need;
          q.guar = o.guar;
        }
      };
      if (ok) {
        fill(rowq_c, s->tc_list);
        fill(rowq_p, s->tp_list);
      }
      s->use_q = ok && !qkeys.empty();
      s->n_qkeys = (int)qkeys.size();
    }
    std::vector<int32_t> pod_tc(std::max<size_t>(P * C_MAX, 1), -1), pod_tp(std::max<size_t>(P, 1), -1);
    for (size_t p = 0; p < P; ++p) {
      for (int cidx = 0; cidx < C_MAX; ++cidx) {
        const int32_t u = s->pod_vec[p * (C_MAX + 1) + cidx];
        if (u >= 0) pod_tc[p * C_MAX + cidx] = s->tc_of[(size_t)u];
      }
      const int32_t u = s->pod_vec[p * (C_MAX + 1) + C_MAX];
      if (u >= 0) pod_tp[p] = s->tp_of[(size_t)u];
    }
    // through a double-buffered pinned block: no stream synchronisation, so the host prepares pod chunk i + 1 while
    // the device works on chunk i
    const size_t stage_need = 256 * 12 + recs.size() * sizeof(VecRec) + (s->pod_vec.size() + pod_tc.size() + pod_tp.size() +
                              s->tc_list.size() + s->tp_list.size()) * 4 + qkeys.size() * sizeof(QKey) +
                              (rowq_c.size() + rowq_p.size()) * sizeof(RowQ);
    void* stage_base = nullptr;
    B200S_CUDA_TRY(c, s->stage.acquire(stage_need, &stage_base));
    size_t stage_off = 0;
    auto up = [&](DevBuf& d, const void* src, size_t bytes) -> int {
      B200S_CUDA_TRY(c, d.ensure(std::max<size_t>(bytes, 16)));
      if (bytes) {
        char* st = static_cast<char*>(stage_base) + stage_off;
        memcpy(st, src, bytes);
        stage_off += (bytes + 255) / 256 * 256;
        B200S_CUDA_TRY(c, cudaMemcpyAsync(d.p, st, bytes, cudaMemcpyHostToDevice, c->stream));
      }
      return B200S_OK;
    };
    B200S_TRY(up(s->vecrec, recs.data(), recs.size() * sizeof(VecRec)));
    B200S_TRY(up(s->d_pod_vec, s->pod_vec.data(), s->pod_vec.size() * 4));
    B200S_TRY(up(s->d_pod_tc, pod_tc.data(), pod_tc.size() * 4));
    B200S_TRY(up(s->d_pod_tp, pod_tp.data(), pod_tp.size() * 4));
    B200S_TRY(up(s->d_tc_list, s->tc_list.data(), s->tc_list.size() * 4));
    B200S_TRY(up(s->d_tp_list, s->tp_list.data(), s->tp_list.size() * 4));
    if (s->use_q) {
      B200S_TRY(up(s->qkeys, qkeys.data(), qkeys.size() * sizeof(QKey)));
      B200S_TRY(up(s->rowq_c, rowq_c.data(), rowq_c.size() * sizeof(RowQ)));
      B200S_TRY(up(s->rowq_p, rowq_p.data(), rowq_p.size() * sizeof(RowQ)));
    }
    B200S_CUDA_TRY(c, s->stage.commit(c->stream));
    s->prep_pods_serial = s->pods_serial;
    s->prep_cfg = cfg;
    memcpy(s->prep_gm, gm, sizeof(gm));
    memcpy(s->prep_gv, gv, sizeof(gv));
    s->prep_ok = true;
  }
  B200S_CUDA_TRY(c, s->Tc.ensure(tc_bytes));
  B200S_CUDA_TRY(c, s->Tp.ensure(tp_bytes));
  if (s->use_q) B200S_CUDA_TRY(c, s->Q.ensure((size_t)s->n_qkeys * S * 4));
  if (s->use_q)
    s->note = wide ? "batched, quotient tables, 64-bit Value() ratios" : "batched, quotient tables";
  return 1;
}

namespace {
template <int SC, bool WIDE>
void launch_tables(b200s_ctx* c, Nrt2* s) {
  const int S = s->Sc + s->Sp;
  TableArgs a;
  a.vecs = s->vecrec.as<VecRec>();
  a.S = S;
  a.most = c->nrt_strategy == B200S_NRT_MOST_ALLOCATED;
  if (!s->tc_list.empty() && s->Nc > 0) {
    a.filt = s->filt.as<int32_t>();
    a.capv = s->capv.as<int32_t>();
    a.magic = s->magic.as<uint32_t>();
    a.nzs = s->nzs.as<uint8_t>();
    a.nrm = s->nrm.as<uint8_t>();
    a.rows = s->d_tc_list.as<int32_t>();
    a.nrows = (int)s->tc_list.size();
    a.count = s->Sc;
    dim3 grid((unsigned)(s->Sc / 128), (unsigned)((a.nrows + UT - 1) / UT));
    nrt2_table_kernel<4, 4, SC, false, WIDE><<<grid, 128, 0, c->stream>>>(a, s->Tc.as<uint16_t>());
    c->launches += 1;
  }
  if (!s->tp_list.empty() && s->Np > 0) {
    a.filt = s->filt.as<int32_t>() + s->Sc;
    a.capv = s->capv.as<int32_t>() + s->Sc;
    a.magic = s->magic.as<uint32_t>() + s->Sc;
    a.nzs = s->nzs.as<uint8_t>() + s->Sc;
    a.nrm = s->nrm.as<uint8_t>() + s->Sc;
    a.rows = s->d_tp_list.as<int32_t>();
    a.nrows = (int)s->tp_list.size();
    a.count = s->Sp;
    dim3 grid((unsigned)(s->Sp / 128), (unsigned)((a.nrows + UT - 1) / UT));
    nrt2_table_kernel<4, 4, SC, true, WIDE><<<grid, 128, 0, c->stream>>>(a, s->Tp.as<uint8_t>());
    c->launches += 1;
  }
}

// Least / MostAllocated through the quotient tables: one pass over (key, slot), then both tables from it.
void launch_tables_q(b200s_ctx* c, Nrt2* s) {
  const int S = s->Sc + s->Sp;
  QArgs qa;
  qa.filt = s->filt.as<int32_t>();
  qa.capv = s->capv.as<int32_t>();
  qa.magic = s->magic.as<uint32_t>();
  qa.keys = s->qkeys.as<QKey>();
  qa.K = s->n_qkeys;
  qa.S = S;
  qa.most = c->nrt_strategy == B200S_NRT_MOST_ALLOCATED;
  dim3 qgrid((unsigned)(S / 128), (unsigned)((qa.K + QT - 1) / QT));
  if (s->wide)
    nrt2_q_kernel<4, 4, true><<<qgrid, 128, 0, c->stream>>>(qa, s->Q.as<uint32_t>());
  else
    nrt2_q_kernel<4, 4, false><<<qgrid, 128, 0, c->stream>>>(qa, s->Q.as<uint32_t>());
  c->launches += 1;
  TableQArgs a;
  a.S = S;
  if (!s->tc_list.empty() && s->Nc > 0) {
    a.Q = s->Q.as<uint32_t>();
    a.nzs = s->nzs.as<uint8_t>();
    a.nrm = s->nrm.as<uint8_t>();
    a.rows = s->rowq_c.as<RowQ>();
    a.nrows = (int)s->tc_list.size();
    a.count = s->Sc;
    dim3 grid((unsigned)(s->Sc / 128), (unsigned)((a.nrows + RT - 1) / RT));
    nrt2_tableq_kernel<false><<<grid, 128, 0, c->stream>>>(a, s->Tc.as<uint16_t>());
    c->launches += 1;
  }
  if (!s->tp_list.empty() && s->Np > 0) {
    a.Q = s->Q.as<uint32_t>() + s->Sc;
    a.nzs = s->nzs.as<uint8_t>() + s->Sc;
    a.nrm = s->nrm.as<uint8_t>() + s->Sc;
    a.rows = s->rowq_p.as<RowQ>();
    a.nrows = (int)s->tp_list.size();
    a.count = s->Sp;
    dim3 grid((unsigned)(s->Sp / 128), (unsigned)((a.nrows + RT - 1) / RT));
    nrt2_tableq_kernel<true><<<grid, 128, 0, c->stream>>>(a, s->Tp.as<uint8_t>());
    c->launches += 1;
  }
}
}  // namespace

// Runs the batched path (nrt2_prepare returned 1; outputs ensured by the caller).
int nrt2_eval(b200s_ctx* c, int dtype) {
  Nrt2* s = nrt2_get(c);
  PluginOut& o = c->out[B200S_PLUGIN_NRT];
  if (s->use_q && c->nrt_strategy != B200S_NRT_BALANCED_ALLOCATION)
    launch_tables_q(c, s);
  else if (c->nrt_strategy == B200S_NRT_BALANCED_ALLOCATION)
    launch_tables<1, false>(c, s);  // fractions in float64: no 32-bit ratio to widen
  else if (s->wide)
    launch_tables<0, true>(c, s);
  else
    launch_tables<0, false>(c, s);
  ExpandArgs a;
  a.node_flags = c->nrt_node_flags.as<uint8_t>();
  a.slot = s->slot.as<int32_t>();
  a.tile_n0 = s->tile_n0.as<int32_t>();
  a.tile_c0 = s->tile_c0.as<int32_t>();
  a.filt = s->filt.as<int32_t>();
  a.nrm = s->nrm.as<uint8_t>();
  a.S = s->Sc + s->Sp;
  a.Sc = s->Sc;
  a.Sp = s->Sp;
  a.qos = c->nrt_pod_qos.as<uint8_t>();
  a.flags = c->nrt_pod_flags.as<uint8_t>();
  a.n_init = c->nrt_pod_ninit.as<uint8_t>();
  a.n_app = c->nrt_pod_napp.as<uint8_t>();
  a.kind = c->nrt_pod_kind.as<uint8_t>();
  a.pod_vec = s->d_pod_vec.as<int32_t>();
  a.pod_tc = s->d_pod_tc.as<int32_t>();
  a.pod_tp = s->d_pod_tp.as<int32_t>();
  a.vecs = s->vecrec.as<VecRec>();
  a.Tc = s->Tc.as<uint16_t>();
  a.Tp = s->Tp.as<uint8_t>();
  a.upstream = c->upstream_mask();
  a.words = c->Npad / 64;
  a.N = c->N;
  a.Npad = c->Npad;
  a.P = c->P;
  dim3 grid((unsigned)s->n_tiles, (unsigned)((c->P + PT - 1) / PT));
  if (dtype == B200S_OUT_I64)
    nrt2_expand_kernel<4, 4, int64_t><<<grid, TILE, 0, c->stream>>>(a, o.scores.as<int64_t>(), o.feas.as<uint32_t>(),
                                                                    o.reasons.as<uint8_t>());
  else
    nrt2_expand_kernel<4, 4, uint8_t><<<grid, TILE, 0, c->stream>>>(a, o.scores.as<uint8_t>(), o.feas.as<uint32_t>(),
                                                                    o.reasons.as<uint8_t>());
  c->launches += 1;
  B200S_CUDA_TRY(c, cudaGetLastError());
  s->last_path = 2;
  return B200S_OK;
}

}  // namespace b200s

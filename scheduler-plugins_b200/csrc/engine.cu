// libb200sched: lifecycle, snapshot columns, pod batches, result fetch — the C-ABI of
// include/b200sched.h.  Kernels live in alloc.cu / trimaran.cu / nrt.cu / netoh.cu / combined.cu.
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>
#include <unordered_map>
#include <vector>

using namespace b200s;

namespace {

thread_local std::string tl_err;

struct Guard {
  b200s_ctx* c;
  std::unique_lock<std::mutex> lk;
  explicit Guard(b200s_ctx* ctx) : c(ctx), lk(ctx->mu) { cudaSetDevice(ctx->device); }
};

// Copies a host column of n elements into a device buffer padded to npad elements (pad = 0).
template <class T>
int upload_col(b200s_ctx* c, DevBuf& dst, size_t dst_off_elems, const T* src, int n, int npad) {
  T* d = dst.as<T>() + dst_off_elems;
  B200S_CUDA_TRY(c, cudaMemcpyAsync(d, src, sizeof(T) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  if (npad > n) B200S_CUDA_TRY(c, cudaMemsetAsync(d + n, 0, sizeof(T) * (size_t)(npad - n), c->stream));
  return B200S_OK;
}

// Row sanity for the score-table path of the Trimaran plugins: with finite, non-negative inputs every score lies in
// 0..100 (targetloadpacking.go:170-186; analysis.go:41-59 clamps) and fits a byte table; anything else (a NaN metric, a
// negative capacity) keeps the direct int64 kernels, which reproduce Go's MinInt64 / wrapped results.
bool tlp_rows_sane(const double* util, const int64_t* cap, const int64_t* missing, int n) {
  for (int i = 0; i < n; ++i)
    if (!(util[i] >= 0 && util[i] <= 1e12) || cap[i] < 0 || missing[i] < 0) return false;
  return true;
}
bool lvrb_rows_sane(const double* a, const double* b, const double* c2, const double* d, const int64_t* ac, const int64_t* am, int n) {
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(a[i]) || !std::isfinite(b[i]) || !std::isfinite(c2[i]) || !std::isfinite(d[i]) || ac[i] < 0 || am[i] < 0)
      return false;
  return true;
}

// distinct values of a pod column (pairs: a second column): row_of_pod + the distinct keys in first-seen order
void dedup_keys(const int64_t* k0, const int64_t* k1, int P, std::vector<int32_t>& row, std::vector<int64_t>& u0,
                std::vector<int64_t>& u1) {
  row.resize((size_t)P);
  u0.clear();
  u1.clear();
  size_t cap = 64;
  while (cap < (size_t)P * 2) cap <<= 1;
  std::vector<int32_t> table(cap, -1);
  for (int p = 0; p < P; ++p) {
    const uint64_t a = (uint64_t)k0[p], b = k1 ? (uint64_t)k1[p] : 0;
    uint64_t h = (a * 0x9e3779b97f4a7c15ull) ^ ((b + 0x7f4a7c15ull) * 0xbf58476d1ce4e5b9ull);
    h ^= h >> 29;
    size_t i = (size_t)h & (cap - 1);
    for (;;) {
      const int32_t id = table[i];
      if (id < 0) {
        table[i] = (int32_t)u0.size();
        row[(size_t)p] = (int32_t)u0.size();
        u0.push_back(k0[p]);
        if (k1) u1.push_back(k1[p]);
        break;
      }
      if (u0[(size_t)id] == k0[p] && (!k1 || u1[(size_t)id] == k1[p])) {
        row[(size_t)p] = id;
        break;
      }
      i = (i + 1) & (cap - 1);
    }
  }
}

// NRT thread-slot permutation.  Major key: the control-flow class (flags, zone count), so that pod-scope and
// container-scope nodes do not serialise inside a warp.  Minor key: how much the node's roomiest zone has left of
// resource slots 0 and 1 -- whether a request fits a zone is what decides Filter, so lanes of a warp tend to agree
// and the Score branch (taken only by surviving pairs) runs with full warps.  The permutation only shapes the
// schedule; results never depend on it.
uint64_t nrt_sort_key(uint8_t node_flags, uint8_t nz, int Z, int R, const uint8_t* zmask, const int64_t* avail,
                      size_t zstride, size_t rstride) {  // zmask[z * zstride], avail[(z * R + r) * rstride]
  int64_t cap[2] = {0, 0};
  for (int z = 0; z < Z && z < nz; ++z)
    for (int r = 0; r < 2 && r < R; ++r)
      if ((zmask[(size_t)z * zstride] >> r) & 1u) cap[r] = std::max(cap[r], avail[((size_t)z * R + r) * rstride]);
  const uint64_t q0 = (uint64_t)std::min<int64_t>(std::max<int64_t>(cap[0], 0) / 1000, 0xFFFFF);
  const uint64_t q1 = (uint64_t)std::min<int64_t>((std::max<int64_t>(cap[1], 0) / 1000) >> 28, 0xFFFFF);
  return ((uint64_t)(((uint32_t)node_flags << 8) | nz) << 40) | (q0 << 20) | q1;
}
inline uint32_t nrt_class_of(uint64_t key) { return (uint32_t)(key >> 40); }

int upload_nrt_perm(b200s_ctx* c) {
  const size_t np = c->Npad, n = (size_t)std::max(c->N, 0);
  // stable LSD radix sort of (key, node) over the 56 key bits, 4 passes of 14 bits
  std::vector<int32_t> perm(np), tmp(n);
  for (size_t i = 0; i < np; ++i) perm[i] = (int32_t)i;
  std::vector<uint32_t> start((1u << 14) + 1);
  for (int pass = 0; pass < 4; ++pass) {
    const int sh = 14 * pass;
    std::fill(start.begin(), start.end(), 0u);
    for (size_t i = 0; i < n; ++i) start[((c->nrt_key_h[i] >> sh) & 0x3FFF) + 1]++;
    for (size_t k = 1; k < start.size(); ++k) start[k] += start[k - 1];
    for (size_t i = 0; i < n; ++i) tmp[start[(c->nrt_key_h[perm[i]] >> sh) & 0x3FFF]++] = perm[i];
    std::copy(tmp.begin(), tmp.end(), perm.begin());
  }
  B200S_CUDA_TRY(c, c->nrt_perm.ensure(np * 4));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->nrt_perm.p, perm.data(), np * 4, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // perm dies at return
  c->nrt_perm_dirty = false;
  return B200S_OK;
}

// NetworkOverhead pair dictionary: id of the (region, zone) label pair, appended on first sight.
int32_t netoh_pair_of(b200s_ctx* c, uint16_t region, uint16_t zone) {
  const uint32_t key = ((uint32_t)region << 16) | zone;
  auto it = c->netoh_dict.find(key);
  if (it != c->netoh_dict.end()) return it->second;
  const int32_t id = (int32_t)c->netoh_pair_r_h.size();
  c->netoh_dict.emplace(key, id);
  c->netoh_pair_r_h.push_back(region);
  c->netoh_pair_z_h.push_back(zone);
  c->netoh_pairs_dirty = true;
  return id;
}

int upload_netoh_pairs(b200s_ctx* c) {
  if (c->netoh_pair_r_h.empty()) netoh_pair_of(c, 0, 0);
  const size_t nq = c->netoh_pair_r_h.size();
  c->netoh_NQ = (int)nq;
  B200S_CUDA_TRY(c, c->netoh_pair_r.ensure(nq * 2));
  B200S_CUDA_TRY(c, c->netoh_pair_z.ensure(nq * 2));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->netoh_pair_r.p, c->netoh_pair_r_h.data(), nq * 2, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->netoh_pair_z.p, c->netoh_pair_z_h.data(), nq * 2, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // pageable sources: do not let them change under the copy
  c->netoh_pairs_dirty = false;
  return B200S_OK;
}

// ---- snapshot patch: `count` rows of up to a few hundred columns, one copy + one scatter launch
struct PatchCol {
  void* dst;          // device column base (element 0 = node 0)
  uint32_t src_off;   // byte offset of the column's `count` values inside the staged block
  uint32_t elem;      // element size: 1, 2, 4 or 8
};

__global__ void patch_scatter_kernel(const unsigned char* __restrict__ stage, const PatchCol* __restrict__ cols,
                                     const int32_t* __restrict__ idx, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const PatchCol pc = cols[blockIdx.y];
  const size_t n = (size_t)idx[i];
  const unsigned char* src = stage + pc.src_off;
  switch (pc.elem) {
    case 8: static_cast<uint64_t*>(pc.dst)[n] = reinterpret_cast<const uint64_t*>(src)[i]; break;
    case 4: static_cast<uint32_t*>(pc.dst)[n] = reinterpret_cast<const uint32_t*>(src)[i]; break;
    case 2: static_cast<uint16_t*>(pc.dst)[n] = reinterpret_cast<const uint16_t*>(src)[i]; break;
    default: static_cast<uint8_t*>(pc.dst)[n] = src[i]; break;
  }
}

// OverReserve deduction (cache/store.go:129-160): one thread per (row, zone, resource) cell
__global__ void nrt_deduct_kernel(const int32_t* __restrict__ idx, const uint8_t* __restrict__ res_mask,
                                  const int64_t* __restrict__ deduct, int count, int Z, int R, size_t npad,
                                  const uint8_t* __restrict__ zmask, int64_t* __restrict__ avail) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count * Z * R) return;
  const int i = t / (Z * R), z = (t / R) % Z, r = t % R;
  const size_t n = (size_t)idx[i];
  if (!((zmask[(size_t)z * npad + n] >> r) & 1u) || !((res_mask[i] >> r) & 1u)) return;
  int64_t* a = avail + ((size_t)z * R + r) * npad + n;
  const int64_t q = deduct[(size_t)r * count + i], v = *a;
  *a = v < q ? 0 : v - q;
}

// Collects the columns of one patch call, packs them behind the index list and launches the scatter.
// Repeated indices: rows are de-duplicated on the host (last wins) so that the scatter has no write race.
struct Patch {
  b200s_ctx* c;
  int count;
  std::vector<int32_t> keep;  // positions of the surviving rows, ascending
  struct Item {
    void* dst;
    const void* src;
    uint32_t elem;
  };
  std::vector<Item> items;
  Patch(b200s_ctx* ctx, int n) : c(ctx), count(n) {}
  int prepare(const int32_t* node_idx) {
    if (count < 0 || (count > 0 && !node_idx)) return c->set_err(B200S_ERR_INVALID, "snapshot_patch: bad count / null node_idx");
    std::unordered_map<int32_t, int32_t> last;
    for (int i = 0; i < count; ++i) {
      if (node_idx[i] < 0 || node_idx[i] >= c->N) return c->set_err(B200S_ERR_INVALID, "snapshot_patch: node index out of range");
      last[node_idx[i]] = i;
    }
    keep.reserve(last.size());
    for (int i = 0; i < count; ++i)
      if (last[node_idx[i]] == i) keep.push_back(i);
    return B200S_OK;
  }
  template <class T>
  void add(DevBuf& dst, size_t dst_off_elems, const T* src) {
    items.push_back({dst.as<T>() + dst_off_elems, src, (uint32_t)sizeof(T)});
  }
  int run(const int32_t* node_idx) {
    const int m = (int)keep.size();
    if (m == 0 || items.empty()) return B200S_OK;
    auto up8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
    const size_t idx_bytes = up8((size_t)m * 4), desc_bytes = up8(items.size() * sizeof(PatchCol));
    size_t total = idx_bytes + desc_bytes;
    std::vector<PatchCol> desc(items.size());
    for (size_t k = 0; k < items.size(); ++k) {
      desc[k] = {items[k].dst, (uint32_t)total, items[k].elem};
      total += up8((size_t)m * items[k].elem);
    }
    if (total > ((size_t)1 << 31)) return c->set_err(B200S_ERR_INVALID, "snapshot_patch: too many rows for one call");
    if (total > c->patch_stage_cap) {
      if (c->patch_stage) cudaFreeHost(c->patch_stage);
      c->patch_stage = nullptr;
      c->patch_stage_cap = 0;
      B200S_CUDA_TRY(c, cudaHostAlloc(&c->patch_stage, total * 2, cudaHostAllocDefault));
      c->patch_stage_cap = total * 2;
    }
    B200S_CUDA_TRY(c, c->patch_dev.ensure(total));
    char* st = static_cast<char*>(c->patch_stage);
    int32_t* sidx = reinterpret_cast<int32_t*>(st);
    for (int j = 0; j < m; ++j) sidx[j] = node_idx[keep[j]];
    memcpy(st + idx_bytes, desc.data(), items.size() * sizeof(PatchCol));
    for (size_t k = 0; k < items.size(); ++k) {
      char* d = st + desc[k].src_off;
      const char* sp = static_cast<const char*>(items[k].src);
      const uint32_t e = items[k].elem;
      for (int j = 0; j < m; ++j) memcpy(d + (size_t)j * e, sp + (size_t)keep[j] * e, e);
    }
    B200S_CUDA_TRY(c, cudaMemcpyAsync(c->patch_dev.p, st, total, cudaMemcpyHostToDevice, c->stream));
    const unsigned char* dev = c->patch_dev.as<unsigned char>();
    dim3 grid((m + 255) / 256, (unsigned)items.size());
    patch_scatter_kernel<<<grid, 256, 0, c->stream>>>(dev, reinterpret_cast<const PatchCol*>(dev + idx_bytes),
                                                      reinterpret_cast<const int32_t*>(dev), m);
    B200S_CUDA_TRY(c, cudaGetLastError());
    // the staging buffer is re-used by the next patch call of this session
    B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->launches++;
    return B200S_OK;
  }
};

int require_patching(b200s_ctx* c, bool has, const char* what) {
  if (!c->snap_open || !c->snap_patching)
    return c->set_err(B200S_ERR_STATE, std::string(what) + " outside b200s_snapshot_patch_begin/commit");
  if (!has) return c->set_err(B200S_ERR_STATE, std::string(what) + ": these columns were never uploaded in full");
  return B200S_OK;
}

int require_open(b200s_ctx* c) {
  if (!c->snap_open) return c->set_err(B200S_ERR_STATE, "snapshot column set outside b200s_snapshot_begin/commit");
  return B200S_OK;
}

}  // namespace

namespace b200s {

int ensure_out(b200s_ctx* c, int plugin, int dtype, bool feas, bool reasons) {
  PluginOut& o = c->out[plugin];
  size_t elems = (size_t)c->P * (size_t)c->Npad;
  B200S_CUDA_TRY(c, o.scores.ensure(elems * (dtype == B200S_OUT_I64 ? 8 : 1)));
  if (feas) B200S_CUDA_TRY(c, o.feas.ensure((size_t)c->P * (size_t)(c->Npad / 64) * 8));
  if (reasons) B200S_CUDA_TRY(c, o.reasons.ensure(elems));
  o.dtype = dtype;
  o.P = c->P;
  o.has_feas = feas;
  o.has_reasons = reasons;
  o.valid = false;
  return B200S_OK;
}

}  // namespace b200s

extern "C" {

int b200s_version(void) { return B200S_VERSION; }

int b200s_init(int device, b200s_ctx** out) {
  if (!out) return B200S_ERR_INVALID;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    tl_err = std::string("b200s_init: no CUDA device (") + cudaGetErrorString(e) +
             "); the engine has no CPU fallback";
    return B200S_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) {
    tl_err = "b200s_init: device index out of range";
    return B200S_ERR_INVALID;
  }
  b200s_ctx* c = new (std::nothrow) b200s_ctx();
  if (!c) return B200S_ERR_NOMEM;
  c->device = device;
  e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    tl_err = std::string("b200s_init: ") + cudaGetErrorString(e);
    delete c;
    return B200S_ERR_CUDA;
  }
  *out = c;
  return B200S_OK;
}

void b200s_shutdown(b200s_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  comm_destroy(c);
  nrt2_destroy(c);
  if (c->comm_stream) {
    cudaStreamSynchronize(c->comm_stream);
    cudaStreamDestroy(c->comm_stream);
    for (int i = 0; i < 4; ++i) {
      cudaEventDestroy(c->ev_chunk[i]);
      cudaEventDestroy(c->ev_reduced[i]);
    }
    for (int i = 0; i < 2; ++i) {
      cudaEventDestroy(c->ev_params[i]);
      cudaEventDestroy(c->ev_norm_done[i]);
    }
  }
  if (c->ev_inputs) cudaEventDestroy(c->ev_inputs);
  c->pod_lo_alt.release();
  c->cycle_scratch.release();
  cycle_graph_free(c);
  c->pods_stage2.release();
  c->feasible_alt.release();
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  for (cudaEvent_t e : {c->ev_mark[0], c->ev_mark[1], c->ev_copied})
    if (e) cudaEventDestroy(e);
  if (c->small_bounce) cudaFreeHost(c->small_bounce);
  c->norm_params_alt.release();
  DevBuf* bufs[] = {&c->alloc_cols,      &c->alloc_raw,        &c->alloc_sorted_raw, &c->alloc_order,
                    &c->alloc_iota,      &c->sort_tmp,         &c->tlp_util,         &c->tlp_cap,
                    &c->tlp_missing,     &c->tlp_flags,        &c->lvrb_f64,         &c->lvrb_i64,
                    &c->lvrb_flags,      &c->nrt_node_flags,   &c->nrt_max_numa,     &c->nrt_nz,
                    &c->nrt_node_res_mask, &c->nrt_zone_res_mask, &c->nrt_avail,     &c->nrt_cost,
                    &c->netoh_region,    &c->netoh_zone,       &c->netoh_zone_cost,  &c->netoh_region_cost,
                    &c->feasible_in,     &c->tlp_pod_cpu,      &c->lvrb_req_cpu,     &c->lvrb_req_mem,
                    &c->nrt_pod_qos,     &c->nrt_pod_flags,    &c->nrt_pod_ninit,    &c->nrt_pod_napp,
                    &c->nrt_pod_kind,    &c->nrt_pod_req_mask, &c->nrt_pod_req,      &c->netoh_equal,
                    &c->netoh_dep_off,   &c->netoh_deps,       &c->pod_lo,           &c->pod_hi,
                    &c->norm_params,     &c->raw_scores,       &c->total,            &c->total_feas,
                    &c->topk_local,      &c->topk_all,         &c->topk_final,       &c->netoh_counts,
                    &c->netoh_pair_id,   &c->netoh_pair_r,     &c->netoh_pair_z,     &c->netoh_pair_cost,
                    &c->netoh_pair_sv,   &c->nrt_perm,         &c->topk_slices,      &c->peaks_util,
                    &c->peaks_cap,       &c->peaks_flags,      &c->peaks_k,          &c->lowrisk_f64,
                    &c->lowrisk_i64,     &c->lowrisk_flags,    &c->lowrisk_load,     &c->peaks_pod_cpu,
                    &c->lowrisk_pod,     &c->tlp_uniq,         &c->tlp_row,          &c->lvrb_uniq_cpu,
                    &c->lvrb_uniq_mem,   &c->lvrb_row,         &c->score_table};
  for (DevBuf* b : bufs) b->release();
  for (auto& o : c->out) {
    o.scores.release();
    o.feas.release();
    o.reasons.release();
  }
  for (auto& v : c->prof_pending)
    for (auto& pr : v) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
  c->pods_arena.release();
  if (c->pods_stage) cudaFreeHost(c->pods_stage);
  c->patch_dev.release();
  if (c->patch_stage) cudaFreeHost(c->patch_stage);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  for (cudaEvent_t e : {c->ev_kernels, c->ev_d2h, c->ev_h2d})
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char* b200s_last_error(b200s_ctx* c) {
  if (!c) return tl_err.c_str();
  std::lock_guard<std::mutex> lk(c->mu);
  tl_err = c->err;
  return tl_err.c_str();
}

void* b200s_stream(b200s_ctx* c) { return c ? (void*)c->stream : nullptr; }

int b200s_sync(b200s_ctx* c) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200S_OK;
}

uint64_t b200s_launch_count(b200s_ctx* c) { return c ? c->launches : 0; }

int32_t b200s_npad(b200s_ctx* c) { return c ? c->Npad : 0; }

int b200s_config_nrt_path(b200s_ctx* c, int path) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (path < B200S_NRT_PATH_AUTO || path > B200S_NRT_PATH_BATCHED) return c->set_err(B200S_ERR_INVALID, "config_nrt_path: unknown path");
  nrt2_set_force(c, path);
  return B200S_OK;
}

int b200s_nrt_last_path(b200s_ctx* c) {
  if (!c) return 0;
  Guard g(c);
  return nrt2_last_path(c);
}

const char* b200s_nrt_path_note(b200s_ctx* c) {
  if (!c) return "";
  Guard g(c);
  return nrt2_note(c);
}

int b200s_set_profiling(b200s_ctx* c, int on) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->profiling = on != 0;
  return B200S_OK;
}

static int timer_read(b200s_ctx* c, int slot, double* total_ms, uint64_t* launches) {
  Guard g(c);
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  double ms = 0;
  uint64_t n = 0;
  for (auto& pr : c->prof_pending[slot]) {
    float t = 0;
    if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) {
      ms += t;
      n++;
    }
    c->prof_pool.push_back(pr.first);
    c->prof_pool.push_back(pr.second);
  }
  c->prof_pending[slot].clear();
  *total_ms = ms;
  *launches = n;
  return B200S_OK;
}

int b200s_kernel_time(b200s_ctx* c, b200s_plugin plugin, double* total_ms, uint64_t* launches) {
  if (!c || (int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT || !total_ms || !launches) return B200S_ERR_INVALID;
  return timer_read(c, (int)plugin, total_ms, launches);
}

int b200s_phase_time(b200s_ctx* c, int phase, double* total_ms, uint64_t* count) {
  if (!c || phase < 0 || phase >= B200S_PHASE_COUNT || !total_ms || !count) return B200S_ERR_INVALID;
  return timer_read(c, B200S_PLUGIN_COUNT + phase, total_ms, count);
}

void* b200s_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void b200s_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------- snapshot
int b200s_snapshot_begin(b200s_ctx* c, uint64_t generation, int32_t n_nodes, int32_t node_offset,
                         int32_t n_nodes_global) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (n_nodes < 0 || node_offset < 0 || n_nodes_global < n_nodes + node_offset)
    return c->set_err(B200S_ERR_INVALID, "snapshot_begin: bad node counts");
  c->snap_open = true;
  c->snap_patching = false;
  c->snap_valid = false;
  c->pods_valid = false;
  c->gen = generation;
  c->N = n_nodes;
  c->Npad = round_up(n_nodes > 0 ? n_nodes : 1, B200S_NODE_ALIGN);
  c->node_off = node_offset;
  c->Nglobal = n_nodes_global;
  c->has_alloc = c->has_tlp = c->has_lvrb = c->has_nrt = c->has_netoh = c->has_peaks = c->has_lowrisk = false;
  for (auto& o : c->out) o.valid = false;
  c->total_valid = c->topk_valid = c->feas_valid = false;
  return B200S_OK;
}

int b200s_snapshot_allocatable(b200s_ctx* c, int32_t n_res, const int64_t* const* alloc) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (n_res < 1 || n_res > 16 || !alloc) return c->set_err(B200S_ERR_INVALID, "snapshot_allocatable: 1..16 resources");
  B200S_CUDA_TRY(c, c->alloc_cols.ensure((size_t)n_res * c->Npad * 8));
  for (int r = 0; r < n_res; ++r) {
    if (!alloc[r]) return c->set_err(B200S_ERR_INVALID, "snapshot_allocatable: null column");
    B200S_TRY(upload_col<int64_t>(c, c->alloc_cols, (size_t)r * c->Npad, alloc[r], c->N, c->Npad));
  }
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->alloc_R = n_res;
  c->has_alloc = true;
  return B200S_OK;
}

int b200s_snapshot_tlp(b200s_ctx* c, const double* util, const int64_t* cap, const int64_t* missing,
                       const uint8_t* flags) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!util || !cap || !missing || !flags) return c->set_err(B200S_ERR_INVALID, "snapshot_tlp: null column");
  B200S_CUDA_TRY(c, c->tlp_util.ensure((size_t)c->Npad * 8));
  B200S_CUDA_TRY(c, c->tlp_cap.ensure((size_t)c->Npad * 8));
  B200S_CUDA_TRY(c, c->tlp_missing.ensure((size_t)c->Npad * 8));
  B200S_CUDA_TRY(c, c->tlp_flags.ensure((size_t)c->Npad));
  B200S_TRY(upload_col<double>(c, c->tlp_util, 0, util, c->N, c->Npad));
  B200S_TRY(upload_col<int64_t>(c, c->tlp_cap, 0, cap, c->N, c->Npad));
  B200S_TRY(upload_col<int64_t>(c, c->tlp_missing, 0, missing, c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->tlp_flags, 0, flags, c->N, c->Npad));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->has_tlp = true;
  c->tlp_sane = tlp_rows_sane(util, cap, missing, c->N);
  return B200S_OK;
}

int b200s_snapshot_lvrb(b200s_ctx* c, const double* cpu_avg, const double* cpu_std, const double* mem_avg,
                        const double* mem_std, const int64_t* alloc_cpu, const int64_t* alloc_mem,
                        const uint8_t* flags) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!cpu_avg || !cpu_std || !mem_avg || !mem_std || !alloc_cpu || !alloc_mem || !flags)
    return c->set_err(B200S_ERR_INVALID, "snapshot_lvrb: null column");
  size_t np = c->Npad;
  B200S_CUDA_TRY(c, c->lvrb_f64.ensure(4 * np * 8));
  B200S_CUDA_TRY(c, c->lvrb_i64.ensure(2 * np * 8));
  B200S_CUDA_TRY(c, c->lvrb_flags.ensure(np));
  B200S_TRY(upload_col<double>(c, c->lvrb_f64, 0 * np, cpu_avg, c->N, c->Npad));
  B200S_TRY(upload_col<double>(c, c->lvrb_f64, 1 * np, cpu_std, c->N, c->Npad));
  B200S_TRY(upload_col<double>(c, c->lvrb_f64, 2 * np, mem_avg, c->N, c->Npad));
  B200S_TRY(upload_col<double>(c, c->lvrb_f64, 3 * np, mem_std, c->N, c->Npad));
  B200S_TRY(upload_col<int64_t>(c, c->lvrb_i64, 0 * np, alloc_cpu, c->N, c->Npad));
  B200S_TRY(upload_col<int64_t>(c, c->lvrb_i64, 1 * np, alloc_mem, c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->lvrb_flags, 0, flags, c->N, c->Npad));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->has_lvrb = true;
  c->lvrb_sane = lvrb_rows_sane(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, c->N);
  c->lvrb_sane = lvrb_rows_sane(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, c->N);
  return B200S_OK;
}

int b200s_snapshot_peaks(b200s_ctx* c, const double* util, const int64_t* cap, const uint8_t* flags, const double* k1,
                         const double* k2) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!util || !cap || !flags || !k1 || !k2) return c->set_err(B200S_ERR_INVALID, "snapshot_peaks: null column");
  const size_t np = c->Npad;
  B200S_CUDA_TRY(c, c->peaks_util.ensure(np * 8));
  B200S_CUDA_TRY(c, c->peaks_cap.ensure(np * 8));
  B200S_CUDA_TRY(c, c->peaks_flags.ensure(np));
  B200S_CUDA_TRY(c, c->peaks_k.ensure(2 * np * 8));
  B200S_TRY(upload_col<double>(c, c->peaks_util, 0, util, c->N, c->Npad));
  B200S_TRY(upload_col<int64_t>(c, c->peaks_cap, 0, cap, c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->peaks_flags, 0, flags, c->N, c->Npad));
  B200S_TRY(upload_col<double>(c, c->peaks_k, 0 * np, k1, c->N, c->Npad));
  B200S_TRY(upload_col<double>(c, c->peaks_k, 1 * np, k2, c->N, c->Npad));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->has_peaks = true;
  return B200S_OK;
}

int b200s_snapshot_low_risk(b200s_ctx* c, const double* cpu_avg, const double* cpu_std, const double* mem_avg,
                            const double* mem_std, const int64_t* alloc_cpu, const int64_t* alloc_mem,
                            const uint8_t* flags, const int64_t* node_req_cpu, const int64_t* node_req_mem,
                            const int64_t* node_lim_cpu, const int64_t* node_lim_mem) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!cpu_avg || !cpu_std || !mem_avg || !mem_std || !alloc_cpu || !alloc_mem || !flags || !node_req_cpu ||
      !node_req_mem || !node_lim_cpu || !node_lim_mem)
    return c->set_err(B200S_ERR_INVALID, "snapshot_low_risk: null column");
  const size_t np = c->Npad;
  B200S_CUDA_TRY(c, c->lowrisk_f64.ensure(4 * np * 8));
  B200S_CUDA_TRY(c, c->lowrisk_i64.ensure(6 * np * 8));
  B200S_CUDA_TRY(c, c->lowrisk_flags.ensure(np));
  const double* f[4] = {cpu_avg, cpu_std, mem_avg, mem_std};
  const int64_t* iv[6] = {alloc_cpu, alloc_mem, node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem};
  for (int k = 0; k < 4; ++k) B200S_TRY(upload_col<double>(c, c->lowrisk_f64, k * np, f[k], c->N, c->Npad));
  for (int k = 0; k < 6; ++k) B200S_TRY(upload_col<int64_t>(c, c->lowrisk_i64, k * np, iv[k], c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->lowrisk_flags, 0, flags, c->N, c->Npad));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->has_lowrisk = true;
  return B200S_OK;
}

int b200s_snapshot_nrt(b200s_ctx* c, const b200s_nrt_nodes* nn) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!nn) return c->set_err(B200S_ERR_INVALID, "snapshot_nrt: null");
  int Z = nn->n_zones, R = nn->n_res;
  if (Z < 1 || Z > B200S_NRT_MAX_ZONES || R < 1 || R > B200S_NRT_MAX_RES)
    return c->set_err(B200S_ERR_UNSUPPORTED, "snapshot_nrt: zones/resources outside 1..8");
  if (!nn->res_flags || !nn->node_flags || !nn->max_numa || !nn->n_zones_node || !nn->node_res_mask ||
      !nn->zone_res_mask || !nn->avail)
    return c->set_err(B200S_ERR_INVALID, "snapshot_nrt: null column");
  size_t np = c->Npad;
  memcpy(c->nrt_res_flags, nn->res_flags, R);
  B200S_CUDA_TRY(c, c->nrt_node_flags.ensure(np));
  B200S_CUDA_TRY(c, c->nrt_max_numa.ensure(np * 2));
  B200S_CUDA_TRY(c, c->nrt_nz.ensure(np));
  B200S_CUDA_TRY(c, c->nrt_node_res_mask.ensure(np));
  B200S_CUDA_TRY(c, c->nrt_zone_res_mask.ensure((size_t)Z * np));
  B200S_CUDA_TRY(c, c->nrt_avail.ensure((size_t)Z * R * np * 8));
  B200S_TRY(upload_col<uint8_t>(c, c->nrt_node_flags, 0, nn->node_flags, c->N, c->Npad));
  B200S_TRY(upload_col<uint16_t>(c, c->nrt_max_numa, 0, nn->max_numa, c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->nrt_nz, 0, nn->n_zones_node, c->N, c->Npad));
  B200S_TRY(upload_col<uint8_t>(c, c->nrt_node_res_mask, 0, nn->node_res_mask, c->N, c->Npad));
  for (int z = 0; z < Z; ++z)
    B200S_TRY(upload_col<uint8_t>(c, c->nrt_zone_res_mask, (size_t)z * np, nn->zone_res_mask + (size_t)z * c->N,
                                  c->N, c->Npad));
  for (int i = 0; i < Z * R; ++i)
    B200S_TRY(upload_col<int64_t>(c, c->nrt_avail, (size_t)i * np, nn->avail + (size_t)i * c->N, c->N, c->Npad));
  c->nrt_key_h.resize((size_t)std::max(c->N, 0));
  for (int i = 0; i < c->N; ++i)
    c->nrt_key_h[i] = nrt_sort_key(nn->node_flags[i], nn->n_zones_node[i], Z, R, nn->zone_res_mask + i, nn->avail + i,
                                   (size_t)c->N, (size_t)c->N);
  B200S_TRY(upload_nrt_perm(c));
  c->nrt_has_cost = nn->cost != nullptr;
  if (nn->cost) {
    B200S_CUDA_TRY(c, c->nrt_cost.ensure((size_t)Z * Z * np * 4));
    for (int i = 0; i < Z * Z; ++i)
      B200S_TRY(upload_col<int32_t>(c, c->nrt_cost, (size_t)i * np, nn->cost + (size_t)i * c->N, c->N, c->Npad));
  }
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->nrt_Z = Z;
  c->nrt_R = R;
  c->has_nrt = true;
  nrt2_on_snapshot_full(c, nn);
  return B200S_OK;
}

int b200s_snapshot_network_overhead(b200s_ctx* c, const uint16_t* region_id, const uint16_t* zone_id,
                                    int32_t n_names, const int64_t* zone_cost, const int64_t* region_cost) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_open(c));
  if (!region_id || !zone_id || !zone_cost || !region_cost || n_names < 1 || n_names > 4096)
    return c->set_err(B200S_ERR_INVALID, "snapshot_network_overhead: bad arguments");
  size_t np = c->Npad, kk = (size_t)n_names * n_names;
  for (int i = 0; i < c->N; ++i)  // the ids index the two cost tables on the device
    if (region_id[i] >= n_names || zone_id[i] >= n_names)
      return c->set_err(B200S_ERR_INVALID, "snapshot_network_overhead: label id outside the name dictionary");
  B200S_CUDA_TRY(c, c->netoh_region.ensure(np * 2));
  B200S_CUDA_TRY(c, c->netoh_zone.ensure(np * 2));
  B200S_CUDA_TRY(c, c->netoh_zone_cost.ensure(kk * 8));
  B200S_CUDA_TRY(c, c->netoh_region_cost.ensure(kk * 8));
  B200S_TRY(upload_col<uint16_t>(c, c->netoh_region, 0, region_id, c->N, c->Npad));
  B200S_TRY(upload_col<uint16_t>(c, c->netoh_zone, 0, zone_id, c->N, c->Npad));
  {  // pair dictionary (host side of the flattening, O(N))
    std::vector<int32_t> pid(np, 0);
    c->netoh_dict.clear();
    c->netoh_pair_r_h.clear();
    c->netoh_pair_z_h.clear();
    for (int i = 0; i < c->N; ++i) pid[i] = netoh_pair_of(c, region_id[i], zone_id[i]);
    B200S_CUDA_TRY(c, c->netoh_pair_id.ensure(np * 4));
    B200S_CUDA_TRY(c, cudaMemcpyAsync(c->netoh_pair_id.p, pid.data(), np * 4, cudaMemcpyHostToDevice, c->stream));
    B200S_TRY(upload_netoh_pairs(c));  // synchronises: pid dies at the end of this scope
  }
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->netoh_zone_cost.p, zone_cost, kk * 8, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->netoh_region_cost.p, region_cost, kk * 8, cudaMemcpyHostToDevice, c->stream));
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  c->netoh_K = n_names;
  c->has_netoh = true;
  return B200S_OK;
}

int b200s_snapshot_commit(b200s_ctx* c) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->snap_open) return c->set_err(B200S_ERR_STATE, "snapshot_commit without begin");
  if (c->has_nrt && c->nrt_perm_dirty) B200S_TRY(upload_nrt_perm(c));
  if (c->has_netoh && c->netoh_pairs_dirty) B200S_TRY(upload_netoh_pairs(c));
  c->snap_open = false;
  c->snap_patching = false;
  c->snap_valid = true;
  c->snap_serial++;
  B200S_CUDA_TRY(c, cudaEventRecord(c->ev_inputs, c->stream));
  return B200S_OK;
}

// ---------------------------------------------------------------- incremental snapshot
int b200s_snapshot_patch_begin(b200s_ctx* c, uint64_t generation) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (c->snap_open) return c->set_err(B200S_ERR_STATE, "snapshot_patch_begin: a snapshot is already open");
  if (!c->snap_valid) return c->set_err(B200S_ERR_STATE, "snapshot_patch_begin: no committed snapshot to patch");
  c->snap_open = true;
  c->snap_patching = true;
  c->snap_valid = false;
  c->gen = generation;
  for (auto& o : c->out) o.valid = false;
  c->total_valid = c->topk_valid = c->feas_valid = false;
  c->netoh_raw_P = -1;
  return B200S_OK;
}

int b200s_snapshot_patch_allocatable(b200s_ctx* c, int32_t count, const int32_t* node_idx, int32_t n_res,
                                     const int64_t* const* alloc) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_alloc, "snapshot_patch_allocatable"));
  if (n_res != c->alloc_R || !alloc) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_allocatable: resource count differs from the snapshot");
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  for (int r = 0; r < n_res; ++r) {
    if (!alloc[r]) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_allocatable: null column");
    p.add<int64_t>(c->alloc_cols, (size_t)r * c->Npad, alloc[r]);
  }
  return p.run(node_idx);
}

int b200s_snapshot_patch_tlp(b200s_ctx* c, int32_t count, const int32_t* node_idx, const double* util,
                             const int64_t* cap, const int64_t* missing, const uint8_t* flags) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_tlp, "snapshot_patch_tlp"));
  if (!util || !cap || !missing || !flags) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_tlp: null column");
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  p.add<double>(c->tlp_util, 0, util);
  p.add<int64_t>(c->tlp_cap, 0, cap);
  p.add<int64_t>(c->tlp_missing, 0, missing);
  p.add<uint8_t>(c->tlp_flags, 0, flags);
  c->tlp_sane = c->tlp_sane && tlp_rows_sane(util, cap, missing, count);
  return p.run(node_idx);
}

int b200s_snapshot_patch_lvrb(b200s_ctx* c, int32_t count, const int32_t* node_idx, const double* cpu_avg,
                              const double* cpu_std, const double* mem_avg, const double* mem_std,
                              const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_lvrb, "snapshot_patch_lvrb"));
  if (!cpu_avg || !cpu_std || !mem_avg || !mem_std || !alloc_cpu || !alloc_mem || !flags)
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_lvrb: null column");
  const size_t np = c->Npad;
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  p.add<double>(c->lvrb_f64, 0 * np, cpu_avg);
  p.add<double>(c->lvrb_f64, 1 * np, cpu_std);
  p.add<double>(c->lvrb_f64, 2 * np, mem_avg);
  p.add<double>(c->lvrb_f64, 3 * np, mem_std);
  p.add<int64_t>(c->lvrb_i64, 0 * np, alloc_cpu);
  p.add<int64_t>(c->lvrb_i64, 1 * np, alloc_mem);
  p.add<uint8_t>(c->lvrb_flags, 0, flags);
  c->lvrb_sane = c->lvrb_sane && lvrb_rows_sane(cpu_avg, cpu_std, mem_avg, mem_std, alloc_cpu, alloc_mem, count);
  return p.run(node_idx);
}

int b200s_snapshot_patch_peaks(b200s_ctx* c, int32_t count, const int32_t* node_idx, const double* util,
                               const int64_t* cap, const uint8_t* flags, const double* k1, const double* k2) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_peaks, "snapshot_patch_peaks"));
  if (!util || !cap || !flags || !k1 || !k2) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_peaks: null column");
  const size_t np = c->Npad;
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  p.add<double>(c->peaks_util, 0, util);
  p.add<int64_t>(c->peaks_cap, 0, cap);
  p.add<uint8_t>(c->peaks_flags, 0, flags);
  p.add<double>(c->peaks_k, 0 * np, k1);
  p.add<double>(c->peaks_k, 1 * np, k2);
  return p.run(node_idx);
}

int b200s_snapshot_patch_low_risk(b200s_ctx* c, int32_t count, const int32_t* node_idx, const double* cpu_avg,
                                  const double* cpu_std, const double* mem_avg, const double* mem_std,
                                  const int64_t* alloc_cpu, const int64_t* alloc_mem, const uint8_t* flags,
                                  const int64_t* node_req_cpu, const int64_t* node_req_mem, const int64_t* node_lim_cpu,
                                  const int64_t* node_lim_mem) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_lowrisk, "snapshot_patch_low_risk"));
  if (!cpu_avg || !cpu_std || !mem_avg || !mem_std || !alloc_cpu || !alloc_mem || !flags || !node_req_cpu ||
      !node_req_mem || !node_lim_cpu || !node_lim_mem)
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_low_risk: null column");
  const size_t np = c->Npad;
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  const double* f[4] = {cpu_avg, cpu_std, mem_avg, mem_std};
  const int64_t* iv[6] = {alloc_cpu, alloc_mem, node_req_cpu, node_req_mem, node_lim_cpu, node_lim_mem};
  for (int k = 0; k < 4; ++k) p.add<double>(c->lowrisk_f64, k * np, f[k]);
  for (int k = 0; k < 6; ++k) p.add<int64_t>(c->lowrisk_i64, k * np, iv[k]);
  p.add<uint8_t>(c->lowrisk_flags, 0, flags);
  return p.run(node_idx);  // the per-node risk columns are re-derived at the next eval (keyed by the snapshot serial)
}

int b200s_snapshot_patch_nrt(b200s_ctx* c, int32_t count, const int32_t* node_idx, const b200s_nrt_nodes* nn) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_nrt, "snapshot_patch_nrt"));
  if (!nn) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt: null");
  const int Z = c->nrt_Z, R = c->nrt_R;
  if (nn->n_zones != Z || nn->n_res != R)
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt: zone/resource shape differs from the snapshot");
  if (!nn->node_flags || !nn->max_numa || !nn->n_zones_node || !nn->node_res_mask || !nn->zone_res_mask || !nn->avail)
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt: null column");
  if ((nn->cost != nullptr) != c->nrt_has_cost)
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt: cost columns must match the snapshot");
  const size_t np = c->Npad, cnt = (size_t)std::max(count, 0);
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  p.add<uint8_t>(c->nrt_node_flags, 0, nn->node_flags);
  p.add<uint16_t>(c->nrt_max_numa, 0, nn->max_numa);
  p.add<uint8_t>(c->nrt_nz, 0, nn->n_zones_node);
  p.add<uint8_t>(c->nrt_node_res_mask, 0, nn->node_res_mask);
  for (int z = 0; z < Z; ++z) p.add<uint8_t>(c->nrt_zone_res_mask, (size_t)z * np, nn->zone_res_mask + (size_t)z * cnt);
  for (int i = 0; i < Z * R; ++i) p.add<int64_t>(c->nrt_avail, (size_t)i * np, nn->avail + (size_t)i * cnt);
  if (nn->cost)
    for (int i = 0; i < Z * Z; ++i) p.add<int32_t>(c->nrt_cost, (size_t)i * np, nn->cost + (size_t)i * cnt);
  B200S_TRY(p.run(node_idx));
  for (int i : p.keep) {  // sort keys behind the thread permutation
    const uint64_t key = nrt_sort_key(nn->node_flags[i], nn->n_zones_node[i], Z, R, nn->zone_res_mask + i, nn->avail + i,
                                      cnt, cnt);
    // a node that changed class must move (warps are class-pure); a node whose free capacity drifted stays where
    // it is until the next full upload -- the order within a class is only a hint
    if (nrt_class_of(c->nrt_key_h[node_idx[i]]) != nrt_class_of(key)) {
      c->nrt_perm_dirty = true;
      nrt2_on_class_change(c);
    }
    c->nrt_key_h[node_idx[i]] = key;
  }
  nrt2_on_patch_rows(c, count, nn);
  return B200S_OK;
}

int b200s_snapshot_patch_nrt_deduct(b200s_ctx* c, int32_t count, const int32_t* node_idx, const uint8_t* res_mask,
                                    const int64_t* deduct) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_nrt, "snapshot_patch_nrt_deduct"));
  if (count < 0 || (count > 0 && (!node_idx || !res_mask || !deduct)))
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt_deduct: bad count / null column");
  if (count == 0) return B200S_OK;
  std::vector<int32_t> seen(node_idx, node_idx + count);
  std::sort(seen.begin(), seen.end());
  if (seen.front() < 0 || seen.back() >= c->N || std::adjacent_find(seen.begin(), seen.end()) != seen.end())
    return c->set_err(B200S_ERR_INVALID, "snapshot_patch_nrt_deduct: node index out of range or repeated");
  const int Z = c->nrt_Z, R = c->nrt_R;
  auto up8 = [](size_t v) { return (v + 7) & ~(size_t)7; };
  const size_t o_idx = 0, o_mask = up8((size_t)count * 4), o_ded = o_mask + up8((size_t)count),
               total = o_ded + (size_t)R * count * 8;
  if (total > c->patch_stage_cap) {
    if (c->patch_stage) cudaFreeHost(c->patch_stage);
    c->patch_stage = nullptr;
    c->patch_stage_cap = 0;
    B200S_CUDA_TRY(c, cudaHostAlloc(&c->patch_stage, total * 2, cudaHostAllocDefault));
    c->patch_stage_cap = total * 2;
  }
  B200S_CUDA_TRY(c, c->patch_dev.ensure(total));
  char* st = static_cast<char*>(c->patch_stage);
  memcpy(st + o_idx, node_idx, (size_t)count * 4);
  memcpy(st + o_mask, res_mask, (size_t)count);
  memcpy(st + o_ded, deduct, (size_t)R * count * 8);
  B200S_CUDA_TRY(c, cudaMemcpyAsync(c->patch_dev.p, st, total, cudaMemcpyHostToDevice, c->stream));
  const char* dev = c->patch_dev.as<char>();
  const int cells = count * Z * R;
  nrt_deduct_kernel<<<(cells + 255) / 256, 256, 0, c->stream>>>(
      reinterpret_cast<const int32_t*>(dev + o_idx), reinterpret_cast<const uint8_t*>(dev + o_mask),
      reinterpret_cast<const int64_t*>(dev + o_ded), count, Z, R, (size_t)c->Npad, c->nrt_zone_res_mask.as<uint8_t>(),
      c->nrt_avail.as<int64_t>());
  B200S_CUDA_TRY(c, cudaGetLastError());
  B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));  // the staging block is re-used by the next patch call
  c->launches++;
  nrt2_on_deduct(c, count, deduct);
  return B200S_OK;
}

int b200s_snapshot_patch_network_overhead(b200s_ctx* c, int32_t count, const int32_t* node_idx,
                                          const uint16_t* region_id, const uint16_t* zone_id) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  B200S_TRY(require_patching(c, c->has_netoh, "snapshot_patch_network_overhead"));
  if (!region_id || !zone_id) return c->set_err(B200S_ERR_INVALID, "snapshot_patch_network_overhead: null column");
  Patch p(c, count);
  B200S_TRY(p.prepare(node_idx));
  for (int i : p.keep)
    if (region_id[i] >= c->netoh_K || zone_id[i] >= c->netoh_K)
      return c->set_err(B200S_ERR_INVALID, "snapshot_patch_network_overhead: label id outside the name dictionary");
  std::vector<int32_t> pid((size_t)std::max(count, 1), 0);
  for (int i : p.keep) pid[i] = netoh_pair_of(c, region_id[i], zone_id[i]);
  p.add<uint16_t>(c->netoh_region, 0, region_id);
  p.add<uint16_t>(c->netoh_zone, 0, zone_id);
  p.add<int32_t>(c->netoh_pair_id, 0, pid.data());
  return p.run(node_idx);
}

// ---------------------------------------------------------------- plugin args
int b200s_config_allocatable(b200s_ctx* c, int mode, int32_t n_res, const int64_t* weights) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (n_res < 1 || n_res > 16 || !weights) return c->set_err(B200S_ERR_INVALID, "config_allocatable: 1..16 resources");
  // validateResources: allocatable.go:53-61 — weights must be positive.
  for (int r = 0; r < n_res; ++r)
    if (weights[r] <= 0) return c->set_err(B200S_ERR_INVALID, "resource Weight should be a positive value");
  c->alloc_mode = mode;
  c->alloc_cfg_R = n_res;
  for (int r = 0; r < n_res; ++r) c->alloc_w[r] = weights[r];
  c->alloc_cfg = true;
  c->alloc_cfg_gen++;
  return B200S_OK;
}

int b200s_config_tlp(b200s_ctx* c, int64_t target) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->tlp_target = target;
  c->tlp_cfg = true;
  return B200S_OK;
}

int b200s_config_lvrb(b200s_ctx* c, double margin, double sens) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->lvrb_margin = margin;
  c->lvrb_sens = sens;
  c->lvrb_cfg = true;
  return B200S_OK;
}

int b200s_config_low_risk(b200s_ctx* c, int64_t window, double w_cpu, double w_mem) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (window <= 0) return c->set_err(B200S_ERR_INVALID, "config_low_risk: smoothingWindowSize must be positive");
  if (!(w_cpu >= 0 && w_cpu <= 1) || !(w_mem >= 0 && w_mem <= 1))
    return c->set_err(B200S_ERR_INVALID, "config_low_risk: riskLimitWeights must be in [0, 1]");
  c->lowrisk_window = window;
  c->lowrisk_w_cpu = w_cpu;
  c->lowrisk_w_mem = w_mem;
  c->lowrisk_cfg = true;
  c->lowrisk_cfg_gen++;
  c->out[B200S_PLUGIN_LOW_RISK].valid = false;
  return B200S_OK;
}

int b200s_config_network_overhead(b200s_ctx* c, int want_counts, int apply_own_filter) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->netoh_want_counts = want_counts != 0;
  c->netoh_apply_filter = apply_own_filter != 0;
  return B200S_OK;
}

int b200s_config_nrt(b200s_ctx* c, int strategy, int32_t n_res, const int64_t* weights) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (strategy < 0 || strategy > B200S_NRT_LEAST_NUMA_NODES)
    return c->set_err(B200S_ERR_INVALID, "illegal scoring strategy found");  // score.go:138
  if (n_res < 0 || n_res > B200S_NRT_MAX_RES) return c->set_err(B200S_ERR_UNSUPPORTED, "config_nrt: > 8 resources");
  c->nrt_strategy = strategy;
  for (int r = 0; r < B200S_NRT_MAX_RES; ++r) c->nrt_w[r] = 1;
  for (int r = 0; r < n_res; ++r) c->nrt_w[r] = (weights && weights[r] >= 1) ? weights[r] : 1;  // score.go:49-60
  c->nrt_cfg = true;
  c->nrt_cfg_gen++;
  return B200S_OK;
}

// ---------------------------------------------------------------- pods
namespace {

// Collects the columns of one pod batch; small ones are packed into one pinned staging buffer and travel
// as a single cudaMemcpyAsync into the pod arena, large ones (the feasibility words of a big batch) go direct.
struct PodUpload {
  struct Item {
    DevBuf* d;
    const void* s;
    size_t bytes;
  };
  std::vector<Item> items;
  void add(DevBuf* d, const void* s, size_t bytes) { items.push_back({d, s, bytes}); }
  int flush(b200s_ctx* c) {
    // asynchronous uploads stage more themselves: a copy from the caller's pageable memory would wait for the stream
    const bool dbl = c->async_upload && !c->hold_upload;
    const size_t kSmall = dbl ? (size_t)4 << 20 : (size_t)64 * 1024;
    constexpr size_t kAlign = 256;
    size_t total = 0;
    for (auto& it : items)
      if (it.bytes <= kSmall) total += (it.bytes + kAlign - 1) / kAlign * kAlign;
    // held for the cycle graph: the copy is rounded up to 4 KiB so that its size rarely changes between cycles
    const size_t need = c->hold_upload ? (total + 4095) / 4096 * 4096 : total;
    void* stage = nullptr;
    if (total > 0 && dbl) {
      B200S_CUDA_TRY(c, c->pods_stage2.acquire(need, &stage));
      B200S_CUDA_TRY(c, c->pods_arena.ensure(need));
    } else if (total > 0) {
      if (need > c->pods_stage_cap) {
        if (c->pods_stage) cudaFreeHost(c->pods_stage);
        c->pods_stage = nullptr;
        c->pods_stage_cap = 0;
        B200S_CUDA_TRY(c, cudaHostAlloc(&c->pods_stage, need * 2, cudaHostAllocDefault));
        c->pods_stage_cap = need * 2;
      }
      B200S_CUDA_TRY(c, c->pods_arena.ensure(need));
      stage = c->pods_stage;
    }
    size_t off = 0;
    for (auto& it : items) {
      if (it.bytes <= kSmall) {
        memcpy(static_cast<char*>(stage) + off, it.s, it.bytes);
        it.d->view(static_cast<char*>(c->pods_arena.p) + off);
        off += (it.bytes + kAlign - 1) / kAlign * kAlign;
      } else if (dbl && it.d == &c->feasible_in && c->h2d_stream) {
        // the buffer the batch before the previous one used: free once everything queued before the previous upload ran
        std::swap(c->feasible_in, c->feasible_alt);
        B200S_CUDA_TRY(c, c->feasible_in.ensure(it.bytes));
        B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->h2d_stream, c->ev_mark[(c->upload_seq - 1) & 1], 0));
        B200S_CUDA_TRY(c, cudaMemcpyAsync(c->feasible_in.p, it.s, it.bytes, cudaMemcpyHostToDevice, c->h2d_stream));
        B200S_CUDA_TRY(c, cudaEventRecord(c->ev_copied, c->h2d_stream));
        B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_copied, 0));
      } else {
        B200S_CUDA_TRY(c, it.d->ensure(it.bytes));
        B200S_CUDA_TRY(c, cudaMemcpyAsync(it.d->p, it.s, it.bytes, cudaMemcpyHostToDevice, c->stream));
      }
    }
    c->held_bytes = 0;
    if (total > 0 && c->hold_upload) {  // b200s_schedule_batch: the copy is the first node of the cycle graph
      c->held_dst = c->pods_arena.p;
      c->held_src = stage;
      c->held_bytes = need;
    } else if (total > 0) {
      B200S_CUDA_TRY(c, cudaMemcpyAsync(c->pods_arena.p, stage, total, cudaMemcpyHostToDevice, c->stream));
      if (dbl) B200S_CUDA_TRY(c, c->pods_stage2.commit(c->stream));
    }
    return B200S_OK;
  }
};

}  // namespace

static int pods_upload_locked(b200s_ctx* c, const b200s_pod_batch* b) {
  if (!c->snap_valid) return c->set_err(B200S_ERR_STATE, "pods_upload: no committed snapshot");
  if (!b || b->n_pods < 0) return c->set_err(B200S_ERR_INVALID, "pods_upload: bad batch");
  int P = b->n_pods;
  c->pods_valid = false;
  c->P = P;
  if (c->async_upload && !c->hold_upload) {  // marks "everything queued for the earlier batches" on the engine stream
    if (!c->h2d_stream) {
      B200S_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
      for (cudaEvent_t* e : {&c->ev_mark[0], &c->ev_mark[1], &c->ev_copied})
        B200S_CUDA_TRY(c, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    c->upload_seq++;
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_mark[c->upload_seq & 1], c->stream));
  }
  size_t words = (size_t)(c->Npad / 64);
  PodUpload up;
  c->has_feasible = b->feasible != nullptr;
  if (b->feasible && P > 0) up.add(&c->feasible_in, b->feasible, (size_t)P * words * 8);
  c->has_tlp_pods = b->tlp_pod_cpu_milli != nullptr;
  if (c->has_tlp_pods && P > 0) up.add(&c->tlp_pod_cpu, b->tlp_pod_cpu_milli, (size_t)P * 8);
  c->has_lvrb_pods = b->lvrb_req_cpu_milli && b->lvrb_req_mem_bytes;
  if (c->has_lvrb_pods && P > 0) {
    up.add(&c->lvrb_req_cpu, b->lvrb_req_cpu_milli, (size_t)P * 8);
    up.add(&c->lvrb_req_mem, b->lvrb_req_mem_bytes, (size_t)P * 8);
  }
  // distinct request keys of a large batch (score-table path of TargetLoadPacking / LoadVariationRiskBalancing)
  c->tlp_U = c->lvrb_U = 0;
  constexpr int kDedupMinPods = 256;
  if (c->has_tlp_pods && P >= kDedupMinPods) {
    bool ok = true;
    for (int p = 0; p < P && ok; ++p) ok = b->tlp_pod_cpu_milli[p] >= 0 && b->tlp_pod_cpu_milli[p] < ((int64_t)1 << 53);
    if (ok) {
      std::vector<int64_t> none;
      dedup_keys(b->tlp_pod_cpu_milli, nullptr, P, c->tlp_row_h, c->tlp_uniq_h, none);
      c->tlp_U = (int)c->tlp_uniq_h.size();
      up.add(&c->tlp_uniq, c->tlp_uniq_h.data(), (size_t)c->tlp_U * 8);
      up.add(&c->tlp_row, c->tlp_row_h.data(), (size_t)P * 4);
    }
  }
  if (c->has_lvrb_pods && P >= kDedupMinPods) {
    dedup_keys(b->lvrb_req_cpu_milli, b->lvrb_req_mem_bytes, P, c->lvrb_row_h, c->lvrb_uniq_cpu_h, c->lvrb_uniq_mem_h);
    c->lvrb_U = (int)c->lvrb_uniq_cpu_h.size();
    up.add(&c->lvrb_uniq_cpu, c->lvrb_uniq_cpu_h.data(), (size_t)c->lvrb_U * 8);
    up.add(&c->lvrb_uniq_mem, c->lvrb_uniq_mem_h.data(), (size_t)c->lvrb_U * 8);
    up.add(&c->lvrb_row, c->lvrb_row_h.data(), (size_t)P * 4);
  }
  c->has_peaks_pods = b->peaks_pod_cpu_milli != nullptr;
  if (c->has_peaks_pods && P > 0) up.add(&c->peaks_pod_cpu, b->peaks_pod_cpu_milli, (size_t)P * 8);
  c->has_lowrisk_pods = b->low_risk_pod != nullptr;
  if (c->has_lowrisk_pods && P > 0) up.add(&c->lowrisk_pod, b->low_risk_pod, (size_t)P * 4 * 8);
  c->has_nrt_pods = b->nrt != nullptr;
  if (b->nrt && P > 0) {
    const b200s_nrt_pods* q = b->nrt;
    if (!c->has_nrt) return c->set_err(B200S_ERR_STATE, "pods_upload: NRT pods without NRT snapshot columns");
    if (!q->qos || !q->flags || !q->n_init || !q->n_app || !q->cont_kind || !q->req_mask || !q->req)
      return c->set_err(B200S_ERR_INVALID, "pods_upload: null NRT pod column");
    const int C = B200S_NRT_MAX_CONT, R = c->nrt_R;
    for (int p = 0; p < P; ++p)  // the kernels index the container slots with these counts
      if ((int)q->n_init[p] + (int)q->n_app[p] > C)
        return c->set_err(B200S_ERR_INVALID, "pods_upload: more than 8 containers (flag the pod UNSUPPORTED and zero its counts)");
    up.add(&c->nrt_pod_qos, q->qos, (size_t)P);
    up.add(&c->nrt_pod_flags, q->flags, (size_t)P);
    up.add(&c->nrt_pod_ninit, q->n_init, (size_t)P);
    up.add(&c->nrt_pod_napp, q->n_app, (size_t)P);
    up.add(&c->nrt_pod_kind, q->cont_kind, (size_t)P * C);
    up.add(&c->nrt_pod_req_mask, q->req_mask, (size_t)P * (C + 1));
    up.add(&c->nrt_pod_req, q->req, (size_t)P * (C + 1) * R * 8);
    nrt2_on_pods(c, q, P);  // distinct request vectors of the batch (host dictionary of the batched path)
  } else {
    nrt2_on_pods(c, nullptr, 0);
  }
  c->has_netoh_pods = b->netoh != nullptr;
  if (b->netoh && P > 0) {
    const b200s_netoh_pods* q = b->netoh;
    if (!q->score_equally || !q->dep_offset) return c->set_err(B200S_ERR_INVALID, "pods_upload: null NetworkOverhead column");
    int total = q->dep_offset[P];
    if (total < 0 || (total > 0 && !q->deps)) return c->set_err(B200S_ERR_INVALID, "pods_upload: bad dependency CSR");
    if (!c->has_netoh) return c->set_err(B200S_ERR_STATE, "pods_upload: NetworkOverhead pods without NetworkOverhead snapshot columns");
    for (int p = 0; p < P; ++p)
      if (q->dep_offset[p + 1] < q->dep_offset[p]) return c->set_err(B200S_ERR_INVALID, "pods_upload: dependency offsets must not decrease");
    for (int i = 0; i < total; ++i)  // host labels index the cost tables on the device
      if (q->deps[i].host_region >= c->netoh_K || q->deps[i].host_zone >= c->netoh_K)
        return c->set_err(B200S_ERR_INVALID, "pods_upload: dependency host label id outside the name dictionary");
    up.add(&c->netoh_equal, q->score_equally, (size_t)P);
    up.add(&c->netoh_dep_off, q->dep_offset, (size_t)(P + 1) * 4);
    if (total > 0)
      up.add(&c->netoh_deps, q->deps, (size_t)total * sizeof(b200s_netoh_dep));
    else
      B200S_CUDA_TRY(c, c->netoh_deps.ensure(sizeof(b200s_netoh_dep)));
    c->netoh_total_deps = total;
    c->netoh_max_deps = 0;
    for (int p = 0; p < P; ++p) c->netoh_max_deps = std::max(c->netoh_max_deps, q->dep_offset[p + 1] - q->dep_offset[p]);
  }
  B200S_TRY(up.flush(c));
  B200S_CUDA_TRY(c, cudaEventRecord(c->ev_inputs, c->stream));  // the batch's columns are queued up to here
  // Inputs may be pinned (truly async copies): the caller may reuse them after we return.
  if (!c->defer_sync && !c->async_upload) B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for (auto& o : c->out) o.valid = false;
  c->total_valid = c->topk_valid = c->feas_valid = false;
  c->pods_valid = true;
  return B200S_OK;
}

int b200s_pods_upload(b200s_ctx* c, const b200s_pod_batch* b) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  return pods_upload_locked(c, b);
}

// ---------------------------------------------------------------- eval
static int eval_locked(b200s_ctx* c, b200s_plugin plugin, b200s_out_dtype dtype) {
  if (!c->snap_valid) return c->set_err(B200S_ERR_STATE, "eval: no committed snapshot");
  if (!c->pods_valid) return c->set_err(B200S_ERR_STATE, "eval: no pod batch uploaded");
  if (dtype != B200S_OUT_I64 && dtype != B200S_OUT_U8) return c->set_err(B200S_ERR_INVALID, "eval: bad dtype");
  switch (plugin) {
    case B200S_PLUGIN_ALLOCATABLE: return alloc_eval(c, dtype);
    case B200S_PLUGIN_TLP: return tlp_eval(c, dtype);
    case B200S_PLUGIN_LVRB: return lvrb_eval(c, dtype);
    case B200S_PLUGIN_NRT: return nrt_eval(c, dtype);
    case B200S_PLUGIN_NETWORK_OVERHEAD: return netoh_eval(c, dtype);
    case B200S_PLUGIN_PEAKS: return peaks_eval(c, dtype);
    case B200S_PLUGIN_LOW_RISK: return lowrisk_eval(c, dtype);
    default: return c->set_err(B200S_ERR_INVALID, "eval: unknown plugin");
  }
}

int b200s_eval(b200s_ctx* c, b200s_plugin plugin, b200s_out_dtype dtype) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  return eval_locked(c, plugin, dtype);
}

static int fetch(b200s_ctx* c, const DevBuf& src, void* out, size_t bytes, size_t want, const char* what) {
  if (!out) return c->set_err(B200S_ERR_INVALID, std::string(what) + ": null output");
  if (bytes < want) return c->set_err(B200S_ERR_INVALID, std::string(what) + ": output buffer too small");
  if (want == 0) return B200S_OK;
  if (want <= (size_t)4096 && !c->defer_sync) {
    // small results (a cycle's winners): a device-to-host copy into PAGEABLE memory takes the driver's slow staged
    // path (~10 us); bounce through the engine's pinned page instead
    if (!c->small_bounce) B200S_CUDA_TRY(c, cudaHostAlloc(&c->small_bounce, 4096, cudaHostAllocMapped));
    B200S_CUDA_TRY(c, cudaMemcpyAsync(c->small_bounce, src.p, want, cudaMemcpyDeviceToHost, c->stream));
    B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    memcpy(out, c->small_bounce, want);
    return B200S_OK;
  }
  B200S_CUDA_TRY(c, cudaMemcpyAsync(out, src.p, want, cudaMemcpyDeviceToHost, c->stream));
  if (!c->defer_sync) B200S_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return B200S_OK;
}

static int fetch_scores_locked(b200s_ctx* c, b200s_plugin plugin, void* out, size_t bytes) {
  if ((int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT) return c->set_err(B200S_ERR_INVALID, "fetch: unknown plugin");
  PluginOut& o = c->out[plugin];
  if (!o.valid) return c->set_err(B200S_ERR_STATE, "fetch_scores: plugin not evaluated");
  size_t want = (size_t)o.P * c->Npad * (o.dtype == B200S_OUT_I64 ? 8 : 1);
  return fetch(c, o.scores, out, bytes, want, "fetch_scores");
}

int b200s_fetch_scores(b200s_ctx* c, b200s_plugin plugin, void* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  return fetch_scores_locked(c, plugin, out, bytes);
}

static int fetch_feasible_locked(b200s_ctx* c, b200s_plugin plugin, uint64_t* out, size_t bytes) {
  if ((int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT) return c->set_err(B200S_ERR_INVALID, "fetch: unknown plugin");
  PluginOut& o = c->out[plugin];
  if (!o.valid || !o.has_feas) return c->set_err(B200S_ERR_STATE, "fetch_feasible: no feasibility words for plugin");
  return fetch(c, o.feas, out, bytes, (size_t)o.P * (c->Npad / 64) * 8, "fetch_feasible");
}

int b200s_fetch_feasible(b200s_ctx* c, b200s_plugin plugin, uint64_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  return fetch_feasible_locked(c, plugin, out, bytes);
}

static int fetch_reasons_locked(b200s_ctx* c, b200s_plugin plugin, uint8_t* out, size_t bytes) {
  if ((int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT) return c->set_err(B200S_ERR_INVALID, "fetch: unknown plugin");
  PluginOut& o = c->out[plugin];
  if (!o.valid || !o.has_reasons) return c->set_err(B200S_ERR_STATE, "fetch_reasons: no reason codes for plugin");
  return fetch(c, o.reasons, out, bytes, (size_t)o.P * c->Npad, "fetch_reasons");
}

int b200s_fetch_reasons(b200s_ctx* c, b200s_plugin plugin, uint8_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  return fetch_reasons_locked(c, plugin, out, bytes);
}

int b200s_fetch_network_overhead_raw(b200s_ctx* c, int64_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->out[B200S_PLUGIN_NETWORK_OVERHEAD].valid || c->netoh_raw_P != c->P)
    return c->set_err(B200S_ERR_STATE, "fetch_network_overhead_raw: plugin not evaluated");
  return fetch(c, c->raw_scores, out, bytes, (size_t)c->P * c->Npad * 8, "fetch_network_overhead_raw");
}

int b200s_fetch_network_overhead_counts(b200s_ctx* c, uint32_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->out[B200S_PLUGIN_NETWORK_OVERHEAD].valid || c->netoh_raw_P != c->P || !c->netoh_want_counts)
    return c->set_err(B200S_ERR_STATE, "fetch_network_overhead_counts: counts not enabled or plugin not evaluated");
  return fetch(c, c->netoh_counts, out, bytes, (size_t)c->P * c->Npad * 4, "fetch_network_overhead_counts");
}

void* b200s_device_scores(b200s_ctx* c, b200s_plugin plugin) {
  if (!c || (int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT || !c->out[plugin].valid) return nullptr;
  return c->out[plugin].scores.p;
}
uint64_t* b200s_device_feasible(b200s_ctx* c, b200s_plugin plugin) {
  if (!c || (int)plugin < 0 || plugin >= B200S_PLUGIN_COUNT || !c->out[plugin].valid || !c->out[plugin].has_feas)
    return nullptr;
  return c->out[plugin].feas.as<uint64_t>();
}

int b200s_eval_combined(b200s_ctx* c, uint32_t plugin_mask, const int64_t* weights, int32_t k, int write_total) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->snap_valid) return c->set_err(B200S_ERR_STATE, "eval_combined: no committed snapshot");
  if (!c->pods_valid) return c->set_err(B200S_ERR_STATE, "eval_combined: no pod batch uploaded");
  return combined_eval(c, plugin_mask, weights, k, write_total);
}

int b200s_fetch_topk(b200s_ctx* c, b200s_topk_entry* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->topk_valid) return c->set_err(B200S_ERR_STATE, "fetch_topk: eval_combined not run");
  return fetch(c, c->topk_final, out, bytes, (size_t)c->P * c->topk_k * sizeof(b200s_topk_entry), "fetch_topk");
}

int b200s_fetch_total(b200s_ctx* c, int64_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->total_valid) return c->set_err(B200S_ERR_STATE, "fetch_total: total matrix not written");
  return fetch(c, c->total, out, bytes, (size_t)c->P * c->Npad * 8, "fetch_total");
}

int b200s_fetch_total_feasible(b200s_ctx* c, uint64_t* out, size_t bytes) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  if (!c->topk_valid && !c->feas_valid) return c->set_err(B200S_ERR_STATE, "fetch_total_feasible: eval_combined not run");
  return fetch(c, c->total_feas, out, bytes, (size_t)c->P * (c->Npad / 64) * 8, "fetch_total_feasible");
}

int b200s_debug_div_check(b200s_ctx* c, const double* x, const double* d, int32_t n, uint64_t* mismatches) {
  if (!c || !x || !d || !mismatches || n < 0) return B200S_ERR_INVALID;
  Guard g(c);
  return debug_div_check(c, x, d, n, mismatches);
}

// Large batches of the score-only plugins travel in pod chunks: per chunk the inputs go in on the main stream, the
// kernels run, and the scores leave on a second stream -- so chunk i's D2H overlaps chunk i+1's H2D (the two PCIe
// directions) instead of the whole mask going in before the first score byte comes out.  Pods are independent
// (every normalisation is per pod), so chunking cannot change a result.
static int score_batch_chunked(b200s_ctx* c, b200s_plugin plugin, const b200s_pod_batch* batch, b200s_out_dtype dtype,
                               void* scores_out, int chunks) {
  if (!c->d2h_stream) {
    B200S_CUDA_TRY(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_kernels, cudaEventDisableTiming));
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_d2h, cudaEventDisableTiming));
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_h2d, cudaEventDisableTiming));
  }
  const int P = batch->n_pods;
  const size_t words = (size_t)(c->Npad / 64), esz = dtype == B200S_OUT_I64 ? 8 : 1;
  const int step = (P + chunks - 1) / chunks;
  int rc = B200S_OK;
  for (int i0 = 0, k = 0; i0 < P && rc == B200S_OK; i0 += step, ++k) {
    b200s_pod_batch sub = *batch;
    sub.n_pods = std::min(step, P - i0);
    if (batch->feasible) sub.feasible = batch->feasible + (size_t)i0 * words;
    if (batch->tlp_pod_cpu_milli) sub.tlp_pod_cpu_milli = batch->tlp_pod_cpu_milli + i0;
    if (batch->lvrb_req_cpu_milli) sub.lvrb_req_cpu_milli = batch->lvrb_req_cpu_milli + i0;
    if (batch->lvrb_req_mem_bytes) sub.lvrb_req_mem_bytes = batch->lvrb_req_mem_bytes + i0;
    if (batch->peaks_pod_cpu_milli) sub.peaks_pod_cpu_milli = batch->peaks_pod_cpu_milli + i0;
    sub.nrt = nullptr;
    sub.netoh = nullptr;
    sub.low_risk_pod = nullptr;  // [4][P] layout: not sliceable by pointer offset
    // the pinned staging block of the small pod columns is re-used: the previous chunk's copy must have left it
    if (k > 0) B200S_CUDA_TRY(c, cudaEventSynchronize(c->ev_h2d));
    rc = pods_upload_locked(c, &sub);
    if (rc != B200S_OK) break;
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_h2d, c->stream));
    if (k > 0) B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_d2h, 0));  // the score matrix is re-used too
    rc = eval_locked(c, plugin, dtype);
    if (rc != B200S_OK) break;
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_kernels, c->stream));
    B200S_CUDA_TRY(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_kernels, 0));
    const size_t bytes = (size_t)sub.n_pods * c->Npad * esz;
    B200S_CUDA_TRY(c, cudaMemcpyAsync(static_cast<char*>(scores_out) + (size_t)i0 * c->Npad * esz,
                                      c->out[plugin].scores.p, bytes, cudaMemcpyDeviceToHost, c->d2h_stream));
    B200S_CUDA_TRY(c, cudaEventRecord(c->ev_d2h, c->d2h_stream));
  }
  return rc;
}

int b200s_config_async_upload(b200s_ctx* c, int on) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->async_upload = on != 0;
  return B200S_OK;
}

int b200s_config_fused_cycle(b200s_ctx* c, int on) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  c->fused_cycle = on == 2 ? 2 : (on != 0 ? 1 : 0);
  return B200S_OK;
}

int b200s_schedule_batch(b200s_ctx* c, const b200s_pod_batch* batch, uint32_t plugin_mask, const int64_t* weights, int32_t k,
                         b200s_topk_entry* topk_out) {
  if (!c || !topk_out) return B200S_ERR_INVALID;
  Guard g(c);
  struct Defer {
    b200s_ctx* c;
    explicit Defer(b200s_ctx* x) : c(x) { c->defer_sync = true; }
    ~Defer() { c->defer_sync = false; }
  } defer(c);
  // a handful of pods on one GPU: copy of the pod columns + both launches of the cycle as ONE graph launch (cycle.cu)
  const bool try_graph = c->fused_cycle == 1 && !c->profiling && weights && batch && batch->n_pods >= 1 && batch->n_pods <= 4 &&
                         k >= 1 && (size_t)batch->n_pods * k * sizeof(b200s_topk_entry) <= 4096;
  c->hold_upload = try_graph;
  int rc = pods_upload_locked(c, batch);
  c->hold_upload = false;
  if (rc == B200S_OK && c->held_bytes > 0) {
    if (!c->small_bounce && cudaHostAlloc(&c->small_bounce, 4096, cudaHostAllocMapped) != cudaSuccess) c->small_bounce = nullptr;
    if (c->small_bounce && cycle_applies(c, plugin_mask, k, 0)) {
      const size_t bytes = (size_t)c->P * k * sizeof(b200s_topk_entry);
      c->mask_override = nullptr;
      for (auto& o : c->out) o.valid = false;
      rc = cycle_graph_run(c, plugin_mask, weights, k, static_cast<b200s_topk_entry*>(c->small_bounce));
      cudaError_t e = cudaStreamSynchronize(c->stream);
      c->held_bytes = 0;
      if (rc != B200S_OK) return rc;
      if (e != cudaSuccess) return c->set_err(B200S_ERR_CUDA, std::string("schedule_batch: ") + cudaGetErrorString(e));
      memcpy(topk_out, c->small_bounce, bytes);
      return B200S_OK;
    }
    // not a fused-cycle shape after all: move the staged columns now and take the general path
    cudaError_t e = cudaMemcpyAsync(c->held_dst, c->held_src, c->held_bytes, cudaMemcpyHostToDevice, c->stream);
    c->held_bytes = 0;
    if (e != cudaSuccess) rc = c->set_err(B200S_ERR_CUDA, "schedule_batch: copy of the pod columns failed");
  }
  if (rc == B200S_OK) rc = combined_eval(c, plugin_mask, weights, k, 0);
  const size_t want = rc == B200S_OK ? (size_t)c->P * k * sizeof(b200s_topk_entry) : 0;
  void* dst = topk_out;
  if (want > 0 && want <= 4096) {  // pinned bounce page: a copy into pageable memory takes the driver's slow path
    if (!c->small_bounce && cudaHostAlloc(&c->small_bounce, 4096, cudaHostAllocMapped) != cudaSuccess) c->small_bounce = nullptr;
    if (c->small_bounce) dst = c->small_bounce;
  }
  if (want > 0 && cudaMemcpyAsync(dst, c->topk_final.p, want, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
    rc = c->set_err(B200S_ERR_CUDA, "schedule_batch: copy of the winners failed");
  cudaError_t e = cudaStreamSynchronize(c->stream);  // also on the error path: inputs must not be in flight on return
  if (rc != B200S_OK) return rc;
  if (e != cudaSuccess) return c->set_err(B200S_ERR_CUDA, std::string("schedule_batch: ") + cudaGetErrorString(e));
  if (dst != topk_out) memcpy(topk_out, dst, want);
  return B200S_OK;
}

int b200s_schedule_sequence(b200s_ctx* c, const b200s_pod_batch* batch, uint32_t plugin_mask, const int64_t* weights,
                            b200s_topk_entry* winners_out) {
  if (!c || !winners_out || !weights) return B200S_ERR_INVALID;
  Guard g(c);
  struct Defer {
    b200s_ctx* c;
    explicit Defer(b200s_ctx* x) : c(x) { c->defer_sync = true; }
    ~Defer() { c->defer_sync = false; }
  } defer(c);
  int rc = pods_upload_locked(c, batch);
  if (rc == B200S_OK && !cycle_applies(c, plugin_mask, 1, 0, true))
    rc = c->set_err(B200S_ERR_UNSUPPORTED, "schedule_sequence: needs the fused cycle (single GPU, the five plugins, <= 4 zones x <= 4 "
                                           "resource slots, not LeastNUMANodes)");
  const size_t want = rc == B200S_OK ? (size_t)c->P * sizeof(b200s_topk_entry) : 0;
  if (rc == B200S_OK && want > 0) {
    if (cudaSuccess != c->topk_final.ensure(want)) rc = c->set_err(B200S_ERR_NOMEM, "schedule_sequence: winners buffer");
    if (rc == B200S_OK) rc = cycle_sequence(c, plugin_mask, weights, c->topk_final.as<b200s_topk_entry>());
    if (rc == B200S_OK && batch->nrt && (plugin_mask & (1u << B200S_PLUGIN_NRT)) && c->nrt_R <= 4) {
      // the deductions keep the columns multiples of gcd(snapshot unit, request unit): tell the batched path's bookkeeping
      const int R = c->nrt_R, Cn = B200S_NRT_MAX_CONT;
      std::vector<int64_t> ded((size_t)R * c->P);
      for (int p = 0; p < c->P; ++p)
        for (int r = 0; r < R; ++r) ded[(size_t)r * c->P + p] = batch->nrt->req[((size_t)p * (Cn + 1) + Cn) * R + r];
      nrt2_on_deduct(c, c->P, ded.data());
    }
    // the columns changed under the derived data (batched NRT node columns): a new snapshot serial refreshes them
    c->snap_serial++;
    for (auto& o : c->out) o.valid = false;
    c->topk_valid = c->total_valid = c->feas_valid = false;
    if (rc == B200S_OK && cudaMemcpyAsync(winners_out, c->topk_final.p, want, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = c->set_err(B200S_ERR_CUDA, "schedule_sequence: copy of the winners failed");
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (rc != B200S_OK) return rc;
  if (e != cudaSuccess) return c->set_err(B200S_ERR_CUDA, std::string("schedule_sequence: ") + cudaGetErrorString(e));
  return B200S_OK;
}

int b200s_score_batch(b200s_ctx* c, b200s_plugin plugin, const b200s_pod_batch* batch, b200s_out_dtype dtype,
                      void* scores_out, uint64_t* feasible_out, uint8_t* reasons_out) {
  if (!c) return B200S_ERR_INVALID;
  Guard g(c);
  // one synchronisation at the very end: H2D copies, kernels and D2H copies are queued back to back
  struct Defer {
    b200s_ctx* c;
    explicit Defer(b200s_ctx* x) : c(x) { c->defer_sync = true; }
    ~Defer() { c->defer_sync = false; }
  } defer(c);
  const bool score_only = plugin == B200S_PLUGIN_ALLOCATABLE || plugin == B200S_PLUGIN_TLP ||
                          plugin == B200S_PLUGIN_LVRB || plugin == B200S_PLUGIN_PEAKS;
  const size_t esz = dtype == B200S_OUT_I64 ? 8 : 1;
  // The chunk decision must be the same on every rank of a sharded job (each chunk of a normalising plugin is one
  // min/max all-reduce: ranks that disagree on chunked-or-not or on the chunk count would mismatch collectives), and
  // shards differ in size by up to one 128-node block -- so with a communicator it is taken from rank-invariant
  // values only: P, dtype and the nominal shard size ceil(Nglobal / world) rounded to the node alignment.
  const int world = comm_world(c);
  const size_t npad_nominal =
      world > 1 ? (size_t)round_up((std::max(c->Nglobal, 1) + world - 1) / world, B200S_NODE_ALIGN) : (size_t)c->Npad;
  const size_t out_bytes = batch && batch->n_pods > 0 ? (size_t)batch->n_pods * npad_nominal * esz : 0;
  constexpr size_t kChunkBytes = (size_t)48 << 20;
  if (score_only && batch && scores_out && !feasible_out && !reasons_out && c->snap_valid &&
      (dtype == B200S_OUT_I64 || dtype == B200S_OUT_U8) && out_bytes >= 2 * kChunkBytes && batch->n_pods >= 64) {
    const int chunks = (int)std::min<size_t>(std::min<size_t>(16, (size_t)batch->n_pods / 32), out_bytes / kChunkBytes);
    int rc = score_batch_chunked(c, plugin, batch, dtype, scores_out, std::max(chunks, 2));
    cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->d2h_stream);
    // the engine-resident matrices hold the last chunk only: nothing may be fetched from them afterwards
    c->pods_valid = false;
    for (auto& o : c->out) o.valid = false;
    c->total_valid = c->topk_valid = c->feas_valid = false;
    if (rc != B200S_OK) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess)
      return c->set_err(B200S_ERR_CUDA, std::string("score_batch: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return B200S_OK;
  }
  int rc = pods_upload_locked(c, batch);
  if (rc == B200S_OK) rc = eval_locked(c, plugin, dtype);
  size_t elems = (size_t)c->P * c->Npad;
  if (rc == B200S_OK) rc = fetch_scores_locked(c, plugin, scores_out, elems * (dtype == B200S_OUT_I64 ? 8 : 1));
  if (rc == B200S_OK && feasible_out && c->out[plugin].has_feas)
    rc = fetch_feasible_locked(c, plugin, feasible_out, (size_t)c->P * (c->Npad / 64) * 8);
  if (rc == B200S_OK && reasons_out && c->out[plugin].has_reasons) rc = fetch_reasons_locked(c, plugin, reasons_out, elems);
  cudaError_t e = cudaStreamSynchronize(c->stream);  // also on the error path: inputs must not be in flight on return
  if (rc != B200S_OK) return rc;
  if (e != cudaSuccess) return c->set_err(B200S_ERR_CUDA, std::string("score_batch: ") + cudaGetErrorString(e));
  return B200S_OK;
}

}  // extern "C"

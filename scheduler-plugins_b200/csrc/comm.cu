// Node-axis sharding across GPUs: one ctx per rank, NCCL over NVLink5/NVSwitch for the two tiny
// per-pod exchanges:
//   * min/max all-reduce of the per-pod NormalizeScore bounds (int64[P] each), only for the two
//     normalising plugins (allocatable.go:145-155, networkoverhead.go:421-435);
//   * ONE all-gather of the per-pod top-k (score,node) winners at the end.
// NCCL is resolved with dlopen so that a single-GPU scheduler does not need libnccl at all;
// the minimal prototypes below are NCCL's stable C ABI (nccl.h, 2.x).
//
// Both exchanges are KBs: an NCCL collective costs ~25 us of launch + protocol latency when it runs alone and ~100 us
// when its CTAs have to squeeze in beside a P x N kernel.  Once the ranks have exchanged the CUDA IPC handles of
// their SYMMETRIC buffers (b200s_comm_peer_export / _import; one process per GPU), the exchanges go over NVLink peer
// memory instead: p2p_gather_kernel stores this rank's block straight into slot [rank] of every peer's buffer
// (remote stores through NVSwitch), raises a sequence flag on every peer, and waits for the peers' flags -- one
// small kernel on the compute stream, no proxy thread, no extra stream; the fold (min/max or top-k) reads the local
// buffer.  Payloads above the slot capacity, or ranks without peer mappings, keep NCCL.
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "engine.h"

namespace b200s {

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Api {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};

Api* api() {
  static Api a;
  static bool tried = false;
  if (tried) return &a;
  tried = true;
  const char* env = getenv("B200S_NCCL_LIB");
  const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.lib) break;
  }
  if (!a.lib) {
    a.why = std::string("cannot dlopen libnccl.so.2 (set B200S_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
    return &a;
  }
#define LOAD(field, sym)                                                     \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));          \
  if (!a.field) {                                                            \
    a.why = std::string("libnccl lacks symbol ") + sym;                      \
    a.lib = nullptr;                                                         \
    return &a;                                                               \
  }
  LOAD(GetUniqueId, "ncclGetUniqueId")
  LOAD(CommInitRank, "ncclCommInitRank")
  LOAD(CommDestroy, "ncclCommDestroy")
  LOAD(AllReduce, "ncclAllReduce")
  LOAD(AllGather, "ncclAllGather")
  LOAD(GroupStart, "ncclGroupStart")
  LOAD(GroupEnd, "ncclGroupEnd")
  LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
  return &a;
}

}  // namespace

constexpr size_t kPeerSlotBytes = (size_t)4 << 20;  // one rank's block of one exchange
constexpr int kPeerMaxWorld = 16;

// Symmetric buffer of a rank: flags[2][kPeerMaxWorld] (u64 sequence numbers, one per source rank and parity), a
// block counter, then data[2][world][kPeerSlotBytes].  Two parities: rank A can only start exchange s+2 after every
// rank finished WRITING s+1, which each rank does after it finished READING s -- so slot parity s is free again.
struct PeerHeader {
  unsigned long long flags[2][kPeerMaxWorld];
  unsigned int blocks_done;
  unsigned int pad[31];
};
static_assert(sizeof(PeerHeader) % 16 == 0, "data region stays 16-byte aligned");

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory path
  void* sym = nullptr;             // this rank's symmetric buffer (cudaMalloc)
  void* peer[kPeerMaxWorld] = {};  // every rank's buffer mapped into this process (peer[rank] == sym)
  void** peer_dev = nullptr;       // the same table on the device
  bool peers_ready = false;
  unsigned long long seq = 0;
};

#define B200S_NCCL_TRY(ctx, expr)                                                                    \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != 0) return (ctx)->set_err(B200S_ERR_NCCL, std::string(#expr) + ": " + api()->GetErrorString(_r)); \
  } while (0)

int comm_rank(b200s_ctx* c) { return c->comm ? c->comm->rank : 0; }
int comm_world(b200s_ctx* c) { return c->comm ? c->comm->world : 1; }

void comm_destroy(b200s_ctx* c) {
  if (!c->comm) return;
  Comm* cm = c->comm;
  if (cm->peers_ready)
    for (int r = 0; r < cm->world; ++r)
      if (r != cm->rank && cm->peer[r]) cudaIpcCloseMemHandle(cm->peer[r]);
  if (cm->peer_dev) cudaFree(cm->peer_dev);
  if (cm->sym) cudaFree(cm->sym);
  if (cm->comm && api()->lib) api()->CommDestroy(cm->comm);
  delete cm;
  c->comm = nullptr;
}

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All-gather over peer memory: src (bytes, multiple of 16) -> data[parity][rank] of EVERY rank's symmetric buffer,
// then the flag; the last CTA to finish raises the flags and waits for the peers' -- when the kernel has completed,
// this rank's data[parity][0..world) holds every rank's block.
__global__ void __launch_bounds__(512) p2p_gather_kernel(const uint4* __restrict__ src, size_t n16, void* const* peers,
                                                         int rank, int world, int parity, unsigned long long seq,
                                                         size_t slot_bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t data_off = sizeof(PeerHeader) + ((size_t)parity * world + rank) * slot_bytes;
  for (int r = 0; r < world; ++r) {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(peers[r]) + data_off);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  PeerHeader* mine = static_cast<PeerHeader*>(peers[rank]);
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(&mine->blocks_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) mine->blocks_done = 0;
  __threadfence_system();
  if (threadIdx.x < world) {
    st_release_sys(&static_cast<PeerHeader*>(peers[threadIdx.x])->flags[parity][rank], seq);  // my block is in your buffer
    while (ld_acquire_sys(&mine->flags[parity][threadIdx.x]) < seq) {  // rank threadIdx.x's block is in mine
    }
  }
}

// [world][2 * count] gathered {lo | ~hi} blocks -> lo / hi in place (min over ranks; max = ~min(~hi))
__global__ void minmax_fold_kernel(const int64_t* __restrict__ gathered, size_t slot_elems, int world, int count,
                                   int64_t* __restrict__ lo, int64_t* __restrict__ hi) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int64_t l = INT64_MAX, h = INT64_MIN;
  for (int r = 0; r < world; ++r) {
    const int64_t* g = gathered + (size_t)r * slot_elems;
    l = min(l, __ldcg(g + i));
    h = max(h, __ldcg(g + count + i));
  }
  lo[i] = l;
  hi[i] = h;
}

}  // namespace

// Gathers `bytes` (multiple of 16, <= slot capacity) from every rank over peer memory; *gathered = this rank's
// [world][kPeerSlotBytes] region holding the blocks (valid until the exchange after the next).
static int p2p_gather(b200s_ctx* c, cudaStream_t stream, const void* src, size_t bytes, const void** gathered) {
  Comm* cm = c->comm;
  const int parity = (int)(cm->seq & 1);
  cm->seq++;
  const size_t n16 = bytes / 16;
  const int blocks = (int)std::min<size_t>(16, (n16 + 511) / 512 ? (n16 + 511) / 512 : 1);
  p2p_gather_kernel<<<blocks, 512, 0, stream>>>(static_cast<const uint4*>(src), n16, cm->peer_dev, cm->rank, cm->world,
                                                parity, cm->seq, kPeerSlotBytes);
  c->launches++;
  B200S_CUDA_TRY(c, cudaGetLastError());
  *gathered = static_cast<const char*>(cm->sym) + sizeof(PeerHeader) + (size_t)parity * cm->world * kPeerSlotBytes;
  return B200S_OK;
}
bool comm_has_peers(b200s_ctx* c) { return c->comm && c->comm->peers_ready; }

__global__ void complement_kernel(int64_t* v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = ~v[i];
}

// lo and hi are the two halves of ONE buffer (hi == lo + count).  max(hi) = ~min(~hi): bitwise NOT is an
// order-reversing bijection on int64 (no overflow, unlike negation), so a single ncclMin all-reduce over
// 2*count values does both reductions — one NCCL launch latency per step instead of two.
int comm_allreduce_minmax_on(b200s_ctx* c, cudaStream_t stream, int64_t* lo, int64_t* hi, int count) {
  if (!c->comm || c->comm->world == 1 || count == 0) return B200S_OK;
  Api* a = api();
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_ALLREDUCE, stream);
  if (c->comm->peers_ready && hi == lo + count && (size_t)count * 16 <= kPeerSlotBytes) {
    // peer-memory exchange: every rank's [lo | hi] block lands in every rank's buffer, then a local fold
    const void* g = nullptr;
    B200S_TRY(p2p_gather(c, stream, lo, (size_t)count * 16, &g));
    minmax_fold_kernel<<<(count + 255) / 256, 256, 0, stream>>>(static_cast<const int64_t*>(g), kPeerSlotBytes / 8,
                                                                c->comm->world, count, lo, hi);
    c->launches++;
    B200S_CUDA_TRY(c, cudaGetLastError());
    return B200S_OK;
  }
  if (hi != lo + count) {  // not contiguous: two reductions
    B200S_NCCL_TRY(c, a->GroupStart());
    B200S_NCCL_TRY(c, a->AllReduce(lo, lo, (size_t)count, ncclInt64, ncclMin, c->comm->comm, stream));
    B200S_NCCL_TRY(c, a->AllReduce(hi, hi, (size_t)count, ncclInt64, ncclMax, c->comm->comm, stream));
    B200S_NCCL_TRY(c, a->GroupEnd());
    return B200S_OK;
  }
  complement_kernel<<<(count + 255) / 256, 256, 0, stream>>>(hi, count);
  B200S_NCCL_TRY(c, a->AllReduce(lo, lo, (size_t)count * 2, ncclInt64, ncclMin, c->comm->comm, stream));
  complement_kernel<<<(count + 255) / 256, 256, 0, stream>>>(hi, count);
  c->launches += 2;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

int comm_allreduce_minmax(b200s_ctx* c, int64_t* lo, int64_t* hi, int count) {
  return comm_allreduce_minmax_on(c, c->stream, lo, hi, count);
}

int comm_ensure_streams(b200s_ctx* c) {
  if (c->comm_stream) return B200S_OK;
  int lo_pri = 0, hi_pri = 0;  // the collective's few CTAs must not queue behind a P x N kernel's thousands
  cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);
  B200S_CUDA_TRY(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi_pri));
  for (int i = 0; i < 4; ++i) {
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_reduced[i], cudaEventDisableTiming));
  }
  for (int i = 0; i < 2; ++i) {
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_params[i], cudaEventDisableTiming));
    B200S_CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_norm_done[i], cudaEventDisableTiming));
  }
  return B200S_OK;
}

int comm_allgather(b200s_ctx* c, const void* send, void* recv, size_t bytes_per_rank) {
  if (!c->comm || c->comm->world == 1) {
    if (send != recv)
      B200S_CUDA_TRY(c, cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
    return B200S_OK;
  }
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_ALLGATHER);
  if (c->comm->peers_ready && bytes_per_rank % 16 == 0 && bytes_per_rank <= kPeerSlotBytes) {
    const void* g = nullptr;
    B200S_TRY(p2p_gather(c, c->stream, send, bytes_per_rank, &g));
    B200S_CUDA_TRY(c, cudaMemcpy2DAsync(recv, bytes_per_rank, g, kPeerSlotBytes, bytes_per_rank, (size_t)c->comm->world,
                                        cudaMemcpyDeviceToDevice, c->stream));
    return B200S_OK;
  }
  B200S_NCCL_TRY(c, api()->AllGather(send, recv, bytes_per_rank, ncclUint8, c->comm->comm, c->stream));
  return B200S_OK;
}

}  // namespace b200s

using namespace b200s;

extern "C" {

int b200s_comm_unique_id(void* out_id) {
  if (!out_id) return B200S_ERR_INVALID;
  Api* a = api();
  if (!a->lib) return B200S_ERR_NCCL;
  ncclUniqueId id;
  if (a->GetUniqueId(&id) != 0) return B200S_ERR_NCCL;
  static_assert(sizeof(id) == B200S_UNIQUE_ID_BYTES, "ncclUniqueId size");
  memcpy(out_id, &id, sizeof(id));
  return B200S_OK;
}

int b200s_comm_init(b200s_ctx* c, const void* id, int rank, int world) {
  if (!c || !id || world < 1 || rank < 0 || rank >= world) return B200S_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  comm_destroy(c);
  Api* a = api();
  if (!a->lib) return c->set_err(B200S_ERR_NCCL, a->why);
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  Comm* cm = new Comm();
  cm->rank = rank;
  cm->world = world;
  ncclResult_t r = a->CommInitRank(&cm->comm, world, uid, rank);
  if (r != 0) {
    delete cm;
    return c->set_err(B200S_ERR_NCCL, std::string("ncclCommInitRank: ") + a->GetErrorString(r));
  }
  c->comm = cm;
  return B200S_OK;
}

int b200s_comm_peer_export(b200s_ctx* c, void* out_handle) {
  if (!c || !out_handle) return B200S_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  if (!c->comm) return c->set_err(B200S_ERR_STATE, "comm_peer_export: b200s_comm_init first");
  Comm* cm = c->comm;
  if (cm->world > kPeerMaxWorld) return c->set_err(B200S_ERR_UNSUPPORTED, "comm_peer_export: more than 16 ranks");
  if (!cm->sym) {
    const size_t bytes = sizeof(PeerHeader) + 2 * (size_t)cm->world * kPeerSlotBytes;
    B200S_CUDA_TRY(c, cudaMalloc(&cm->sym, bytes));
    B200S_CUDA_TRY(c, cudaMemset(cm->sym, 0, sizeof(PeerHeader)));
    B200S_CUDA_TRY(c, cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  B200S_CUDA_TRY(c, cudaIpcGetMemHandle(&h, cm->sym));
  static_assert(sizeof(h) == B200S_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  memcpy(out_handle, &h, sizeof(h));
  return B200S_OK;
}

int b200s_comm_peer_import(b200s_ctx* c, const void* handles) {
  if (!c || !handles) return B200S_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaSetDevice(c->device);
  if (!c->comm || !c->comm->sym) return c->set_err(B200S_ERR_STATE, "comm_peer_import: b200s_comm_peer_export first");
  Comm* cm = c->comm;
  for (int r = 0; r < cm->world; ++r) {
    if (r == cm->rank) {
      cm->peer[r] = cm->sym;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)r * B200S_PEER_HANDLE_BYTES, sizeof(h));
    B200S_CUDA_TRY(c, cudaIpcOpenMemHandle(&cm->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  B200S_CUDA_TRY(c, cudaMalloc(reinterpret_cast<void**>(&cm->peer_dev), sizeof(void*) * kPeerMaxWorld));
  B200S_CUDA_TRY(c, cudaMemcpy(cm->peer_dev, cm->peer, sizeof(void*) * kPeerMaxWorld, cudaMemcpyHostToDevice));
  cm->peers_ready = true;
  return B200S_OK;
}

int b200s_comm_rank(b200s_ctx* c) { return c ? comm_rank(c) : 0; }
int b200s_comm_world(b200s_ctx* c) { return c ? comm_world(c) : 1; }

}  // extern "C"

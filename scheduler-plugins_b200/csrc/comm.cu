// Node-axis sharding across GPUs: one ctx per rank, NCCL over NVLink5/NVSwitch for the two tiny
// per-pod exchanges:
//   * min/max all-reduce of the per-pod NormalizeScore bounds (int64[P] each), only for the two
//     normalising plugins (allocatable.go:145-155, networkoverhead.go:421-435);
//   * ONE all-gather of the per-pod top-k (score,node) winners at the end.
// NCCL is resolved with dlopen so that a single-GPU scheduler does not need libnccl at all;
// the minimal prototypes below are NCCL's stable C ABI (nccl.h, 2.x).
#include <dlfcn.h>

#include <cstdlib>
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

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
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
  if (c->comm->comm && api()->lib) api()->CommDestroy(c->comm->comm);
  delete c->comm;
  c->comm = nullptr;
}

__global__ void complement_kernel(int64_t* v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = ~v[i];
}

// lo and hi are the two halves of ONE buffer (hi == lo + count).  max(hi) = ~min(~hi): bitwise NOT is an
// order-reversing bijection on int64 (no overflow, unlike negation), so a single ncclMin all-reduce over
// 2*count values does both reductions — one NCCL launch latency per step instead of two.
int comm_allreduce_minmax(b200s_ctx* c, int64_t* lo, int64_t* hi, int count) {
  if (!c->comm || c->comm->world == 1 || count == 0) return B200S_OK;
  Api* a = api();
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_ALLREDUCE);
  if (hi != lo + count) {  // not contiguous: two reductions
    B200S_NCCL_TRY(c, a->GroupStart());
    B200S_NCCL_TRY(c, a->AllReduce(lo, lo, (size_t)count, ncclInt64, ncclMin, c->comm->comm, c->stream));
    B200S_NCCL_TRY(c, a->AllReduce(hi, hi, (size_t)count, ncclInt64, ncclMax, c->comm->comm, c->stream));
    B200S_NCCL_TRY(c, a->GroupEnd());
    return B200S_OK;
  }
  complement_kernel<<<(count + 255) / 256, 256, 0, c->stream>>>(hi, count);
  B200S_NCCL_TRY(c, a->AllReduce(lo, lo, (size_t)count * 2, ncclInt64, ncclMin, c->comm->comm, c->stream));
  complement_kernel<<<(count + 255) / 256, 256, 0, c->stream>>>(hi, count);
  c->launches += 2;
  B200S_CUDA_TRY(c, cudaGetLastError());
  return B200S_OK;
}

int comm_allgather(b200s_ctx* c, const void* send, void* recv, size_t bytes_per_rank) {
  if (!c->comm || c->comm->world == 1) {
    if (send != recv)
      B200S_CUDA_TRY(c, cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
    return B200S_OK;
  }
  KernelTimer kt(c, B200S_PLUGIN_COUNT + B200S_PHASE_ALLGATHER);
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

int b200s_comm_rank(b200s_ctx* c) { return c ? comm_rank(c) : 0; }
int b200s_comm_world(b200s_ctx* c) { return c ? comm_world(c) : 1; }

}  // extern "C"

// Shared device/host helpers of the b200sched engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b200sched.h"

#define B200S_SM_COUNT 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of it

namespace b200s {

// ---- Go integer semantics ---------------------------------------------------
// Go's int64 arithmetic wraps and `/` truncates toward zero.
// All wrapping arithmetic is done in uint64_t; division special-cases the two
// inputs for which C/CUDA are undefined.
__host__ __device__ __forceinline__ int64_t wrap_add(int64_t a, int64_t b) {
  return (int64_t)((uint64_t)a + (uint64_t)b);
}
__host__ __device__ __forceinline__ int64_t wrap_sub(int64_t a, int64_t b) {
  return (int64_t)((uint64_t)a - (uint64_t)b);
}
__host__ __device__ __forceinline__ int64_t wrap_mul(int64_t a, int64_t b) {
  return (int64_t)((uint64_t)a * (uint64_t)b);
}
// Go: x / y for y != 0; MinInt64 / -1 == MinInt64 (no trap).  y == 0 panics in
// Go; callers never reach it (guarded), we return 0.
__host__ __device__ __forceinline__ int64_t go_div(int64_t x, int64_t y) {
  if (y == 0) return 0;
  if (y == -1) return (int64_t)(0ull - (uint64_t)x);
  return x / y;
}

// ---- streaming global stores --------------------------------------------------
// Score matrices are written once and read by nobody on the device: evict-first.
__device__ __forceinline__ void st_stream_v2(int64_t* p, int64_t a, int64_t b) {
  asm volatile("st.global.cs.v2.s64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_stream_u16(uint8_t* p, uint32_t v) {
  asm volatile("st.global.cs.u16 [%0], %1;" ::"l"(p), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void* p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream_u64(void* p, uint64_t v) {
  asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// One thread's NPT consecutive scores -> one vector store.  int64 rows: NPT = 2 (16 B);
// u8 rows: NPT = 4 (4 B) or 8 (8 B).
template <class OutT, int NPT>
struct Store;
template <>
struct Store<int64_t, 2> {
  __device__ static __forceinline__ void put32(int64_t* p, const uint32_t* q) {
    st_stream_v2(p, (int64_t)q[0], (int64_t)q[1]);
  }
  __device__ static __forceinline__ void put64(int64_t* p, const int64_t* q) { st_stream_v2(p, q[0], q[1]); }
};
template <>
struct Store<uint8_t, 2> {
  __device__ static __forceinline__ void put32(uint8_t* p, const uint32_t* q) {
    st_stream_u16(p, (q[0] & 255u) | ((q[1] & 255u) << 8));
  }
  __device__ static __forceinline__ void put64(uint8_t* p, const int64_t* q) {
    uint32_t v[2] = {(uint32_t)q[0], (uint32_t)q[1]};
    put32(p, v);
  }
};
template <>
struct Store<uint8_t, 4> {
  __device__ static __forceinline__ void put32(uint8_t* p, const uint32_t* q) {
    st_stream_u32(p, (q[0] & 255u) | ((q[1] & 255u) << 8) | ((q[2] & 255u) << 16) | (q[3] << 24));
  }
  __device__ static __forceinline__ void put64(uint8_t* p, const int64_t* q) {
    uint32_t v[4] = {(uint32_t)q[0], (uint32_t)q[1], (uint32_t)q[2], (uint32_t)q[3]};
    put32(p, v);
  }
};
template <>
struct Store<uint8_t, 8> {
  __device__ static __forceinline__ void put32(uint8_t* p, const uint32_t* q) {
    uint32_t a = (q[0] & 255u) | ((q[1] & 255u) << 8) | ((q[2] & 255u) << 16) | (q[3] << 24);
    uint32_t b = (q[4] & 255u) | ((q[5] & 255u) << 8) | ((q[6] & 255u) << 16) | (q[7] << 24);
    st_stream_u64(p, (uint64_t)a | ((uint64_t)b << 32));
  }
  __device__ static __forceinline__ void put64(uint8_t* p, const int64_t* q) {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (uint32_t)q[j];
    put32(p, v);
  }
};

}  // namespace b200s

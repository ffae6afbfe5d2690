// Device side of NetworkOverhead for ONE (pod, node) pair: checkMaxNetworkCostRequirements + getAccumulatedCost
// (networkoverhead.go:500-638) over the pod's flattened dependency entries.  Shared by netoh.cu and the fused
// single-cycle kernel (cycle.cu); moved here verbatim from netoh.cu.
#pragma once
#include "engine.h"

namespace b200s {
namespace netdev {

constexpr int64_t MAX_COST = 100;  // :52
constexpr int64_t SAME_ZONE = 1;   // :58

struct Topo {
  const int64_t* zc;
  const int64_t* rc;
  int K;
};

__device__ __forceinline__ bool lookup_zone(const Topo& t, int r, int z, int d, int64_t& c) {
  if (z != 0) {
    c = __ldg(&t.zc[(size_t)z * t.K + d]);
    if (c != B200S_NETOH_MISSING) return true;
    if (r == z) {
      c = __ldg(&t.rc[(size_t)r * t.K + d]);
      if (c != B200S_NETOH_MISSING) return true;
    }
  }
  return false;
}
__device__ __forceinline__ bool lookup_region(const Topo& t, int r, int z, int d, int64_t& c) {
  if (r != 0) {
    if (z == r) {
      c = __ldg(&t.zc[(size_t)z * t.K + d]);
      if (c != B200S_NETOH_MISSING) return true;
    }
    c = __ldg(&t.rc[(size_t)r * t.K + d]);
    if (c != B200S_NETOH_MISSING) return true;
  }
  return false;
}

// One (pod, node): the two reference loops fused (same traversal, :500-573 and :576-638).
__device__ __forceinline__ void eval_node(const Topo& t, int node_global, int r, int z,
                                          const b200s_netoh_dep* __restrict__ deps, int nd, int64_t& sat,
                                          int64_t& viol, int64_t& cost) {
  sat = viol = cost = 0;
  for (int i = 0; i < nd; ++i) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(deps + i));  // 16 B entry, warp-uniform address
    const int host = raw.x;
    const int hr = raw.y & 0xffff, hz = (raw.y >> 16) & 0xffff;
    const int64_t maxc = (int64_t)(((uint64_t)(uint32_t)raw.w << 32) | (uint32_t)raw.z);
    int64_t c;
    if (host == node_global) {
      sat += 1;  // cost += SameHostname (0)
    } else if (hr == 0 && hz == 0) {
      viol += 1;
      cost = wrap_add(cost, MAX_COST);
    } else if (r == hr) {
      if (z == hz) {
        sat += 1;
        cost = wrap_add(cost, SAME_ZONE);
      } else if (lookup_zone(t, r, z, hz, c)) {
        if (c <= maxc) sat += 1; else viol += 1;
        cost = wrap_add(cost, c);
      } else {
        cost = wrap_add(cost, MAX_COST);  // missing: Filter counts neither, Score adds MaxCost
      }
    } else if (lookup_region(t, r, z, hr, c)) {
      if (c <= maxc) sat += 1; else viol += 1;
      cost = wrap_add(cost, c);
    } else {
      cost = wrap_add(cost, MAX_COST);
    }
  }
}

}  // namespace netdev
}  // namespace b200s

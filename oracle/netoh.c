/* Oracle: NetworkOverhead PreFilter node loop + Filter + Score + NormalizeScore.
 * TEST INFRASTRUCTURE — see oracle.h.
 *
 * Flattened inputs (what the host layer produces from the AppGroup / NetworkTopology CRs):
 *   region_id[n], zone_id[n]  ids in ONE name dictionary (0 = empty label): the reference keeps
 *                             region and zone costs in one (origin,destination) map
 *                             (networkoverhead.go:472-474, 491-493), so names share a namespace.
 *   zone_cost[o*K+d], region_cost[o*K+d]  cost lists per topology key, ORC_NETOH_MISSING if absent.
 *   deps[]: one entry per (placed pod, dependency) pair with matching selector, in the
 *           reference's iteration order (scheduledList outer, dependencyList inner).
 */
#include <stdlib.h>

#include "oracle.h"

#define MAX_COST 100     /* networkoverhead.go:52 */
#define SAME_HOSTNAME 0  /* :55 */
#define SAME_ZONE 1      /* :58 */

/* costMap[{origin: zone, destination: d}] for a node labelled (region r, zone z).
 * populateCostMap (:448-497) inserts region-origin entries first (only if r != ""), then
 * zone-origin entries (only if z != ""), the latter overwriting equal keys. */
static int lookup_zone(const orc_netoh_topology* t, int r, int z, int d, int64_t* cost) {
  if (z != 0) {
    int64_t c = t->zone_cost[(size_t)z * t->n_names + d];
    if (c != ORC_NETOH_MISSING) { *cost = c; return 1; }
    if (r == z) { /* same string used as region and zone name: the region list filled this origin */
      c = t->region_cost[(size_t)r * t->n_names + d];
      if (c != ORC_NETOH_MISSING) { *cost = c; return 1; }
    }
  }
  return 0;
}
static int lookup_region(const orc_netoh_topology* t, int r, int z, int d, int64_t* cost) {
  if (r != 0) {
    if (z == r) {
      int64_t c = t->zone_cost[(size_t)z * t->n_names + d];
      if (c != ORC_NETOH_MISSING) { *cost = c; return 1; } /* zone entry overwrote the region entry */
    }
    int64_t c = t->region_cost[(size_t)r * t->n_names + d];
    if (c != ORC_NETOH_MISSING) { *cost = c; return 1; }
  }
  return 0;
}

/* checkMaxNetworkCostRequirements (:500-573) and getAccumulatedCost (:576-638) for one node. */
void orc_netoh_node(const orc_netoh_topology* t, int node_global, int region, int zone, const orc_netoh_dep* deps,
                    int n_deps, int64_t* satisfied, int64_t* violated, int64_t* cost_out) {
  int64_t sat = 0, viol = 0, cost = 0;
  for (int i = 0; i < n_deps; ++i) {
    const orc_netoh_dep* d = &deps[i];
    int64_t c;
    /* ---- checkMaxNetworkCostRequirements ---- */
    if (d->host_node == node_global) { /* :522-525 */
      sat += 1;
    } else if (d->host_region == 0 && d->host_zone == 0) { /* :538-539 */
      viol += 1;
    } else if (region == d->host_region) { /* :540, string equality incl. the empty string */
      if (zone == d->host_zone) { /* :541-542 */
        sat += 1;
      } else if (lookup_zone(t, region, zone, d->host_zone, &c)) { /* :544-555: missing -> neither */
        if (c <= d->max_network_cost) sat += 1; else viol += 1;
      }
    } else if (lookup_region(t, region, zone, d->host_region, &c)) { /* :557-569 */
      if (c <= d->max_network_cost) sat += 1; else viol += 1;
    }
    /* ---- getAccumulatedCost ---- */
    if (d->host_node == node_global) { /* :594-595 */
      cost = orc_wrap_add(cost, SAME_HOSTNAME);
    } else if (d->host_region == 0 && d->host_zone == 0) { /* :607-608 */
      cost = orc_wrap_add(cost, MAX_COST);
    } else if (region == d->host_region) {
      if (zone == d->host_zone) cost = orc_wrap_add(cost, SAME_ZONE); /* :610-611 */
      else if (lookup_zone(t, region, zone, d->host_zone, &c)) cost = orc_wrap_add(cost, c); /* :613-619 */
      else cost = orc_wrap_add(cost, MAX_COST); /* :620-621: missing -> MaxCost */
    } else {
      if (lookup_region(t, region, zone, d->host_region, &c)) cost = orc_wrap_add(cost, c); /* :625-631 */
      else cost = orc_wrap_add(cost, MAX_COST); /* :632-633 */
    }
  }
  *satisfied = sat;
  *violated = viol;
  *cost_out = cost;
}

/* NormalizeScore (:389-418) with getMinMaxScores (:421-435), in place over one pod's list. */
void orc_netoh_normalize(int64_t* scores, int n) {
  int64_t max = INT64_MIN, min = INT64_MAX;
  for (int i = 0; i < n; ++i) {
    if (scores[i] > max) max = scores[i];
    if (scores[i] < min) min = scores[i];
  }
  if (min == 0 && max == 0) return; /* :400-402 */
  for (int i = 0; i < n; ++i) {
    double norm;
    if (max != min) { /* :406-410 */
      norm = 100.0 * (double)orc_wrap_sub(scores[i], min) / (double)orc_wrap_sub(max, min);
    } else { /* :411-413 */
      norm = (double)orc_wrap_sub(scores[i], min);
    }
    /* int64(normCost): Go/amd64 truncation; out of range -> MinInt64 */
    int64_t tr = (norm >= -9223372036854775808.0 && norm < 9223372036854775808.0) ? (int64_t)norm : INT64_MIN;
    scores[i] = orc_wrap_sub(100, tr);
  }
}

/* One scheduling cycle per pod: PreFilter over every node (:243-280), Filter (:326-359) on the
 * upstream-feasible nodes, Score (:362-386) + NormalizeScore on the nodes that passed. */
void orc_netoh_batch(const orc_netoh_topology* t, const uint16_t* region_id, const uint16_t* zone_id, int N,
                     int node_offset, const uint8_t* score_equally, const int32_t* dep_offset,
                     const orc_netoh_dep* deps, int P, const uint64_t* feasible, int words, int64_t* out_scores,
                     uint64_t* out_feasible, uint8_t* out_reasons, int pitch) {
  int64_t* list = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
  int* idx = (int*)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
  int owords = pitch / 64;
  for (int p = 0; p < P; ++p) {
    int m = 0;
    for (int w = 0; w < owords; ++w) out_feasible[(size_t)p * owords + w] = 0;
    for (int n = 0; n < pitch; ++n) {
      out_scores[(size_t)p * pitch + n] = 0;
      if (out_reasons) out_reasons[(size_t)p * pitch + n] = 0;
    }
    const orc_netoh_dep* pd = deps + dep_offset[p];
    int nd = dep_offset[p + 1] - dep_offset[p];
    for (int n = 0; n < N; ++n) {
      int up = !feasible || ((feasible[(size_t)p * words + (n >> 6)] >> (n & 63)) & 1ull);
      int64_t sat = 0, viol = 0, cost = 0;
      if (!score_equally[p]) orc_netoh_node(t, node_offset + n, region_id[n], zone_id[n], pd, nd, &sat, &viol, &cost);
      int pass = score_equally[p] || !(viol > sat); /* :340-357 */
      if (out_reasons) out_reasons[(size_t)p * pitch + n] = !pass ? 7 : (!up ? 8 : 0);
      if (!pass || !up) continue;
      out_feasible[(size_t)p * owords + (n >> 6)] |= 1ull << (n & 63);
      list[m] = score_equally[p] ? 0 : cost; /* :376-384 */
      idx[m++] = n;
    }
    orc_netoh_normalize(list, m);
    for (int i = 0; i < m; ++i) out_scores[(size_t)p * pitch + idx[i]] = list[i];
  }
  free(list);
  free(idx);
}

// NodeResourceTopologyMatch for ONE (pod, node) pair on the host object model, any shape: NUMA ids in any order,
// sparse or beyond the zone count, more than 8 zones / resource names / containers.  This is where reason code 9
// (B200S_REASON_UNSUPPORTED -- "shape outside the dense encoding") goes in the C++ host mirror; in the Go shim the
// embedded original plugin plays this role (go/pkg/noderesourcetopology/topologymatch_b200.go).  It follows the
// reference statement by statement (files cited per function); it is product code and shares nothing with oracle/.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "objects.hpp"

namespace b200host {

struct NUMANode {  // numaresources.go:31-35
  int numa_id = 0;
  ResourceList resources;     // zone Available, milli-units
  std::map<int, int> costs;   // destination NUMA id -> distance
};
using NUMANodeList = std::vector<NUMANode>;

NUMANodeList CreateNUMANodeList(const NodeResourceTopology& nrt);  // pluginhelpers.go:105-161

// TopologyMatch.Filter after the freshness / nil-NRT gates (filter.go:203-224): policy, scope, handlers
Status ScalarFilter(const Pod& pod, const NodeInfo& node, const NodeResourceTopology& nrt);
// TopologyMatch.Score after the QoS / freshness / nil-NRT gates (score.go:88-101); strategy = B200S_NRT_* id,
// weights by resource name (values < 1 mean 1, score.go:49-60).  Throws std::out_of_range where Go would panic
// (numaScores[NUMAID], numaNodes[bit] with an id beyond the list).
int64_t ScalarScore(const Pod& pod, const NodeResourceTopology& nrt, int strategy, const std::map<std::string, int64_t>& weights);

// numaNodesRequired (least_numa.go:159-174): NUMA ids of the smallest fitting combination (empty = cannot fit) and
// whether its average distance is the minimum for that size
bool OnlyNonNUMAResources(const NUMANodeList& numa_nodes, const ResourceList& resources);  // pluginhelpers.go:163-173
std::vector<int> NumaNodesRequired(QOS qos, const NUMANodeList& numa_nodes, const ResourceList& resources, bool* is_min_distance);

}  // namespace b200host

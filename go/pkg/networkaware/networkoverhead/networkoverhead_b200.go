/*
Copyright The Kubernetes Authors.

Licensed under the Apache License, Version 2.0 (the "License");
you may not use this file except in compliance with the License.
You may obtain a copy of the License at

    http://www.apache.org/licenses/LICENSE-2.0

Unless required by applicable law or agreed to in writing, software
distributed under the License is distributed on an "AS IS" BASIS,
WITHOUT WARRANTIES OR CONDITIONS OF ANY KIND, either express or implied.
See the License for the specific language governing permissions and
limitations under the License.
*/

// NetworkOverhead behind libb200sched (package-internal wrapper; never compiled here: no Go toolchain).
//
// Only the per-node loop of PreFilter (networkoverhead.go:243-280: populateCostMap + checkMaxNetworkCostRequirements +
// getAccumulatedCost for EVERY node, with a sort and two binary searches per node) moves to the engine.  Everything
// before it (AppGroup / NetworkTopology look-ups, dependency and scheduled lists, the scoreEqually early-outs) is the
// original code, and the engine's outputs are written into the SAME PreFilterState maps (satisfiedMap, violatedMap,
// finalCostMap), so Filter (:326-359), Score (:362-386) and NormalizeScore (:389-435) are the embedded plugin's own
// methods, unchanged -- including the exact Filter message.
package networkoverhead

import (
	"context"
	"fmt"
	"sort"
	"unsafe"

	corev1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/labels"
	"k8s.io/apimachinery/pkg/runtime"
	"k8s.io/klog/v2"
	fwk "k8s.io/kube-scheduler/framework"

	agv1alpha1 "github.com/diktyo-io/appgroup-api/pkg/apis/appgroup/v1alpha1"
	ntv1alpha1 "github.com/diktyo-io/networktopology-api/pkg/apis/networktopology/v1alpha1"

	"sigs.k8s.io/scheduler-plugins/pkg/b200sched"
	networkawareutil "sigs.k8s.io/scheduler-plugins/pkg/networkaware/util"
)

const netohMissing = -1 << 63 // B200S_NETOH_MISSING

type NetworkOverheadB200 struct {
	*NetworkOverhead
	eng *b200sched.Engine
}

func NewB200(ctx context.Context, obj runtime.Object, handle fwk.Handle) (fwk.Plugin, error) {
	inner, err := New(ctx, obj, handle)
	if err != nil {
		return nil, err
	}
	no := inner.(*NetworkOverhead)
	eng, err := b200sched.New(0)
	if err != nil {
		klog.FromContext(ctx).Error(err, "b200sched unavailable, NetworkOverhead stays on the Go path")
		return no, nil
	}
	// want_counts: satisfiedMap / violatedMap feed the Filter message; apply_own_filter is irrelevant here because
	// only the raw PreFilterState is fetched (normalisation stays in the embedded NormalizeScore)
	if err := eng.ConfigNetworkOverhead(true, true); err != nil {
		return nil, err
	}
	return &NetworkOverheadB200{NetworkOverhead: no, eng: eng}, nil
}

// name dictionary shared by region and zone label values (the reference keeps both in ONE (origin, destination)-keyed
// map per node, networkoverhead.go:472-493); id 0 = empty label
type dict struct {
	id map[string]uint16
}

func (d *dict) of(s string) uint16 {
	if s == "" {
		return 0
	}
	if v, ok := d.id[s]; ok {
		return v
	}
	v := uint16(len(d.id) + 1)
	d.id[s] = v
	return v
}

func (p *NetworkOverheadB200) PreFilter(ctx context.Context, state fwk.CycleState, pod *corev1.Pod, nodes []fwk.NodeInfo) (*fwk.PreFilterResult, *fwk.Status) {
	preFilterState := &PreFilterState{scoreEqually: true}
	logger := klog.FromContext(klog.NewContext(ctx, p.logger)).WithValues("ExtensionPoint", "PreFilter")
	state.Write(preFilterStateKey, preFilterState)
	// ---- the original's head, verbatim in behaviour (networkoverhead.go:186-232)
	agName := networkawareutil.GetPodAppGroupLabel(pod)
	if len(agName) == 0 {
		return nil, fwk.NewStatus(fwk.Success, "Pod does not belong to an AppGroup, return")
	}
	appGroup := p.findAppGroupNetworkOverhead(ctx, agName)
	networkTopology := p.findNetworkTopologyNetworkOverhead(ctx)
	p.sortNetworkTopologyCosts(networkTopology)
	dependencyList := networkawareutil.GetDependencyList(pod, appGroup)
	if dependencyList == nil {
		return nil, fwk.NewStatus(fwk.Success, "Pod has no dependencies, return")
	}
	selector := labels.Set(map[string]string{agv1alpha1.AppGroupLabel: agName}).AsSelector()
	pods, err := p.podLister.List(selector)
	if err != nil {
		return nil, fwk.NewStatus(fwk.Success, "Error while returning pods from appGroup, return")
	}
	if len(pods) == 0 {
		return nil, fwk.NewStatus(fwk.Success, "No pods yet allocated, return")
	}
	scheduledList := networkawareutil.GetScheduledList(pods)
	if len(scheduledList) == 0 {
		logger.Error(nil, "Scheduled list is empty, return")
		return nil, fwk.NewStatus(fwk.Success, "Scheduled list is empty, return")
	}
	nodeList, err := p.handle.SnapshotSharedLister().NodeInfos().List()
	if err != nil {
		return nil, fwk.NewStatus(fwk.Error, fmt.Sprintf("Error getting the nodelist: %v", err))
	}
	// ---- the node loop (:243-280) as one engine call
	satisfiedMap, violatedMap, finalCostMap, err := p.nodeLoop(nodeList, networkTopology, scheduledList, dependencyList)
	if err != nil {
		logger.V(2).Info("b200sched: falling back to the Go node loop for this cycle", "err", err)
		return p.NetworkOverhead.PreFilter(ctx, state, pod, nodes)
	}
	state.Write(preFilterStateKey, &PreFilterState{
		scoreEqually: false, agName: agName, appGroup: appGroup, networkTopology: networkTopology,
		dependencyList: dependencyList, scheduledList: scheduledList,
		nodeCostMap:  map[string]map[networkawareutil.CostKey]int64{}, // only read inside the loop the engine replaced
		satisfiedMap: satisfiedMap, violatedMap: violatedMap, finalCostMap: finalCostMap,
	})
	return nil, fwk.NewStatus(fwk.Success, "PreFilter State updated")
}

func (p *NetworkOverheadB200) nodeLoop(nodeList []fwk.NodeInfo, nt *ntv1alpha1.NetworkTopology, scheduled networkawareutil.ScheduledList,
	deps []agv1alpha1.DependenciesInfo) (sat, viol, cost map[string]int64, err error) {
	n, npad := len(nodeList), b200sched.NPad(len(nodeList))
	d := &dict{id: map[string]uint16{}}
	var pins []*b200sched.Pinned
	defer func() {
		for _, b := range pins {
			b.Free()
		}
	}()
	pin := func(bytes int) *b200sched.Pinned {
		b, e := b200sched.AllocPinned(bytes)
		if e != nil && err == nil {
			err = e
		}
		pins = append(pins, b)
		return b
	}
	region, zone := pin(2*npad), pin(2*npad)
	if err != nil {
		return
	}
	index := make(map[string]int32, n)
	rg, zn := region.Uint16s(n), zone.Uint16s(n)
	for i, ni := range nodeList {
		index[ni.Node().Name] = int32(i)
		rg[i] = d.of(networkawareutil.GetNodeRegion(ni.Node()))
		zn[i] = d.of(networkawareutil.GetNodeZone(ni.Node()))
	}
	// destinations named in the cost lists join the dictionary too
	type entry struct {
		key          ntv1alpha1.TopologyKey
		origin, dest string
		cost         int64
	}
	var entries []entry
	for _, w := range nt.Spec.Weights {
		if w.Name != p.weightsName {
			continue
		}
		for _, key := range []ntv1alpha1.TopologyKey{ntv1alpha1.NetworkTopologyRegion, ntv1alpha1.NetworkTopologyZone} {
			topologyList := networkawareutil.FindTopologyKey(w.TopologyList, key)
			if p.weightsName != ntv1alpha1.NetworkTopologyNetperfCosts {
				sort.Sort(networkawareutil.ByOrigin(topologyList))
			}
			// only origins that some node carries are ever looked up -- with the reference's own binary search, so
			// an unsorted NetperfCosts list misses exactly the origins it misses there
			seen := map[string]bool{}
			for _, ni := range nodeList {
				o := networkawareutil.GetNodeRegion(ni.Node())
				if key == ntv1alpha1.NetworkTopologyZone {
					o = networkawareutil.GetNodeZone(ni.Node())
				}
				if o == "" || seen[o] {
					continue
				}
				seen[o] = true
				for _, c := range networkawareutil.FindOriginCosts(topologyList, o) {
					entries = append(entries, entry{key, o, c.Destination, c.NetworkCost})
					d.of(c.Destination)
				}
			}
		}
	}
	K := len(d.id) + 1
	zc, rc := pin(8*K*K), pin(8*K*K)
	if err != nil {
		return
	}
	zcm, rcm := zc.Int64s(K*K), rc.Int64s(K*K)
	for i := range zcm {
		zcm[i], rcm[i] = netohMissing, netohMissing
	}
	for _, e := range entries { // later entries overwrite earlier ones, like the map assignment of :472-474, :491-493
		m := rcm
		if e.key == ntv1alpha1.NetworkTopologyZone {
			m = zcm
		}
		m[int(d.of(e.origin))*K+int(d.of(e.dest))] = e.cost
	}
	// ---- the pod: (placed pod, matching dependency) pairs in scheduledList x dependencyList order (:516-519, :589-592)
	type dep struct {
		hostNode                int32
		hostRegion, hostZone    uint16
		maxNetworkCost          int64
	}
	var pairs []dep
	for _, s := range scheduled {
		for _, dp := range deps {
			if s.Selector != dp.Workload.Selector {
				continue
			}
			hi, ok := index[s.Hostname]
			if !ok {
				return nil, nil, nil, fmt.Errorf("pod hostname not found: %s", s.Hostname) // :528-531 -> Error status upstream
			}
			pairs = append(pairs, dep{hi, rg[hi], zn[hi], dp.MaxNetworkCost})
		}
	}
	eq, off, dps := pin(1), pin(8), pin(16*len(pairs)+16)
	if err != nil {
		return
	}
	eq.Bytes(1)[0] = 0
	off.Int32s(2)[0], off.Int32s(2)[1] = 0, int32(len(pairs))
	raw := dps.Bytes(16 * len(pairs))
	for i, pr := range pairs { // b200s_netoh_dep: {int32 host_node; uint16 host_region, host_zone; int64 max_network_cost}
		b := raw[16*i:]
		*(*int32)(ptr(&b[0])) = pr.hostNode
		*(*uint16)(ptr(&b[4])) = pr.hostRegion
		*(*uint16)(ptr(&b[6])) = pr.hostZone
		*(*int64)(ptr(&b[8])) = pr.maxNetworkCost
	}
	scores, rawCost, counts := pin(npad), pin(8*npad), pin(4*npad)
	if err != nil {
		return
	}
	if err = p.eng.SnapshotBegin(0, n, 0, n); err != nil {
		return
	}
	if err = p.eng.SnapshotNetworkOverhead(region, zone, K, zc, rc); err != nil {
		return
	}
	if err = p.eng.SnapshotCommit(); err != nil {
		return
	}
	batch := &b200sched.PodBatch{NPods: 1, NetOH: &b200sched.NetOHPods{ScoreEqually: eq, DepOffset: off, Deps: dps}}
	if err = p.eng.ScoreBatch(b200sched.PluginNetworkOverhead, batch, b200sched.OutU8, scores, nil, nil); err != nil {
		return
	}
	if err = p.eng.FetchNetworkOverheadRaw(rawCost, 8*npad); err != nil {
		return
	}
	if err = p.eng.FetchNetworkOverheadCounts(counts, 4*npad); err != nil {
		return
	}
	sat, viol, cost = make(map[string]int64, n), make(map[string]int64, n), make(map[string]int64, n)
	rcst, cnt := rawCost.Int64s(n), counts.Int32s(n)
	for i, ni := range nodeList {
		name := ni.Node().Name
		cost[name] = rcst[i]
		sat[name] = int64(uint32(cnt[i]) & 0xFFFF)
		viol[name] = int64(uint32(cnt[i]) >> 16)
	}
	return
}

// ptr is the one unsafe cast of this file (packing b200s_netoh_dep records into pinned memory).
func ptr(b *byte) unsafe.Pointer { return unsafe.Pointer(b) }

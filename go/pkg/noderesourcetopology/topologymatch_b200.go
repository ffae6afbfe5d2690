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

// NodeResourceTopologyMatch behind libb200sched (package-internal wrapper; never compiled here: no Go toolchain).
//
// PreFilter flattens the cycle's nodes (NRT objects as the cache hands them out, OverReserve deductions included)
// and the pod into the dense encoding of include/b200sched.h and makes ONE engine call that yields, for every node,
// the Filter verdict (feasibility bit + reason code) and the Score.  Filter / Score are lookups.  Whatever the dense
// encoding cannot express (reason code 9: NUMA ids that are not 0..k-1 in list order, > 8 zones / resources /
// containers) is answered by the embedded original plugin for that (pod, node) pair -- same results, just slower.
// Reserve / Unreserve / PostBind / EventsToRegister are the embedded plugin's own (reserve.go:28-45, postbind.go:28).
package noderesourcetopology

import (
	"context"
	"sort"

	topologyv1alpha2 "github.com/k8stopologyawareschedwg/noderesourcetopology-api/pkg/apis/topology/v1alpha2"
	"github.com/k8stopologyawareschedwg/noderesourcetopology-api/pkg/numanode"
	v1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/api/resource"
	"k8s.io/apimachinery/pkg/runtime"
	"k8s.io/klog/v2"
	fwk "k8s.io/kube-scheduler/framework"
	v1qos "k8s.io/kubernetes/pkg/apis/core/v1/helper/qos"
	kubeletconfig "k8s.io/kubernetes/pkg/kubelet/apis/config"

	apiconfig "sigs.k8s.io/scheduler-plugins/apis/config"
	"sigs.k8s.io/scheduler-plugins/pkg/b200sched"
	"sigs.k8s.io/scheduler-plugins/pkg/noderesourcetopology/nodeconfig"
	"sigs.k8s.io/scheduler-plugins/pkg/noderesourcetopology/resourcerequests"
	"sigs.k8s.io/scheduler-plugins/pkg/util"
)

const (
	nrtCycleKey = "PreFilter" + Name + "B200"
	zMax, rMax, cMax = 8, 8, 8 // B200S_NRT_MAX_ZONES / _RES / _CONT

	nodeHasNRT, nodeFresh, nodeSingleNUMA, nodeScopePod, nodeUnsupported = 1, 2, 4, 8, 16 // B200S_NRT_NODE_*
	resAffine, resHostLevel                                              = 1, 2            // B200S_NRT_RES_*
	podFilterBypass, podUnsupported                                      = 1, 2            // B200S_NRT_POD_*
	contApp, contInit, contSidecar                                       = 0, 1, 2         // B200S_CONT_*
)

var strategyID = map[apiconfig.ScoringStrategyType]int{ // B200S_NRT_* strategies
	apiconfig.MostAllocated: 0, apiconfig.BalancedAllocation: 1, apiconfig.LeastAllocated: 2, apiconfig.LeastNUMANodes: 3,
}

var reasonStatus = map[uint8]*fwk.Status{ // exact messages of filter.go:50-53, 65, 169, 196
	b200sched.ReasonNRTInvalidTopology: fwk.NewStatus(fwk.Unschedulable, "invalid node topology data"),
	b200sched.ReasonNRTAlignPod:        fwk.NewStatus(fwk.Unschedulable, "cannot align pod"),
	b200sched.ReasonNRTAlignContainer:  fwk.NewStatus(fwk.Unschedulable, "cannot align container"),
	b200sched.ReasonNRTAlignInit:       fwk.NewStatus(fwk.Unschedulable, "cannot align init container"),
	b200sched.ReasonNRTAlignSidecar:    fwk.NewStatus(fwk.Unschedulable, "cannot align sidecar container"),
}

type TopologyMatchB200 struct {
	*TopologyMatch
	handle fwk.Handle
	eng    *b200sched.Engine
}

var _ fwk.PreFilterPlugin = &TopologyMatchB200{}

func NewB200(ctx context.Context, args runtime.Object, handle fwk.Handle) (fwk.Plugin, error) {
	inner, err := New(ctx, args, handle)
	if err != nil {
		return nil, err
	}
	tm := inner.(*TopologyMatch)
	eng, err := b200sched.New(0)
	if err != nil {
		klog.FromContext(ctx).Error(err, "b200sched unavailable, NodeResourceTopologyMatch stays on the Go path")
		return tm, nil
	}
	return &TopologyMatchB200{TopologyMatch: tm, handle: handle, eng: eng}, nil
}

func milli(q resource.Quantity, name v1.ResourceName) int64 { // exact milli-units (Appendix B of SURVEY.md)
	if name == v1.ResourceCPU {
		return q.MilliValue()
	}
	return q.Value() * 1000
}

// zones of type Node named node-<id>, id <= 64 (createNUMANodeList, pluginhelpers.go:105-134); dense = ids 0..k-1 in order
func numaZones(nrt *topologyv1alpha2.NodeResourceTopology) (zones []topologyv1alpha2.Zone, dense bool) {
	dense = true
	for _, z := range nrt.Zones {
		if z.Type != "Node" {
			continue
		}
		id, err := numanode.NameToID(z.Name)
		if err != nil || id > maxNUMAId {
			continue
		}
		if id != len(zones) {
			dense = false
		}
		zones = append(zones, z)
	}
	return zones, dense && len(zones) <= zMax
}

type nrtCycle struct {
	b200sched.CycleResult
	pod *v1.Pod
}

func (c *nrtCycle) Clone() fwk.StateData { return c }

func (p *TopologyMatchB200) PreFilter(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, nodes []fwk.NodeInfo) (*fwk.PreFilterResult, *fwk.Status) {
	lh := klog.FromContext(ctx)
	res := &nrtCycle{pod: pod}
	res.Index = make(map[string]int32, len(nodes))
	if err := p.run(ctx, pod, nodes, res); err != nil {
		lh.V(2).Info("b200sched: falling back to the Go path for this cycle", "err", err)
		res.Fallback = true
	}
	cs.Write(nrtCycleKey, res)
	return nil, nil
}

func (p *TopologyMatchB200) PreFilterExtensions() fwk.PreFilterExtensions { return nil }

func (p *TopologyMatchB200) run(ctx context.Context, pod *v1.Pod, nodes []fwk.NodeInfo, out *nrtCycle) error {
	n, npad := len(nodes), b200sched.NPad(len(nodes))
	// ---- resource-slot dictionary of this pod: every requested name, cpu and memory first
	names := []v1.ResourceName{v1.ResourceCPU, v1.ResourceMemory}
	slot := map[v1.ResourceName]int{v1.ResourceCPU: 0, v1.ResourceMemory: 1}
	add := func(rl v1.ResourceList) {
		keys := make([]string, 0, len(rl))
		for k := range rl {
			keys = append(keys, string(k))
		}
		sort.Strings(keys) // deterministic slot order (map iteration is not)
		for _, k := range keys {
			if _, ok := slot[v1.ResourceName(k)]; !ok {
				slot[v1.ResourceName(k)] = len(names)
				names = append(names, v1.ResourceName(k))
			}
		}
	}
	for i := range pod.Spec.InitContainers {
		add(pod.Spec.InitContainers[i].Resources.Requests)
	}
	for i := range pod.Spec.Containers {
		add(pod.Spec.Containers[i].Resources.Requests)
	}
	add(pod.Spec.Overhead)
	R := len(names)
	if R > rMax {
		R = rMax
	}
	// ---- node columns
	type nodeNRT struct {
		zones []topologyv1alpha2.Zone
		dense bool
		nrt   *topologyv1alpha2.NodeResourceTopology
		fresh bool
	}
	per := make([]nodeNRT, n)
	Z := 1
	for i, ni := range nodes {
		out.Index[ni.Node().Name] = int32(i)
		nrt, info := p.nrtCache.GetCachedNRTCopy(ctx, ni.Node().Name, pod) // OverReserve deductions already applied
		per[i] = nodeNRT{nrt: nrt, fresh: info.Fresh, dense: true}
		if nrt != nil {
			per[i].zones, per[i].dense = numaZones(nrt)
			if per[i].dense && len(per[i].zones) > Z {
				Z = len(per[i].zones)
			}
		}
	}
	alloc := func(bytes int) *b200sched.Pinned {
		b, err := b200sched.AllocPinned(bytes)
		if err != nil {
			panic(err) // recovered below
		}
		return b
	}
	var pins []*b200sched.Pinned
	defer func() {
		for _, b := range pins {
			b.Free()
		}
	}()
	pin := func(bytes int) *b200sched.Pinned { b := alloc(bytes); pins = append(pins, b); return b }
	nn := &b200sched.NRTNodes{NZones: Z, NRes: R, ResFlags: pin(R), NodeFlags: pin(npad), MaxNUMA: pin(2 * npad), NZonesNode: pin(npad),
		NodeResMask: pin(npad), ZoneResMask: pin(Z * npad), Avail: pin(8 * Z * R * npad), Cost: pin(4 * Z * Z * npad)}
	rf := nn.ResFlags.Bytes(R)
	for r := 0; r < R; r++ {
		rf[r] = 0
		if isNUMAAffineResource(names[r]) {
			rf[r] |= resAffine
		}
		if isHostLevelResource(names[r]) {
			rf[r] |= resHostLevel
		}
	}
	nf, mx, nz, nrm := nn.NodeFlags.Bytes(n), nn.MaxNUMA.Uint16s(n), nn.NZonesNode.Bytes(n), nn.NodeResMask.Bytes(n)
	zm, av, co := nn.ZoneResMask.Bytes(Z*n), nn.Avail.Int64s(Z*R*n), nn.Cost.Int32s(Z*Z*n)
	for i := range zm {
		zm[i] = 0
	}
	for i := range av {
		av[i] = 0
	}
	for i := range co {
		co[i] = -1
	}
	for i, ni := range nodes {
		var fl uint8
		mx[i], nz[i], nrm[i] = 8, 0, 0
		if per[i].fresh {
			fl |= nodeFresh
		}
		nodeRes := util.ResourceList(ni.GetAllocatable()) // filter.go:97
		for r := 0; r < R; r++ {
			if _, ok := nodeRes[names[r]]; ok {
				nrm[i] |= 1 << r
			}
		}
		if nrt := per[i].nrt; nrt != nil {
			fl |= nodeHasNRT
			conf := nodeconfig.TopologyManagerFromNodeResourceTopology(klog.FromContext(ctx), nrt)
			if conf.Policy == kubeletconfig.SingleNumaNodeTopologyManagerPolicy {
				fl |= nodeSingleNUMA
			}
			if conf.Scope == kubeletconfig.PodTopologyManagerScope {
				fl |= nodeScopePod
			}
			mx[i] = uint16(conf.MaxNUMANodes)
			if !per[i].dense {
				fl |= nodeUnsupported
			} else {
				nz[i] = uint8(len(per[i].zones))
				for z, zone := range per[i].zones {
					for _, ri := range zone.Resources {
						if r, ok := slot[v1.ResourceName(ri.Name)]; ok && r < R {
							zm[z*n+i] |= 1 << r
							av[(z*R+r)*n+i] = milli(ri.Available, v1.ResourceName(ri.Name)) // extractResources: Available
						}
					}
					for _, c := range zone.Costs { // extractCosts (pluginhelpers.go:136-153)
						if id, err := numanode.NameToID(c.Name); err == nil && id < len(per[i].zones) {
							co[(z*Z+id)*n+i] = int32(c.Value)
						}
					}
				}
			}
		}
		nf[i] = fl
	}
	// ---- the pod
	qos := v1qos.GetPodQOS(pod)
	var pflags uint8
	if qos == v1.PodQOSBestEffort && !resourcerequests.IncludeNonNative(pod) {
		pflags |= podFilterBypass // filter.go:180-183
	}
	nInit, nApp := len(pod.Spec.InitContainers), len(pod.Spec.Containers)
	np := &b200sched.NRTPods{QoS: pin(1), Flags: pin(1), NInit: pin(1), NApp: pin(1), ContKind: pin(cMax), ReqMask: pin(cMax + 1),
		Req: pin(8 * (cMax + 1) * R)}
	kind, rmask, req := np.ContKind.Bytes(cMax), np.ReqMask.Bytes(cMax+1), np.Req.Int64s((cMax+1)*R)
	for i := range req {
		req[i] = 0
	}
	for i := range rmask {
		rmask[i] = 0
	}
	put := func(c int, rl v1.ResourceList) {
		for name, q := range rl {
			r := slot[name]
			if r >= R {
				pflags |= podUnsupported
				continue
			}
			rmask[c] |= 1 << r
			req[c*R+r] = milli(q, name)
		}
	}
	if nInit+nApp > cMax {
		pflags |= podUnsupported
		nInit, nApp = 0, 0
	} else {
		for c := range pod.Spec.InitContainers {
			kind[c] = contInit
			if util.IsSidecarInitContainer(&pod.Spec.InitContainers[c]) { // logging.go:68-73
				kind[c] = contSidecar
			}
			put(c, pod.Spec.InitContainers[c].Resources.Requests)
		}
		for c := range pod.Spec.Containers {
			kind[nInit+c] = contApp
			put(nInit+c, pod.Spec.Containers[c].Resources.Requests)
		}
	}
	put(cMax, util.GetPodEffectiveRequest(pod)) // pkg/util/resource.go:51-85
	np.QoS.Bytes(1)[0] = map[v1.PodQOSClass]uint8{v1.PodQOSGuaranteed: 0, v1.PodQOSBurstable: 1, v1.PodQOSBestEffort: 2}[qos]
	np.Flags.Bytes(1)[0], np.NInit.Bytes(1)[0], np.NApp.Bytes(1)[0] = pflags, uint8(nInit), uint8(nApp)
	// ---- engine: snapshot (the NRT cache hands out per-cycle copies, so the columns are per cycle), args, one call
	weights := make([]int64, R)
	for r := 0; r < R; r++ {
		weights[r] = p.resourceToWeightMap.weight(names[r]) // score.go:49-60 (< 1 means 1)
	}
	scores, feas, reasons := pin(npad), pin(npad/8), pin(npad)
	if err := p.eng.SnapshotBegin(0, n, 0, n); err != nil {
		return err
	}
	if err := p.eng.SnapshotNRT(nn); err != nil {
		return err
	}
	if err := p.eng.SnapshotCommit(); err != nil {
		return err
	}
	if err := p.eng.ConfigNRT(strategyID[p.scoreStrategyType], weights); err != nil {
		return err
	}
	if err := p.eng.ScoreBatch(b200sched.PluginNRT, &b200sched.PodBatch{NPods: 1, NRT: np}, b200sched.OutU8, scores, feas, reasons); err != nil {
		return err
	}
	out.Scores = append([]uint8(nil), scores.Bytes(npad)...)
	out.Feasible = append([]uint64(nil), feas.Uint64s(npad/64)...)
	out.Reasons = append([]uint8(nil), reasons.Bytes(npad)...)
	return nil
}

func (p *TopologyMatchB200) cycle(cs fwk.CycleState) *nrtCycle {
	if d, err := cs.Read(nrtCycleKey); err == nil {
		if c, ok := d.(*nrtCycle); ok && !c.Fallback {
			return c
		}
	}
	return nil
}

func (p *TopologyMatchB200) Filter(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, ni fwk.NodeInfo) *fwk.Status {
	if ni.Node() == nil {
		return fwk.NewStatus(fwk.Error, "node not found") // filter.go:177-179
	}
	c := p.cycle(cs)
	if c == nil {
		return p.TopologyMatch.Filter(ctx, cs, pod, ni)
	}
	i, ok := c.Index[ni.Node().Name]
	if !ok || c.Reasons[i] == b200sched.ReasonUnsupported {
		return p.TopologyMatch.Filter(ctx, cs, pod, ni) // outside the dense encoding: the original answers this pair
	}
	switch r := c.Reasons[i]; r {
	case b200sched.ReasonOK:
		return nil
	case b200sched.ReasonNRTAccounting:
		return fwk.NewStatus(fwk.Error, "inconsistent resource accounting") // filter.go:73
	default:
		p.nrtCache.NodeMaybeOverReserved(ni.Node().Name, pod) // filter.go:221-223
		return reasonStatus[r]
	}
}

func (p *TopologyMatchB200) Score(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, ni fwk.NodeInfo) (int64, *fwk.Status) {
	c := p.cycle(cs)
	if c == nil {
		return p.TopologyMatch.Score(ctx, cs, pod, ni)
	}
	i, ok := c.Index[ni.Node().Name]
	if !ok || c.Reasons[i] == b200sched.ReasonUnsupported {
		return p.TopologyMatch.Score(ctx, cs, pod, ni)
	}
	return int64(c.Scores[i]), nil
}

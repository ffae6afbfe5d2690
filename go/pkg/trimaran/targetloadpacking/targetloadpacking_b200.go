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

// TargetLoadPacking behind libb200sched.  Lives in package targetloadpacking (reuses PredictUtilisation and the
// plugin's collector / bind cache).  Never compiled in this repository (no Go toolchain).
package targetloadpacking

import (
	"context"

	"github.com/paypal/load-watcher/pkg/watcher"
	v1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/runtime"
	"k8s.io/klog/v2"
	fwk "k8s.io/kube-scheduler/framework"

	"sigs.k8s.io/scheduler-plugins/pkg/b200sched"
)

const tlpCycleKey = "PreScore" + Name + "B200"

const (
	tlpHasMetrics = 1 // B200S_TLP_HAS_METRICS
	tlpCPUFound   = 2 // B200S_TLP_CPU_FOUND
)

type TargetLoadPackingB200 struct {
	*TargetLoadPacking
	eng                             *b200sched.Engine
	table                           b200sched.NodeTable
	util, cap_, missing, flags, pod *b200sched.Pinned
	row                             *b200sched.Pinned
	metricsEnd                      int64 // Window.End the columns were flattened at
}

var _ fwk.PreScorePlugin = &TargetLoadPackingB200{}

func NewB200(ctx context.Context, obj runtime.Object, handle fwk.Handle) (fwk.Plugin, error) {
	inner, err := New(ctx, obj, handle)
	if err != nil {
		return nil, err
	}
	pl := inner.(*TargetLoadPacking)
	eng, err := b200sched.New(0)
	if err != nil {
		klog.FromContext(ctx).Error(err, "b200sched unavailable, TargetLoadPacking stays on the Go path")
		return pl, nil
	}
	// per-instance target (the reference keeps it in a package global, targetloadpacking.go:49-53)
	if err := eng.ConfigTLP(pl.args.TargetUtilization); err != nil {
		return nil, err
	}
	return &TargetLoadPackingB200{TargetLoadPacking: pl, eng: eng}, nil
}

// one node's row: targetloadpacking.go:112-167 up to (not including) the pod-dependent arithmetic
func (p *TargetLoadPackingB200) flattenRow(logger klog.Logger, ni fwk.NodeInfo) (util float64, capMilli, missingMilli int64, flags uint8) {
	node := ni.Node()
	metrics, all := p.collector.GetNodeMetrics(logger, node.Name)
	capMilli = node.Status.Capacity.Cpu().MilliValue() // Status.Capacity, not Allocatable (:146)
	if metrics == nil {
		return
	}
	flags |= tlpHasMetrics
	for _, m := range metrics { // the LAST Average|Latest cpu metric wins (:131-140)
		if m.Type == watcher.CPU && (m.Operator == watcher.Average || m.Operator == watcher.Latest) {
			util = m.Value
			flags |= tlpCPUFound
		}
	}
	p.eventHandler.RLock()
	for _, info := range p.eventHandler.ScheduledPodsCache[node.Name] { // :151-167
		ts := info.Timestamp.Unix()
		if ts > all.Window.End || ts <= all.Window.End && (all.Window.End-ts) < metricsAgentReportingIntervalSeconds {
			for i := range info.Pod.Spec.Containers {
				missingMilli += PredictUtilisation(&info.Pod.Spec.Containers[i])
			}
			missingMilli += info.Pod.Spec.Overhead.Cpu().MilliValue()
		}
	}
	p.eventHandler.RUnlock()
	return
}

func (p *TargetLoadPackingB200) ensureSnapshot(logger klog.Logger, all []fwk.NodeInfo) error {
	// the columns depend on the node list, on the collector's last fetch and on the bind cache (wall-clock dependent:
	// a bound pod leaves `missing` 60 s after Window.End) -- re-flatten when any of them moved.  The bind cache changes
	// one node at a time: that path is b200s_snapshot_patch_tlp (see the C++ host mirror's PatchSnapshot).
	full, changed := p.table.Diff(all)
	_, allMetrics := p.collector.GetNodeMetrics(logger, "")
	end := int64(0)
	if allMetrics != nil {
		end = allMetrics.Window.End
	}
	if !full && len(changed) == 0 && end == p.metricsEnd {
		return nil
	}
	p.metricsEnd = end
	if full {
		p.table.Reset(all)
		for _, b := range []**b200sched.Pinned{&p.util, &p.cap_, &p.missing, &p.flags, &p.row, &p.pod} {
			if *b != nil {
				(*b).Free()
			}
		}
		var err error
		alloc := func(bytes int) *b200sched.Pinned {
			b, e := b200sched.AllocPinned(bytes)
			if e != nil {
				err = e
			}
			return b
		}
		p.util, p.cap_, p.missing = alloc(8*p.table.NPad), alloc(8*p.table.NPad), alloc(8*p.table.NPad)
		p.flags, p.row, p.pod = alloc(p.table.NPad), alloc(p.table.NPad), alloc(8)
		if err != nil {
			return err
		}
	}
	u, c, m, f := p.util.Float64s(p.table.N), p.cap_.Int64s(p.table.N), p.missing.Int64s(p.table.N), p.flags.Bytes(p.table.N)
	for i, ni := range all {
		u[i], c[i], m[i], f[i] = p.flattenRow(logger, ni)
		p.table.Gen[i] = ni.GetGeneration()
	}
	if err := p.eng.SnapshotBegin(p.table.Epoch, p.table.N, 0, p.table.N); err != nil {
		return err
	}
	if err := p.eng.SnapshotTLP(p.util, p.cap_, p.missing, p.flags); err != nil {
		return err
	}
	return p.eng.SnapshotCommit()
}

func (p *TargetLoadPackingB200) PreScore(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, feasible []fwk.NodeInfo) *fwk.Status {
	logger := klog.FromContext(klog.NewContext(ctx, p.logger)).WithValues("ExtensionPoint", "PreScore")
	all, err := p.handle.SnapshotSharedLister().NodeInfos().List()
	if err == nil {
		err = p.ensureSnapshot(logger, all)
	}
	res := &b200sched.CycleResult{Index: p.table.Index}
	if err == nil {
		var podCPU int64 // :122-129
		for i := range pod.Spec.Containers {
			podCPU += PredictUtilisation(&pod.Spec.Containers[i])
		}
		if pod.Spec.Overhead != nil {
			podCPU += pod.Spec.Overhead.Cpu().MilliValue()
		}
		p.pod.Int64s(1)[0] = podCPU
		err = p.eng.ScoreBatch(b200sched.PluginTLP, &b200sched.PodBatch{NPods: 1, TLPPodCPUMilli: p.pod}, b200sched.OutU8, p.row, nil, nil)
	}
	if err != nil {
		logger.V(2).Info("b200sched: falling back to the Go path for this cycle", "err", err)
		res.Fallback = true
	} else {
		res.Scores = append([]uint8(nil), p.row.Bytes(p.table.NPad)...)
	}
	cs.Write(tlpCycleKey, res)
	return nil
}

func (p *TargetLoadPackingB200) Score(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, ni fwk.NodeInfo) (int64, *fwk.Status) {
	if d, err := cs.Read(tlpCycleKey); err == nil {
		if c, ok := d.(*b200sched.CycleResult); ok {
			if i, ok := c.Lookup(ni.Node().Name); ok {
				return int64(c.Scores[i]), fwk.NewStatus(fwk.Success, "")
			}
		}
	}
	return p.TargetLoadPacking.Score(ctx, cs, pod, ni)
}

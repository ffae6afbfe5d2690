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

// LoadVariationRiskBalancing behind libb200sched (package-internal wrapper; never compiled here: no Go toolchain).
package loadvariationriskbalancing

import (
	"context"

	"github.com/paypal/load-watcher/pkg/watcher"
	v1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/runtime"
	"k8s.io/klog/v2"
	fwk "k8s.io/kube-scheduler/framework"

	"sigs.k8s.io/scheduler-plugins/pkg/b200sched"
	"sigs.k8s.io/scheduler-plugins/pkg/trimaran"
)

const lvrbCycleKey = "PreScore" + Name + "B200"

const (
	lvrbHasMetrics = 1 // B200S_LVRB_HAS_METRICS
	lvrbCPUOK      = 2
	lvrbMemOK      = 4
)

type LoadVariationRiskBalancingB200 struct {
	*LoadVariationRiskBalancing
	eng                                                       *b200sched.Engine
	table                                                     b200sched.NodeTable
	cpuAvg, cpuStd, memAvg, memStd, allocCPU, allocMem, flags *b200sched.Pinned
	reqCPU, reqMem, row                                       *b200sched.Pinned
	metricsEnd                                                int64
}

var _ fwk.PreScorePlugin = &LoadVariationRiskBalancingB200{}

func NewB200(ctx context.Context, obj runtime.Object, handle fwk.Handle) (fwk.Plugin, error) {
	inner, err := New(ctx, obj, handle)
	if err != nil {
		return nil, err
	}
	pl := inner.(*LoadVariationRiskBalancing)
	eng, err := b200sched.New(0)
	if err != nil {
		klog.FromContext(ctx).Error(err, "b200sched unavailable, LoadVariationRiskBalancing stays on the Go path")
		return pl, nil
	}
	if err := eng.ConfigLVRB(pl.args.SafeVarianceMargin, pl.args.SafeVarianceSensitivity); err != nil {
		return nil, err
	}
	return &LoadVariationRiskBalancingB200{LoadVariationRiskBalancing: pl, eng: eng}, nil
}

// GetResourceData (pkg/trimaran/resourcestats.go:89-107): Average sets avg and wins; Std sets the deviation;
// "" / Latest sets avg only while no Average was seen; valid iff any metric of the type exists.
func resourceData(metrics []watcher.Metric, t string) (avg, std float64, ok bool) {
	avgFound := false
	for _, m := range metrics {
		if m.Type != t {
			continue
		}
		ok = true
		switch m.Operator {
		case watcher.Average:
			avg, avgFound = m.Value, true
		case watcher.Std:
			std = m.Value
		case "", watcher.Latest:
			if !avgFound {
				avg = m.Value
			}
		}
	}
	return
}

func (p *LoadVariationRiskBalancingB200) ensureSnapshot(logger klog.Logger, all []fwk.NodeInfo) error {
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
		var err error
		alloc := func(old *b200sched.Pinned, bytes int) *b200sched.Pinned {
			if old != nil {
				old.Free()
			}
			b, e := b200sched.AllocPinned(bytes)
			if e != nil {
				err = e
			}
			return b
		}
		n8 := 8 * p.table.NPad
		p.cpuAvg, p.cpuStd, p.memAvg, p.memStd = alloc(p.cpuAvg, n8), alloc(p.cpuStd, n8), alloc(p.memAvg, n8), alloc(p.memStd, n8)
		p.allocCPU, p.allocMem = alloc(p.allocCPU, n8), alloc(p.allocMem, n8)
		p.flags, p.row = alloc(p.flags, p.table.NPad), alloc(p.row, p.table.NPad)
		p.reqCPU, p.reqMem = alloc(p.reqCPU, 8), alloc(p.reqMem, 8)
		if err != nil {
			return err
		}
	}
	ca, cs, ma, ms := p.cpuAvg.Float64s(p.table.N), p.cpuStd.Float64s(p.table.N), p.memAvg.Float64s(p.table.N), p.memStd.Float64s(p.table.N)
	ac, am, fl := p.allocCPU.Int64s(p.table.N), p.allocMem.Int64s(p.table.N), p.flags.Bytes(p.table.N)
	for i, ni := range all {
		node := ni.Node()
		ca[i], cs[i], ma[i], ms[i], fl[i] = 0, 0, 0, 0, 0
		ac[i] = node.Status.Allocatable.Cpu().MilliValue() // resourcestats.go:55
		am[i] = node.Status.Allocatable.Memory().Value()   // :56
		if metrics, _ := p.collector.GetNodeMetrics(logger, node.Name); metrics != nil {
			fl[i] |= lvrbHasMetrics
			var ok bool
			if ca[i], cs[i], ok = resourceData(metrics, watcher.CPU); ok {
				fl[i] |= lvrbCPUOK
			}
			if ma[i], ms[i], ok = resourceData(metrics, watcher.Memory); ok {
				fl[i] |= lvrbMemOK
			}
		}
		p.table.Gen[i] = ni.GetGeneration()
	}
	if err := p.eng.SnapshotBegin(p.table.Epoch, p.table.N, 0, p.table.N); err != nil {
		return err
	}
	if err := p.eng.SnapshotLVRB(p.cpuAvg, p.cpuStd, p.memAvg, p.memStd, p.allocCPU, p.allocMem, p.flags); err != nil {
		return err
	}
	return p.eng.SnapshotCommit()
}

func (p *LoadVariationRiskBalancingB200) PreScore(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, feasible []fwk.NodeInfo) *fwk.Status {
	logger := klog.FromContext(klog.NewContext(ctx, p.logger)).WithValues("ExtensionPoint", "PreScore")
	all, err := p.handle.SnapshotSharedLister().NodeInfos().List()
	if err == nil {
		err = p.ensureSnapshot(logger, all)
	}
	res := &b200sched.CycleResult{Index: p.table.Index}
	if err == nil {
		req := trimaran.GetResourceRequested(pod) // resourcestats.go:110-146: sum app, max init, + overhead
		p.reqCPU.Int64s(1)[0], p.reqMem.Int64s(1)[0] = req.MilliCPU, req.Memory
		err = p.eng.ScoreBatch(b200sched.PluginLVRB, &b200sched.PodBatch{NPods: 1, LVRBReqCPUMilli: p.reqCPU, LVRBReqMemBytes: p.reqMem},
			b200sched.OutU8, p.row, nil, nil)
	}
	if err != nil {
		logger.V(2).Info("b200sched: falling back to the Go path for this cycle", "err", err)
		res.Fallback = true
	} else {
		res.Scores = append([]uint8(nil), p.row.Bytes(p.table.NPad)...)
	}
	cs.Write(lvrbCycleKey, res)
	return nil
}

func (p *LoadVariationRiskBalancingB200) Score(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, ni fwk.NodeInfo) (int64, *fwk.Status) {
	if d, err := cs.Read(lvrbCycleKey); err == nil {
		if c, ok := d.(*b200sched.CycleResult); ok {
			if i, ok := c.Lookup(ni.Node().Name); ok {
				return int64(c.Scores[i]), fwk.NewStatus(fwk.Success, "")
			}
		}
	}
	return p.LoadVariationRiskBalancing.Score(ctx, cs, pod, ni)
}

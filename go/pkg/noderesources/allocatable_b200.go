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

// NodeResourcesAllocatable behind libb200sched: same registry name, same args type, same Score / NormalizeScore
// results; the per-node arithmetic of allocatable.go:117-168 runs as one kernel launch per scheduling cycle.
// Lives IN package noderesources (next to allocatable.go) because it reuses the unexported scorer types.
// Never compiled in this repository (no Go toolchain) -- see pkg/b200sched/b200sched.go.
package noderesources

import (
	"context"
	"fmt"
	"sort"

	v1 "k8s.io/api/core/v1"
	"k8s.io/apimachinery/pkg/runtime"
	"k8s.io/klog/v2"
	fwk "k8s.io/kube-scheduler/framework"
	schedutil "k8s.io/kubernetes/pkg/scheduler/util"

	"sigs.k8s.io/scheduler-plugins/apis/config"
	"sigs.k8s.io/scheduler-plugins/pkg/b200sched"
)

const allocatableCycleKey = "PreScore" + AllocatableName + "B200"

// AllocatableB200 embeds the original plugin: Name(), ScoreExtensions() and the fallback paths are its own.
type AllocatableB200 struct {
	*Allocatable
	eng    *b200sched.Engine
	table  b200sched.NodeTable
	res    []v1.ResourceName // column order = sorted resource names of resourceToWeightMap
	cols   []*b200sched.Pinned
	row    *b200sched.Pinned // [Npad] u8 scores of the cycle
	feas   *b200sched.Pinned // [Npad/64] words
	patchI *b200sched.Pinned
}

var _ fwk.PreScorePlugin = &AllocatableB200{}
var _ fwk.ScorePlugin = &AllocatableB200{}

// NewAllocatableB200 has the factory signature cmd/scheduler/main.go registers (app.WithPlugin).
func NewAllocatableB200(ctx context.Context, allocArgs runtime.Object, h fwk.Handle) (fwk.Plugin, error) {
	inner, err := NewAllocatable(ctx, allocArgs, h) // validates the args exactly as before
	if err != nil {
		return nil, err
	}
	a := inner.(*Allocatable)
	mode := config.Least
	if args, ok := allocArgs.(*config.NodeResourcesAllocatableArgs); ok && args != nil && args.Mode != "" {
		mode = args.Mode
	}
	eng, err := b200sched.New(0)
	if err != nil {
		klog.FromContext(ctx).Error(err, "b200sched unavailable, NodeResourcesAllocatable stays on the Go path")
		return a, nil
	}
	p := &AllocatableB200{Allocatable: a, eng: eng}
	for r := range a.resourceToWeightMap {
		p.res = append(p.res, r)
	}
	sort.Slice(p.res, func(i, j int) bool { return p.res[i] < p.res[j] })
	weights := make([]int64, len(p.res))
	for i, r := range p.res {
		weights[i] = a.resourceToWeightMap[r]
	}
	m := 0 // B200S_ALLOC_LEAST
	if mode == config.Most {
		m = 1
	}
	if err := eng.ConfigAllocatable(m, weights); err != nil {
		return nil, fmt.Errorf("NodeResourcesAllocatable: %w", err)
	}
	return p, nil
}

// column value of one node: the allocatable side of calculateResourceAllocatableRequest (resource_allocation.go:79-100)
func allocatableColumn(ni fwk.NodeInfo, r v1.ResourceName) int64 {
	switch r {
	case v1.ResourceCPU:
		return ni.GetAllocatable().GetMilliCPU()
	case v1.ResourceMemory:
		return ni.GetAllocatable().GetMemory()
	case v1.ResourceEphemeralStorage:
		return ni.GetAllocatable().GetEphemeralStorage()
	default:
		if schedutil.IsScalarResourceName(r) {
			return ni.GetAllocatable().GetScalarResources()[r]
		}
	}
	return 0
}

func (p *AllocatableB200) ensureSnapshot(all []fwk.NodeInfo) error {
	full, changed := p.table.Diff(all)
	if !full && len(changed) == 0 {
		return nil
	}
	if full {
		p.table.Reset(all)
		for _, c := range p.cols {
			c.Free()
		}
		p.cols = p.cols[:0]
		for range p.res {
			c, err := b200sched.AllocPinned(8 * p.table.NPad)
			if err != nil {
				return err
			}
			p.cols = append(p.cols, c)
		}
		for r, name := range p.res {
			col := p.cols[r].Int64s(p.table.N)
			for i, ni := range all {
				col[i] = allocatableColumn(ni, name)
			}
		}
		if p.row != nil {
			p.row.Free()
			p.feas.Free()
		}
		var err error
		if p.row, err = b200sched.AllocPinned(p.table.NPad); err != nil {
			return err
		}
		if p.feas, err = b200sched.AllocPinned(p.table.NPad / 8); err != nil {
			return err
		}
		if err := p.eng.SnapshotBegin(p.table.Epoch, p.table.N, 0, p.table.N); err != nil {
			return err
		}
		if err := p.eng.SnapshotAllocatable(p.cols); err != nil {
			return err
		}
		return p.eng.SnapshotCommit()
	}
	// a few rows moved: rewrite them in place (b200s_snapshot_patch_allocatable)
	if p.patchI == nil {
		var err error
		if p.patchI, err = b200sched.AllocPinned(4 * p.table.NPad); err != nil {
			return err
		}
	}
	copy(p.patchI.Int32s(len(changed)), changed)
	rows := make([]*b200sched.Pinned, len(p.res))
	for r, name := range p.res {
		buf, err := b200sched.AllocPinned(8 * len(changed))
		if err != nil {
			return err
		}
		defer buf.Free()
		v := buf.Int64s(len(changed))
		for j, i := range changed {
			v[j] = allocatableColumn(all[i], name)
			p.table.Gen[i] = all[i].GetGeneration()
		}
		rows[r] = buf
	}
	if err := p.eng.PatchBegin(p.table.Epoch); err != nil {
		return err
	}
	if err := p.eng.PatchAllocatable(len(changed), p.patchI, rows); err != nil {
		return err
	}
	return p.eng.SnapshotCommit()
}

// PreScore: ONE engine call evaluates every node for this pod; Score is a lookup.
func (p *AllocatableB200) PreScore(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, feasible []fwk.NodeInfo) *fwk.Status {
	all, err := p.handle.SnapshotSharedLister().NodeInfos().List()
	if err == nil {
		err = p.ensureSnapshot(all)
	}
	res := &b200sched.CycleResult{Index: p.table.Index}
	if err == nil {
		p.table.FeasibleWords(p.feas.Uint64s(p.table.NPad/64), feasible)
		err = p.eng.ScoreBatch(b200sched.PluginAllocatable, &b200sched.PodBatch{NPods: 1, Feasible: p.feas}, b200sched.OutU8, p.row, nil, nil)
	}
	if err != nil {
		klog.FromContext(ctx).V(2).Info("b200sched: falling back to the Go path for this cycle", "err", err)
		res.Fallback = true
	} else {
		res.Scores = append([]uint8(nil), p.row.Bytes(p.table.NPad)...) // the pinned row is reused next cycle
	}
	cs.Write(allocatableCycleKey, res)
	return nil
}

func (p *AllocatableB200) Score(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, ni fwk.NodeInfo) (int64, *fwk.Status) {
	if ni.Node() == nil {
		return 0, fwk.NewStatus(fwk.Error, "node not found") // resource_allocation.go:53-56
	}
	if d, err := cs.Read(allocatableCycleKey); err == nil {
		if c, ok := d.(*b200sched.CycleResult); ok {
			if i, ok := c.Lookup(ni.Node().Name); ok {
				return int64(c.Scores[i]), nil // already normalised over the cycle's feasible list
			}
		}
	}
	return p.Allocatable.Score(ctx, cs, pod, ni)
}

// NormalizeScore: the engine normalised over exactly the feasible list of this cycle (allocatable.go:143-168);
// on a fallback cycle the raw Go scores still need the original normalisation.
func (p *AllocatableB200) NormalizeScore(ctx context.Context, cs fwk.CycleState, pod *v1.Pod, scores fwk.NodeScoreList) *fwk.Status {
	if d, err := cs.Read(allocatableCycleKey); err == nil {
		if c, ok := d.(*b200sched.CycleResult); ok && !c.Fallback {
			return nil
		}
	}
	return p.Allocatable.NormalizeScore(ctx, cs, pod, scores)
}

func (p *AllocatableB200) ScoreExtensions() fwk.ScoreExtensions { return p }

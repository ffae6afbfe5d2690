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

package b200sched

import (
	"sync"

	fwk "k8s.io/kube-scheduler/framework"
)

// NodeTable is the host half of one plugin's device snapshot: which NodeInfo sits in which column, and the
// NodeInfo.Generation each row was flattened at.  Upstream refreshes its scheduling snapshot by per-node generation
// (cache.UpdateSnapshot); the wrapper follows it: unchanged rows stay on the device, changed rows go through
// b200s_snapshot_patch_*, a changed node LIST (add / remove / reorder) is a full b200s_snapshot_* upload.
type NodeTable struct {
	mu    sync.Mutex
	Names []string
	Index map[string]int32
	Gen   []int64
	N     int
	NPad  int
	Epoch uint64 // bumps on every full upload: per-cycle results carry it
}

// Diff returns (full, changed): full = the node list differs from the table, changed = rows whose generation moved.
func (t *NodeTable) Diff(nodes []fwk.NodeInfo) (full bool, changed []int32) {
	if len(nodes) != t.N {
		return true, nil
	}
	for i, ni := range nodes {
		n := ni.Node()
		if n == nil || t.Names[i] != n.Name {
			return true, nil
		}
		if g := ni.GetGeneration(); g != t.Gen[i] {
			changed = append(changed, int32(i))
		}
	}
	// past a quarter of the rows one bulk upload is cheaper than the scatter (same rule as the C++ host mirror)
	if len(changed)*4 > t.N {
		return true, nil
	}
	return false, changed
}

func (t *NodeTable) Reset(nodes []fwk.NodeInfo) {
	t.N, t.NPad = len(nodes), NPad(len(nodes))
	t.Names = make([]string, t.N)
	t.Gen = make([]int64, t.N)
	t.Index = make(map[string]int32, t.N)
	for i, ni := range nodes {
		if n := ni.Node(); n != nil {
			t.Names[i] = n.Name
			t.Index[n.Name] = int32(i)
		}
		t.Gen[i] = ni.GetGeneration()
	}
	t.Epoch++
}

// FeasibleWords fills dst ([Npad/64] words, zeroed first) with the nodes upstream's filters kept for this cycle.
func (t *NodeTable) FeasibleWords(dst []uint64, feasible []fwk.NodeInfo) {
	for i := range dst {
		dst[i] = 0
	}
	for _, ni := range feasible {
		if n := ni.Node(); n != nil {
			if i, ok := t.Index[n.Name]; ok {
				dst[i>>6] |= 1 << (uint(i) & 63)
			}
		}
	}
}

// CycleResult is what one engine call leaves in CycleState for the per-node Score / Filter calls: pure lookups, no
// cgo on the per-node path (upstream calls Score / Filter from up to 16 goroutines, once per node).
type CycleResult struct {
	Scores   []uint8  // [Npad] 0..100 (every final score of these plugins)
	Feasible []uint64 // [Npad/64] or nil
	Reasons  []uint8  // [Npad] B200S_REASON_* or nil
	Index    map[string]int32
	Fallback bool // engine error or unsupported shape: the embedded Go plugin answers this cycle
}

func (c *CycleResult) Clone() fwk.StateData { return c }

func (c *CycleResult) Lookup(nodeName string) (idx int32, ok bool) {
	if c == nil || c.Fallback {
		return 0, false
	}
	idx, ok = c.Index[nodeName]
	return
}

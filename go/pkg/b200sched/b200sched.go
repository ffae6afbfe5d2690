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

// Package b200sched binds libb200sched.so (include/b200sched.h) into the scheduler-plugins tree.
//
// STATUS: specification-grade Go.  The image this repository is built in has no Go toolchain and no module cache, so
// these files have never been compiled; they are written against include/b200sched.h (every C identifier below is
// declared there) and against the reference at the revision under /root/reference.  The same call sequences run, and
// are parity-tested, through the ctypes binding (scheduler-plugins_b200/engine.py) and the C++ host mirror
// (scheduler-plugins_b200/host/).  Drop this directory over the reference tree (paths mirror it) and build with
// `make -f Makefile.b200 build-scheduler-b200`.
package b200sched

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/b200sched/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/b200sched/lib -lb200sched -Wl,-rpath,$ORIGIN/lib
#include <stdlib.h>
#include <string.h>
#include "b200sched.h"
*/
import "C"

import (
	"fmt"
	"runtime"
	"sync"
	"unsafe"
)

// Plugin ids, output types and reason codes of include/b200sched.h.
const (
	PluginAllocatable     = C.B200S_PLUGIN_ALLOCATABLE
	PluginTLP             = C.B200S_PLUGIN_TLP
	PluginLVRB            = C.B200S_PLUGIN_LVRB
	PluginNRT             = C.B200S_PLUGIN_NRT
	PluginNetworkOverhead = C.B200S_PLUGIN_NETWORK_OVERHEAD

	OutI64 = C.B200S_OUT_I64
	OutU8  = C.B200S_OUT_U8

	ReasonOK                 = C.B200S_REASON_OK
	ReasonNRTInvalidTopology = C.B200S_REASON_NRT_INVALID_TOPOLOGY
	ReasonNRTAlignPod        = C.B200S_REASON_NRT_ALIGN_POD
	ReasonNRTAlignContainer  = C.B200S_REASON_NRT_ALIGN_CONTAINER
	ReasonNRTAlignInit       = C.B200S_REASON_NRT_ALIGN_INIT
	ReasonNRTAlignSidecar    = C.B200S_REASON_NRT_ALIGN_SIDECAR
	ReasonNRTAccounting      = C.B200S_REASON_NRT_ACCOUNTING
	ReasonNetOHViolated      = C.B200S_REASON_NETOH_VIOLATED
	ReasonUpstream           = C.B200S_REASON_UPSTREAM
	ReasonUnsupported        = C.B200S_REASON_UNSUPPORTED

	NodeAlign = C.B200S_NODE_ALIGN
)

// NPad is the row pitch of every engine matrix for n nodes.
func NPad(n int) int {
	if n < 1 {
		n = 1
	}
	return (n + NodeAlign - 1) / NodeAlign * NodeAlign
}

// Engine is one b200s_ctx: one GPU, one shard of the node axis, one scheduler profile.
// All methods are safe from any goroutine (the ctx serialises internally); cgo may migrate goroutines between OS
// threads and the library sets the device on entry, so no LockOSThread is needed.
type Engine struct {
	ctx *C.b200s_ctx
	mu  sync.Mutex // guards the staging buffers below, not the ctx
}

func New(device int) (*Engine, error) {
	var ctx *C.b200s_ctx
	if rc := C.b200s_init(C.int(device), &ctx); rc != C.B200S_OK {
		return nil, fmt.Errorf("b200s_init(device %d) = %d: no usable GPU (the engine has no CPU fallback)", device, int(rc))
	}
	e := &Engine{ctx: ctx}
	runtime.SetFinalizer(e, func(e *Engine) { e.Close() })
	return e, nil
}

func (e *Engine) Close() {
	if e.ctx != nil {
		C.b200s_shutdown(e.ctx)
		e.ctx = nil
	}
}

func (e *Engine) err(rc C.int, what string) error {
	if rc == C.B200S_OK {
		return nil
	}
	return fmt.Errorf("%s: b200sched error %d: %s", what, int(rc), C.GoString(C.b200s_last_error(e.ctx)))
}

// Pinned is caller-owned page-locked host memory (b200s_alloc_pinned).  cgo forbids C code from retaining Go
// pointers and the engine's H2D copies are asynchronous, so every column travels through one of these.
type Pinned struct {
	p     unsafe.Pointer
	bytes int
}

func AllocPinned(bytes int) (*Pinned, error) {
	p := C.b200s_alloc_pinned(C.size_t(bytes))
	if p == nil {
		return nil, fmt.Errorf("b200s_alloc_pinned(%d) failed", bytes)
	}
	return &Pinned{p: p, bytes: bytes}, nil
}
func (b *Pinned) Free()                  { C.b200s_free_pinned(b.p); b.p = nil }
func (b *Pinned) Ptr() unsafe.Pointer    { return b.p }
func (b *Pinned) Int64s(n int) []int64   { return unsafe.Slice((*int64)(b.p), n) }
func (b *Pinned) Float64s(n int) []float64 { return unsafe.Slice((*float64)(b.p), n) }
func (b *Pinned) Uint64s(n int) []uint64 { return unsafe.Slice((*uint64)(b.p), n) }
func (b *Pinned) Uint16s(n int) []uint16 { return unsafe.Slice((*uint16)(b.p), n) }
func (b *Pinned) Int32s(n int) []int32   { return unsafe.Slice((*int32)(b.p), n) }
func (b *Pinned) Bytes(n int) []byte     { return unsafe.Slice((*byte)(b.p), n) }

// ---- snapshot -----------------------------------------------------------------------------------------------------

func (e *Engine) SnapshotBegin(generation uint64, nNodes, nodeOffset, nNodesGlobal int) error {
	return e.err(C.b200s_snapshot_begin(e.ctx, C.uint64_t(generation), C.int32_t(nNodes), C.int32_t(nodeOffset),
		C.int32_t(nNodesGlobal)), "snapshot_begin")
}

// SnapshotAllocatable: cols[r] holds nNodes int64 (cpu milli, memory / ephemeral-storage bytes, scalars in units:
// calculateResourceAllocatableRequest, pkg/noderesources/resource_allocation.go:79-100).
func (e *Engine) SnapshotAllocatable(cols []*Pinned) error {
	ptrs := (**C.int64_t)(C.malloc(C.size_t(len(cols)) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(ptrs))
	arr := unsafe.Slice((**C.int64_t)(unsafe.Pointer(ptrs)), len(cols))
	for i, c := range cols {
		arr[i] = (*C.int64_t)(c.p)
	}
	return e.err(C.b200s_snapshot_allocatable(e.ctx, C.int32_t(len(cols)), ptrs), "snapshot_allocatable")
}

func (e *Engine) SnapshotTLP(cpuUtilPct, capMilli, missingMilli, flags *Pinned) error {
	return e.err(C.b200s_snapshot_tlp(e.ctx, (*C.double)(cpuUtilPct.p), (*C.int64_t)(capMilli.p), (*C.int64_t)(missingMilli.p),
		(*C.uint8_t)(flags.p)), "snapshot_tlp")
}

func (e *Engine) SnapshotLVRB(cpuAvg, cpuStd, memAvg, memStd, allocCPUMilli, allocMemBytes, flags *Pinned) error {
	return e.err(C.b200s_snapshot_lvrb(e.ctx, (*C.double)(cpuAvg.p), (*C.double)(cpuStd.p), (*C.double)(memAvg.p),
		(*C.double)(memStd.p), (*C.int64_t)(allocCPUMilli.p), (*C.int64_t)(allocMemBytes.p), (*C.uint8_t)(flags.p)), "snapshot_lvrb")
}

// NRTNodes mirrors b200s_nrt_nodes; every slice header points into pinned memory.
type NRTNodes struct {
	NZones, NRes                                                                 int
	ResFlags, NodeFlags, MaxNUMA, NZonesNode, NodeResMask, ZoneResMask, Avail, Cost *Pinned // Cost may be nil
}

func (n *NRTNodes) c() C.b200s_nrt_nodes {
	var s C.b200s_nrt_nodes
	s.n_zones, s.n_res = C.int32_t(n.NZones), C.int32_t(n.NRes)
	s.res_flags = (*C.uint8_t)(n.ResFlags.p)
	s.node_flags = (*C.uint8_t)(n.NodeFlags.p)
	s.max_numa = (*C.uint16_t)(n.MaxNUMA.p)
	s.n_zones_node = (*C.uint8_t)(n.NZonesNode.p)
	s.node_res_mask = (*C.uint8_t)(n.NodeResMask.p)
	s.zone_res_mask = (*C.uint8_t)(n.ZoneResMask.p)
	s.avail = (*C.int64_t)(n.Avail.p)
	if n.Cost != nil {
		s.cost = (*C.int32_t)(n.Cost.p)
	}
	return s
}

func (e *Engine) SnapshotNRT(n *NRTNodes) error {
	s := n.c()
	return e.err(C.b200s_snapshot_nrt(e.ctx, &s), "snapshot_nrt")
}

func (e *Engine) SnapshotNetworkOverhead(regionID, zoneID *Pinned, nNames int, zoneCost, regionCost *Pinned) error {
	return e.err(C.b200s_snapshot_network_overhead(e.ctx, (*C.uint16_t)(regionID.p), (*C.uint16_t)(zoneID.p), C.int32_t(nNames),
		(*C.int64_t)(zoneCost.p), (*C.int64_t)(regionCost.p)), "snapshot_network_overhead")
}

func (e *Engine) SnapshotCommit() error { return e.err(C.b200s_snapshot_commit(e.ctx), "snapshot_commit") }

// ---- incremental snapshot (rows whose NodeInfo.Generation moved) --------------------------------------------------

func (e *Engine) PatchBegin(generation uint64) error {
	return e.err(C.b200s_snapshot_patch_begin(e.ctx, C.uint64_t(generation)), "snapshot_patch_begin")
}
func (e *Engine) PatchAllocatable(count int, nodeIdx *Pinned, cols []*Pinned) error {
	ptrs := (**C.int64_t)(C.malloc(C.size_t(len(cols)) * C.size_t(unsafe.Sizeof(uintptr(0)))))
	defer C.free(unsafe.Pointer(ptrs))
	arr := unsafe.Slice((**C.int64_t)(unsafe.Pointer(ptrs)), len(cols))
	for i, c := range cols {
		arr[i] = (*C.int64_t)(c.p)
	}
	return e.err(C.b200s_snapshot_patch_allocatable(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p), C.int32_t(len(cols)), ptrs),
		"snapshot_patch_allocatable")
}
func (e *Engine) PatchTLP(count int, nodeIdx, cpuUtilPct, capMilli, missingMilli, flags *Pinned) error {
	return e.err(C.b200s_snapshot_patch_tlp(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p), (*C.double)(cpuUtilPct.p),
		(*C.int64_t)(capMilli.p), (*C.int64_t)(missingMilli.p), (*C.uint8_t)(flags.p)), "snapshot_patch_tlp")
}
func (e *Engine) PatchLVRB(count int, nodeIdx, cpuAvg, cpuStd, memAvg, memStd, allocCPUMilli, allocMemBytes, flags *Pinned) error {
	return e.err(C.b200s_snapshot_patch_lvrb(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p), (*C.double)(cpuAvg.p),
		(*C.double)(cpuStd.p), (*C.double)(memAvg.p), (*C.double)(memStd.p), (*C.int64_t)(allocCPUMilli.p),
		(*C.int64_t)(allocMemBytes.p), (*C.uint8_t)(flags.p)), "snapshot_patch_lvrb")
}
func (e *Engine) PatchNRT(count int, nodeIdx *Pinned, rows *NRTNodes) error {
	s := rows.c()
	return e.err(C.b200s_snapshot_patch_nrt(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p), &s), "snapshot_patch_nrt")
}

// PatchNRTDeduct is resourceStore.UpdateNRT (pkg/noderesourcetopology/cache/store.go:129-160) as a column operation.
func (e *Engine) PatchNRTDeduct(count int, nodeIdx, resMask, deduct *Pinned) error {
	return e.err(C.b200s_snapshot_patch_nrt_deduct(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p), (*C.uint8_t)(resMask.p),
		(*C.int64_t)(deduct.p)), "snapshot_patch_nrt_deduct")
}
func (e *Engine) PatchNetworkOverhead(count int, nodeIdx, regionID, zoneID *Pinned) error {
	return e.err(C.b200s_snapshot_patch_network_overhead(e.ctx, C.int32_t(count), (*C.int32_t)(nodeIdx.p),
		(*C.uint16_t)(regionID.p), (*C.uint16_t)(zoneID.p)), "snapshot_patch_network_overhead")
}

// ---- plugin args ---------------------------------------------------------------------------------------------------

func (e *Engine) ConfigAllocatable(mode int, weights []int64) error {
	w := (*C.int64_t)(C.malloc(C.size_t(8 * len(weights))))
	defer C.free(unsafe.Pointer(w))
	copy(unsafe.Slice((*int64)(unsafe.Pointer(w)), len(weights)), weights)
	return e.err(C.b200s_config_allocatable(e.ctx, C.int(mode), C.int32_t(len(weights)), w), "config_allocatable")
}
func (e *Engine) ConfigTLP(targetUtilizationPct int64) error {
	return e.err(C.b200s_config_tlp(e.ctx, C.int64_t(targetUtilizationPct)), "config_tlp")
}
func (e *Engine) ConfigLVRB(margin, sensitivity float64) error {
	return e.err(C.b200s_config_lvrb(e.ctx, C.double(margin), C.double(sensitivity)), "config_lvrb")
}
func (e *Engine) ConfigNRT(strategy int, weights []int64) error {
	var w *C.int64_t
	if len(weights) > 0 {
		w = (*C.int64_t)(C.malloc(C.size_t(8 * len(weights))))
		defer C.free(unsafe.Pointer(w))
		copy(unsafe.Slice((*int64)(unsafe.Pointer(w)), len(weights)), weights)
	}
	return e.err(C.b200s_config_nrt(e.ctx, C.int(strategy), C.int32_t(len(weights)), w), "config_nrt")
}
func (e *Engine) ConfigNetworkOverhead(wantCounts, applyOwnFilter bool) error {
	b := func(v bool) C.int {
		if v {
			return 1
		}
		return 0
	}
	return e.err(C.b200s_config_network_overhead(e.ctx, b(wantCounts), b(applyOwnFilter)), "config_network_overhead")
}

// ---- the per-cycle call ---------------------------------------------------------------------------------------------

// PodBatch mirrors b200s_pod_batch for ONE scheduling cycle (n_pods = 1) or a harness batch.  All pointers are pinned.
type PodBatch struct {
	NPods                                                     int
	Feasible                                                  *Pinned // [P][Npad/64] words or nil = every node
	TLPPodCPUMilli, LVRBReqCPUMilli, LVRBReqMemBytes          *Pinned
	NRT                                                       *NRTPods
	NetOH                                                     *NetOHPods
}
type NRTPods struct{ QoS, Flags, NInit, NApp, ContKind, ReqMask, Req *Pinned }
type NetOHPods struct{ ScoreEqually, DepOffset, Deps *Pinned }

// ScoreBatch = b200s_score_batch: upload + evaluate + fetch, HOST buffers in and out, one synchronisation.
// scores holds NPods*Npad elements of dtype; feasible / reasons may be nil (score-only plugins).
func (e *Engine) ScoreBatch(plugin int, b *PodBatch, dtype int, scores, feasible, reasons *Pinned) error {
	cb := (*C.b200s_pod_batch)(C.calloc(1, C.size_t(unsafe.Sizeof(C.b200s_pod_batch{}))))
	defer C.free(unsafe.Pointer(cb))
	cb.n_pods = C.int32_t(b.NPods)
	if b.Feasible != nil {
		cb.feasible = (*C.uint64_t)(b.Feasible.p)
	}
	if b.TLPPodCPUMilli != nil {
		cb.tlp_pod_cpu_milli = (*C.int64_t)(b.TLPPodCPUMilli.p)
	}
	if b.LVRBReqCPUMilli != nil {
		cb.lvrb_req_cpu_milli = (*C.int64_t)(b.LVRBReqCPUMilli.p)
		cb.lvrb_req_mem_bytes = (*C.int64_t)(b.LVRBReqMemBytes.p)
	}
	if b.NRT != nil {
		n := (*C.b200s_nrt_pods)(C.calloc(1, C.size_t(unsafe.Sizeof(C.b200s_nrt_pods{}))))
		defer C.free(unsafe.Pointer(n))
		n.qos, n.flags = (*C.uint8_t)(b.NRT.QoS.p), (*C.uint8_t)(b.NRT.Flags.p)
		n.n_init, n.n_app = (*C.uint8_t)(b.NRT.NInit.p), (*C.uint8_t)(b.NRT.NApp.p)
		n.cont_kind, n.req_mask = (*C.uint8_t)(b.NRT.ContKind.p), (*C.uint8_t)(b.NRT.ReqMask.p)
		n.req = (*C.int64_t)(b.NRT.Req.p)
		cb.nrt = n
	}
	if b.NetOH != nil {
		n := (*C.b200s_netoh_pods)(C.calloc(1, C.size_t(unsafe.Sizeof(C.b200s_netoh_pods{}))))
		defer C.free(unsafe.Pointer(n))
		n.score_equally = (*C.uint8_t)(b.NetOH.ScoreEqually.p)
		n.dep_offset = (*C.int32_t)(b.NetOH.DepOffset.p)
		n.deps = (*C.b200s_netoh_dep)(b.NetOH.Deps.p)
		cb.netoh = n
	}
	var f *C.uint64_t
	var r *C.uint8_t
	if feasible != nil {
		f = (*C.uint64_t)(feasible.p)
	}
	if reasons != nil {
		r = (*C.uint8_t)(reasons.p)
	}
	return e.err(C.b200s_score_batch(e.ctx, C.b200s_plugin(plugin), cb, C.b200s_out_dtype(dtype), scores.p, f, r), "score_batch")
}

// NetworkOverhead's PreFilterState after the last evaluation: finalCostMap and satisfied | violated << 16.
func (e *Engine) FetchNetworkOverheadRaw(out *Pinned, bytes int) error {
	return e.err(C.b200s_fetch_network_overhead_raw(e.ctx, (*C.int64_t)(out.p), C.size_t(bytes)), "fetch_network_overhead_raw")
}
func (e *Engine) FetchNetworkOverheadCounts(out *Pinned, bytes int) error {
	return e.err(C.b200s_fetch_network_overhead_counts(e.ctx, (*C.uint32_t)(out.p), C.size_t(bytes)), "fetch_network_overhead_counts")
}

// ---- engine-only profiles: weighted sum + per-pod top-k ---------------------------------------------------------------

type TopK struct {
	Score int64
	Node  int32 // GLOBAL node index, -1 = no feasible node
	_     int32
}

func (e *Engine) PodsUpload(b *C.b200s_pod_batch) error { return e.err(C.b200s_pods_upload(e.ctx, b), "pods_upload") }
func (e *Engine) EvalCombined(pluginMask uint32, weights [C.B200S_PLUGIN_COUNT]int64, k int, writeTotal bool) error {
	wt := C.int(0)
	if writeTotal {
		wt = 1
	}
	return e.err(C.b200s_eval_combined(e.ctx, C.uint32_t(pluginMask), (*C.int64_t)(unsafe.Pointer(&weights[0])), C.int32_t(k), wt),
		"eval_combined")
}
func (e *Engine) FetchTopK(out *Pinned, bytes int) error {
	return e.err(C.b200s_fetch_topk(e.ctx, (*C.b200s_topk_entry)(out.p), C.size_t(bytes)), "fetch_topk")
}

// ScheduleBatch is the engine-only profile in one call with one synchronisation: upload the batch's pod columns,
// evaluate the weighted combination, copy the [NPods][k] winners to out (pinned).  For NPods <= 4 on one GPU this is
// two kernel launches and no per-plugin matrix (cycle.cu).
func (e *Engine) ScheduleBatch(b *C.b200s_pod_batch, pluginMask uint32, weights [C.B200S_PLUGIN_COUNT]int64, k int, out *Pinned) error {
	return e.err(C.b200s_schedule_batch(e.ctx, b, C.uint32_t(pluginMask), (*C.int64_t)(unsafe.Pointer(&weights[0])), C.int32_t(k),
		(*C.b200s_topk_entry)(out.p)), "schedule_batch")
}

// ScheduleSequence places a whole batch pod by pod with the assume step on the device (the OverReserve deduction of
// overreserve.go:148-182 / store.go:129-160 and the Trimaran bind cache of handler.go:131-167).  The resident snapshot
// is modified in place: a failed bind resyncs that node's rows through the patch calls.
func (e *Engine) ScheduleSequence(b *C.b200s_pod_batch, pluginMask uint32, weights [C.B200S_PLUGIN_COUNT]int64, out *Pinned) error {
	return e.err(C.b200s_schedule_sequence(e.ctx, b, C.uint32_t(pluginMask), (*C.int64_t)(unsafe.Pointer(&weights[0])),
		(*C.b200s_topk_entry)(out.p)), "schedule_sequence")
}

// ---- multi-GPU: one Engine per shard ---------------------------------------------------------------------------------

func CommUniqueID() ([C.B200S_UNIQUE_ID_BYTES]byte, error) {
	var id [C.B200S_UNIQUE_ID_BYTES]byte
	if rc := C.b200s_comm_unique_id(unsafe.Pointer(&id[0])); rc != C.B200S_OK {
		return id, fmt.Errorf("b200s_comm_unique_id = %d", int(rc))
	}
	return id, nil
}
func (e *Engine) CommInit(id [C.B200S_UNIQUE_ID_BYTES]byte, rank, world int) error {
	return e.err(C.b200s_comm_init(e.ctx, unsafe.Pointer(&id[0]), C.int(rank), C.int(world)), "comm_init")
}
func (e *Engine) PeerExport() ([C.B200S_PEER_HANDLE_BYTES]byte, error) {
	var h [C.B200S_PEER_HANDLE_BYTES]byte
	return h, e.err(C.b200s_comm_peer_export(e.ctx, unsafe.Pointer(&h[0])), "comm_peer_export")
}
func (e *Engine) PeerImport(handles []byte) error {
	p := C.CBytes(handles)
	defer C.free(p)
	return e.err(C.b200s_comm_peer_import(e.ctx, p), "comm_peer_import")
}

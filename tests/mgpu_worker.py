"""Multi-GPU parity worker, launched by torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/mgpu_worker.py
Every rank holds one shard of the node axis in its own engine ctx; the engine's NCCL communicator
does the per-pod min/max all-reduce (NodeResourcesAllocatable / NetworkOverhead NormalizeScore) and
the single all-gather of the per-pod top-k winners.  Each rank compares its shard of every score
matrix and the folded global top-k with the UNSHARDED oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import pyoracle as orc
    from scheduler_plugins_b200 import engine as E
    from scheduler_plugins_b200 import sharding, synth
    from test_gpu_combined import build_inputs, load_engine, oracle_combined

    P, N, K = 40, 128 * 7 * world + 77, 3
    seed = synth.BASE_SEED + 5
    d = build_inputs(seed, P, N)
    feas_full = synth.gen_feasible_words(seed, P, N, E.npad_of(N))
    fb = E.unpack_bits(feas_full, N)
    off, cnt = sharding.shard_bounds(N, world)[rank]
    feas = E.pack_bits(fb[:, off:off + cnt], E.npad_of(cnt))

    eng = E.Engine(local)
    uid = [eng.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    if os.environ.get("B200S_TEST_PEER", "1") == "1":  # the per-pod exchanges over peer memory instead of NCCL
        handles = [None] * world
        dist.all_gather_object(handles, eng.peer_export())
        eng.peer_import(handles)
    load_engine(eng, E, d, cnt, P, feas, node_offset=off, n_global=N)
    weights = [2, 1, 1, 3, 5]

    # --- NodeResourcesAllocatable alone: sharded NormalizeScore == unsharded
    eng.eval(E.PLUGIN_ALLOCATABLE)
    got = eng.fetch_scores(E.PLUGIN_ALLOCATABLE)[:, :cnt]
    want = orc.alloc_batch([d["nodes"]["alloc_cpu_milli"], d["nodes"]["alloc_mem_bytes"]], [1 << 20, 1], 1, P, feas_full,
                           pitch=E.npad_of(N))
    assert np.array_equal(got, want[:, off:off + cnt]), f"rank {rank}: sharded Allocatable != unsharded oracle"

    # --- NetworkOverhead alone (global host-node indices, min/max all-reduce)
    eng.eval(E.PLUGIN_NETWORK_OVERHEAD)
    got = eng.fetch_scores(E.PLUGIN_NETWORK_OVERHEAD)[:, :cnt]
    net = d["net"]
    ws, wf, _ = orc.netoh_batch(net["zone_cost"], net["region_cost"], net["region_all"], net["zone_all"],
                                net["score_equally"], net["dep_offset"], net["deps"], feas_full, pitch=E.npad_of(N))
    assert np.array_equal(got, ws[:, off:off + cnt]), f"rank {rank}: sharded NetworkOverhead != unsharded oracle"

    # --- combined profile: total matrix shard + global top-k on every rank
    eng.eval_combined(0b11111, weights, k=K, write_total=True)
    want_total, want_feas, want_topk = oracle_combined(d, P, N, E.npad_of(N), feas_full, weights, K, 0b11111)
    assert np.array_equal(eng.fetch_total()[:, :cnt], want_total[:, off:off + cnt]), f"rank {rank}: total differs"
    got_topk = eng.fetch_topk()
    for p in range(P):
        assert [(int(e["score"]), int(e["node"])) for e in got_topk[p]] == want_topk[p], (rank, p)
    # --- chunked b200s_score_batch on UNEVEN shards (this N gives the ranks different Npad): the batch is large enough
    # for the pod-chunk pipeline, every chunk is one min/max exchange, so chunked-or-not and the chunk count must come
    # from rank-invariant values -- ranks that disagreed would mismatch the exchanges (hang or wrong normalisation)
    P2 = 13_000
    feas2_full = synth.gen_feasible_words(seed + 1, P2, N, E.npad_of(N))
    feas2 = E.pack_bits(E.unpack_bits(feas2_full, N)[:, off:off + cnt], E.npad_of(cnt))
    batch2, _keep2 = eng.make_batch(P2, feasible=feas2)
    out2 = np.empty((P2, eng.Npad), dtype=np.int64)
    assert P2 * E.npad_of(-(-N // world)) * 8 >= (96 << 20)  # the chunked path (engine.cu: >= 2 x 48 MiB of scores)
    eng.score_batch(E.PLUGIN_ALLOCATABLE, batch2, E.OUT_I64, out2)
    want2 = orc.alloc_batch([d["nodes"]["alloc_cpu_milli"], d["nodes"]["alloc_mem_bytes"]], [1 << 20, 1], 1, P2, feas2_full,
                            pitch=E.npad_of(N))
    assert np.array_equal(out2[:, :cnt], want2[:, off:off + cnt]), f"rank {rank}: chunked sharded score_batch differs"
    # --- Peaks: NormalizeScore's min/max over the feasible set crosses the shards (one all-reduce between the passes)
    nodes = d["nodes"]
    tri, t2 = synth.gen_trimaran(seed, nodes), synth.gen_trimaran2(seed, nodes, P)
    sl = slice(off, off + cnt)
    eng.snapshot_begin(cnt, node_offset=off, n_nodes_global=N)
    eng.snapshot_peaks(tri["cpu_avg"][sl], nodes["cap_cpu_milli"][sl], tri["tlp_flags"][sl], t2["k1"][sl], t2["k2"][sl])
    eng.snapshot_commit()
    eng.pods_upload(P, feasible=feas, peaks_pod_cpu_milli=t2["peaks_pod_cpu_milli"])
    eng.eval(E.PLUGIN_PEAKS)
    want = orc.peaks_batch(tri["cpu_avg"], nodes["cap_cpu_milli"], tri["tlp_flags"], t2["k1"], t2["k2"],
                           t2["peaks_pod_cpu_milli"], feas_full, pitch=E.npad_of(N))
    assert np.array_equal(eng.fetch_scores(E.PLUGIN_PEAKS)[:, :cnt], want[:, sl]), f"rank {rank}: sharded Peaks differs"
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: world={world} P={P} N={N} shards={sharding.shard_bounds(N, world)} "
              f"exchange={'peer' if os.environ.get('B200S_TEST_PEER', '1') == '1' else 'nccl'}")
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

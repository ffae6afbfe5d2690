#!/usr/bin/env python
"""Where does a P = 1 cycle (all five plugins + top-1, 50k nodes) spend its time?  Wall-clock p50 of the pieces of
b200s_schedule_batch, and the two cycle kernels' own duration (CUDA events)."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
from scheduler_plugins_b200 import engine as E, synth  # noqa: E402
from test_gpu_combined import build_inputs, load_engine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
d = build_inputs(synth.BASE_SEED + 5, 1, N)
d["net"]["score_equally"][:] = 0
eng = E.Engine(0)
load_engine(eng, E, d, N, 1, None)
w = [1, 1, 1, 1, 5]
batch, keep = eng.make_batch(1, tlp_pod_cpu_milli=d["pods"]["tlp_pod_cpu_milli"], lvrb_req_cpu_milli=d["pods"]["req_cpu_milli"],
                             lvrb_req_mem_bytes=d["pods"]["req_mem_bytes"], nrt=d["nrt_pods"], netoh=d["net"])
out = np.empty((1, 1), dtype=E.TOPK_DTYPE)


def p50(fn, n=2000):
    for _ in range(50):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.percentile(np.array(ts) * 1e6, 50))


res = {"nodes": N}
res["schedule_batch_us"] = p50(lambda: eng.schedule_batch(batch, 0b11111, w, 1, out))
res["schedule_batch_prepared_call_us"] = p50(eng.prepare_schedule_batch(batch, 0b11111, w, 1, out))
eng.config_fused_cycle(2)
res["schedule_batch_prepared_call_plain_launches_us"] = p50(eng.prepare_schedule_batch(batch, 0b11111, w, 1, out))
eng.config_fused_cycle(True)
res["pods_upload_us"] = p50(lambda: eng._chk(eng.lib.b200s_pods_upload(eng.ctx, C.byref(batch))))
eng.P = 1
res["eval_combined_plus_sync_us"] = p50(lambda: (eng.eval_combined(0b11111, w, 1, False), eng.sync()))
res["fetch_topk_us"] = p50(lambda: eng.fetch_topk())
res["sync_only_us"] = p50(lambda: eng.sync())
for m, name in ((0b00001, "alloc"), (0b00010, "tlp"), (0b00100, "lvrb"), (0b01000, "nrt"), (0b10000, "netoh"), (0b11111, "all")):
    eng.set_profiling(True)
    for _ in range(200):
        eng.eval_combined(m, w, 1, False)
    eng.sync()
    ms, n = eng.phase_time(E.PHASE_COMBINE)
    eng.set_profiling(False)
    res[f"kernel_us_{name}"] = ms * 1e3 / max(n, 1)
print(json.dumps(res))

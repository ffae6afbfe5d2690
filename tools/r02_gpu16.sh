set -u
O=gpurun_out/r02o; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_combined.py tests/test_gpu_fuzz.py tests/test_gpu_host_plugins.py tests/test_gpu_integration_scenarios.py tests/test_gpu_full_size.py tests/test_gpu_nrt.py -q -m gpu -x > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -30 $O/tests.log | cut -c1-300
timeout 900 python bench.py --config c2 --steps 20 --cycles 1000 > $O/bench_c2.json 2> $O/bench_c2.err; python - <<PY
import json
d=json.loads(open("$O/bench_c2.json").read().strip().splitlines()[-1])
print(json.dumps(d["cycle_latency"])[:1500]); print(d["value"], d["parity_checked"], d["parity_errors"])
PY
tail -3 $O/bench_c2.err

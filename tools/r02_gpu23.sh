# cycle graph check: combined tests + c2 cycle latency
set -u
O=gpurun_out/r23; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_combined.py tests/test_gpu_host_plugins.py tests/test_gpu_snapshot_patch.py -q -m gpu -x > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -12 $O/tests.log
timeout 300 python tools/cycle_breakdown.py > $O/breakdown.json 2> $O/breakdown.err; cat $O/breakdown.json
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r23/bench_c2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_checked') if k in d}); print(json.dumps(d.get('cycle_latency'),indent=0)[:1800])
PY
tail -5 $O/bench_c2.err

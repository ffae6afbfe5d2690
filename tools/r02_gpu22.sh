# device-side gcd statistics check: NRT / patch / combined tests + c2 cycle latency + c4 timing
set -u
O=gpurun_out/r22; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nrt.py tests/test_gpu_nrt_batched.py tests/test_gpu_snapshot_patch.py tests/test_gpu_combined.py tests/test_gpu_host_plugins.py tests/test_gpu_full_size.py -q -m gpu > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -8 $O/tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r22/bench_c2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','parity_checked') if k in d}); print(json.dumps(d.get('cycle_latency'),indent=0)[:1500])
PY
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; head -c 700 $O/bench_c4.json

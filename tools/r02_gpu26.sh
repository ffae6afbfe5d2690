# bounded min/max scan with the coalesced fallback: Allocatable parity + c5 timing
set -u
O=gpurun_out/r26; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_combined.py tests/test_gpu_fuzz.py -q -m gpu > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -5 $O/tests.log
timeout 300 python bench.py --config c5 --steps 3 --warmup 1 > $O/bench_c5.json 2> $O/bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r26/bench_c5.json').read().strip().splitlines()[-1])
print('c5', d['ms_per_step'], d.get('parity_checked'), d.get('parity_errors'), d['roofline'].get('per_plugin_kernel_ms'), d['roofline'].get('phase_ms_per_step'), 'e2e', d['e2e'].get('ms_per_step'))
PY

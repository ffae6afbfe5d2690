# quotient-table form of the NRT table kernels: parity + c4 timing (A/B against B200S_NRT2_Q=0)
set -u
O=gpurun_out/r24; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nrt_batched.py tests/test_gpu_nrt.py tests/test_gpu_full_size.py -q -m gpu > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -8 $O/tests.log
timeout 300 python bench.py --config c4 --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err; python - <<'PY'
import json
for f in ('bench_c4',):
    d=json.loads(open(f'gpurun_out/r24/{f}.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d.get('parity_checked'), d.get('per_plugin_kernel_ms'), d['roofline'])
PY
timeout 300 python bench.py --config c5 --steps 2 --warmup 1 > $O/bench_c5.json 2> $O/bench_c5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r24/bench_c5.json').read().strip().splitlines()[-1])
print('c5', d['ms_per_step'], d.get('parity_checked'), d.get('parity_errors'), d['roofline'].get('per_plugin_kernel_ms'), d['roofline'].get('phase_ms_per_step'))
PY

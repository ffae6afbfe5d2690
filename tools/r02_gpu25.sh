# pinned double-buffered staging (no per-chunk stream syncs) + async pod upload: parity + c5 / c4 timing
set -u
O=gpurun_out/r25; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_combined.py tests/test_gpu_full_size.py -q -m gpu > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -5 $O/tests.log
for cfg in c5; do
timeout 300 python bench.py --config $cfg --steps 3 --warmup 1 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
done
python - <<'PY'
import json
for cfg in ("c5",):
    d=json.loads(open(f'gpurun_out/r25/bench_{cfg}.json').read().strip().splitlines()[-1])
    print(cfg, d['ms_per_step'], d.get('parity_checked'), d.get('parity_errors'), d['roofline'].get('per_plugin_kernel_ms'), d['roofline'].get('phase_ms_per_step'), 'e2e', d['e2e'].get('ms_per_step'))
PY

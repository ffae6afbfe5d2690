# iteration helper: NRT tests + c4 timings (arg 1 = output dir name under gpurun_out)
set -u
O=gpurun_out/${1:-nrt_iter}; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_nrt.py tests/test_gpu_nrt_batched.py tests/test_gpu_snapshot_patch.py tests/test_gpu_fuzz.py tests/test_gpu_combined.py tests/test_gpu_divcheck.py tests/test_gpu_host_plugins.py tests/test_gpu_parity.py -q -m gpu > $O/tests.log 2>&1; echo "pytest rc=$?" >> $O/tests.log
tail -15 $O/tests.log
timeout 600 python tools/measure_configs.py --configs c4,c3 --out $O/configs.json > $O/configs.log 2>&1
cut -c1-330 $O/configs.log

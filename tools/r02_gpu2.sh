set -u
O=gpurun_out/r02b; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_nrt.py tests/test_gpu_nrt_batched.py -q -m gpu -x > $O/nrt_tests.log 2>&1; echo "pytest rc=$?" >> $O/nrt_tests.log
tail -25 $O/nrt_tests.log
timeout 600 python tools/measure_configs.py --configs c4 --out $O/configs_c4.json > $O/configs_c4.log 2>&1
cat $O/configs_c4.log | cut -c1-400

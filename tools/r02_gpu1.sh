set -u
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a/gpu.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/r02a/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/gpu_tests.log
tail -3 gpurun_out/r02a/gpu_tests.log
timeout 600 python tools/measure_configs.py --out gpurun_out/r02a/configs_baseline.json > gpurun_out/r02a/configs.log 2>&1
bash tools/r02_ncu_kernels.sh

set -u
O=gpurun_out/r02c; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_nrt.py tests/test_gpu_nrt_batched.py tests/test_gpu_snapshot_patch.py tests/test_gpu_fuzz.py tests/test_gpu_combined.py -q -m gpu > $O/nrt_tests.log 2>&1; echo "pytest rc=$?" >> $O/nrt_tests.log
tail -30 $O/nrt_tests.log
cap() {  # name regex skip count
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $3 -c $4 -f -o $O/$1 \
    python tools/measure_configs.py --configs c4 --steps 1 > $O/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
}
cap nrt2_expand nrt2_expand_kernel 2 1
cap nrt2_table nrt2_table_kernel 4 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nrt2 -c 40 --csv --log-file $O/nrt2_launches.csv python tools/measure_configs.py --configs c4 --steps 1 > $O/launches_run.log 2>&1
grep -c nrt2 $O/nrt2_launches.csv

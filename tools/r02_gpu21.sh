set -u
O=gpurun_out/r02p; mkdir -p $O
# compute-sanitizer over the tests that drive the kernels written this round (batched NRT, fused cycle, sequence)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_nrt.py tests/test_gpu_nrt_batched.py tests/test_gpu_combined.py -q -m gpu -x -k "not full_size" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/memcheck.log
tail -5 $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_combined.py tests/test_gpu_nrt_batched.py -q -m gpu -x -k "fused or sequence or many_shapes" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/racecheck.log
tail -5 $O/racecheck.log
cap() {  # name regex skip count config
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$2" --launch-skip $3 -c $4 -f -o $O/$1 \
    python tools/measure_configs.py --configs $5 --steps 1 > $O/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
}
cap nrt2_expand nrt2_expand_kernel 2 1 c4
cap nrt2_table nrt2_table_kernel 4 2 c4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/c4_launches.csv python tools/measure_configs.py --configs c4 --steps 1 > $O/launches_run.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --kernel-only > $O/bench_launches_run.log 2>&1
timeout 600 python tools/measure_configs.py --out $O/configs.json > $O/configs.log 2>&1
tail -3 $O/configs.log | cut -c1-300
rm -f $O/*.ncu-rep

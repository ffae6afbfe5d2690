#!/bin/bash
# Round-end measurement refresh on ONE B200 (run under gpurun from the repo root); everything lands in gpurun_out/final
# and is copied into profiles/ afterwards.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
for cfg in c3 c4 c5; do
  timeout 400 python bench.py --config $cfg --steps 5 --warmup 3 > $O/bench_$cfg.json 2> $O/bench_$cfg.err
done
timeout 400 python tools/measure_configs.py --steps 5 --out $O/configs.json > $O/configs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --kernel-only > $O/launches_run.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file $O/c5_launches.csv \
  python bench.py --config c5 --steps 1 --warmup 1 --kernel-only > $O/c5_launches_run.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none \
  -k regex:'nrt2_q_kernel|nrt2_tableq_kernel|nrt2_expand_kernel|netoh_fast4_kernel|combine_top1_kernel' -c 8 -o $O/c5_kernels \
  python bench.py --config c5 --steps 1 --warmup 0 --kernel-only > $O/c5_kernels.log 2>&1
for k in nrt2_q_kernel nrt2_tableq_kernel nrt2_expand_kernel netoh_fast4_kernel combine_top1_kernel; do
  ncu -i $O/c5_kernels.ncu-rep --page raw --csv -k regex:$k > $O/r02b_${k}_raw.csv 2>/dev/null
done
rm -f $O/c5_kernels.ncu-rep
tail -2 $O/gpu_tests.log; tail -1 $O/smoke.log; head -c 600 $O/bench.json; echo; head -c 400 $O/bench_reference.json; echo

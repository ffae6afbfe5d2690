#!/bin/bash
# Round-end measurement refresh on ONE B200 (run under gpurun from the repo root); everything lands in gpurun_out/
# and is copied into profiles/ afterwards.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1
timeout 400 python bench.py > $O/bench.json 2> $O/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 500 python tools/measure_configs.py --out $O/configs.json > $O/configs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --kernel-only > $O/launches_run.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:nrt_kernel -c 1 -o $O/nrt_v4 \
  python tools/measure_configs.py --configs c4 > $O/nrt_v4.log 2>&1
tail -2 $O/gpu_tests.log; cat $O/smoke.log | tail -1; head -c 600 $O/bench.json; echo; head -c 400 $O/bench_reference.json; echo

# round-end confirmation on HEAD: full GPU suite + smoke, ncu captures of the P = 1 cycle kernels and the row expansion
set -u
O=gpurun_out/r27; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/gpu_tests.log 2>&1; echo "pytest rc=$?" >> $O/gpu_tests.log
tail -3 $O/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:'cycle_phase' -c 4 -o $O/cycle python tools/cycle_once.py 3 > $O/cycle.log 2>&1
ncu -i $O/cycle.ncu-rep --page raw --csv > $O/r02b_cycle_raw.csv 2>/dev/null
timeout 200 ncu --set full --import-source on --clock-control none -k regex:'expand_rows_kernel' -c 2 -o $O/rows python bench.py --config c3 --steps 1 --kernel-only > $O/rows.log 2>&1
ncu -i $O/rows.ncu-rep --page raw --csv > $O/r02b_expand_rows_raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
wc -c $O/*.csv

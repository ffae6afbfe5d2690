set -u
O=gpurun_out/r02g; mkdir -p $O
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:nrt2_expand_kernel" --launch-skip 2 -c 1 -f -o $O/nrt2_expand \
    python tools/measure_configs.py --configs c4 --steps 1 > $O/cap.log 2>&1
ncu -i $O/nrt2_expand.ncu-rep --page raw --csv > $O/nrt2_expand_raw.csv 2>/dev/null
ncu -i $O/nrt2_expand.ncu-rep --page source --csv > $O/nrt2_expand_source.csv 2>/dev/null
ls -la $O

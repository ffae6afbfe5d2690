#!/bin/bash
# ncu --set full capture of every P x N kernel that had no artefact in round 1 (VERDICT r01 "missing" #8), ONE GPU.
# Each capture runs tools/measure_configs.py for one config with --steps 1 and picks one warm launch of the kernel.
# Output: gpurun_out/r02_ncu/<name>.ncu-rep + <name>_raw.csv (summaries are copied to profiles/ afterwards).
set -u
O=gpurun_out/r02_ncu
mkdir -p $O
cap() {  # name config regex skip count
  timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$3" --launch-skip $4 -c $5 -f -o $O/$1 \
    python tools/measure_configs.py --configs $2 --steps 1 > $O/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
}
cap alloc_norm_u8 c2 alloc_norm_kernel 5 1
cap tlp c3 tlp_kernel 1 1
cap lvrb c3 lvrb_kernel 1 1
cap peaks c3b peaks_kernel 2 2
cap lowrisk c3b lowrisk_kernel 1 1
cap netoh_fast c5s netoh_fast_kernel 2 2
cap combine_topk c5s combine_topk_kernel 1 1
ls -la $O

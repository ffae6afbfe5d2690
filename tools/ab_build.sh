#!/bin/bash
# A/B build of libb200sched.so: recompiles ONE source with extra -D flags and links it with the other objects of
# the regular build.  usage: tools/ab_build.sh <name> <source.cu> "<extra nvcc flags>"  ->  scheduler-plugins_b200/lib/ab/<name>.so
set -e
cd "$(dirname "$0")/../scheduler-plugins_b200/csrc"
make -s >/dev/null
name=$1; src=$2; flags=$3
mkdir -p build/ab ../lib/ab
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC --expt-relaxed-constexpr \
  $flags -c $src -o build/ab/$name.o
objs=$(ls build/*.o | grep -v "build/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/ab/$name.so $objs build/ab/$name.o -cudart shared -ldl -Xlinker -rpath=/usr/local/cuda/lib64
echo "built lib/ab/$name.so"

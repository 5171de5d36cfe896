#!/bin/bash
# A/B helper: build ab/<name>.so = the default objects with the N = 128 kernel units (k_decode7s, k_sweep7s) recompiled with extra flags.
#   scripts/build_variant.sh t896 "-DPB_LIST_THREADS=896 -DPB_SWEEP_THREADS=896 -DPB_RETRY_THREADS=896"
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
obj=polar_code_b200/csrc/obj; out=ab/obj_$name
mkdir -p $out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177"
for u in k_decode7s k_sweep7s; do $NV $flags -c -o $out/$u.o polar_code_b200/csrc/$u.cu & done
wait
others=$(ls $obj/*.o | grep -v "k_decode7s.o\|k_sweep7s.o")
nvcc -shared -o ab/$name.so $others $out/k_decode7s.o $out/k_sweep7s.o
echo built ab/$name.so

#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..."  -> tools/bin/lib_NAME.so (tuning sweeps on the GPU box)
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin /tmp/mcov_var_$1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude"
for s in mcov_api.cu stats_sort.cu experimental.cu synth.cu bam_gpu.cu bamio.cpp transport.cpp; do
  nvcc $F $2 -c metacov_b200/csrc/$s -o /tmp/mcov_var_$1/${s%.*}.o &
done
wait
nvcc -shared -o tools/bin/lib_$1.so /tmp/mcov_var_$1/*.o -lz -gencode arch=compute_100a,code=sm_100a
echo built tools/bin/lib_$1.so

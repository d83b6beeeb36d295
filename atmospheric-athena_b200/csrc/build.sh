#!/bin/bash
# Builds libathena_b200.so for sm_100a IN-TREE (travels to the GPU box with the snapshot).
# --fmad=false: the reference's gcc -O3 x86-64 build has no FMA contraction; keeping products and sums
# separately rounded is what makes the hydro path bit-identical to it (DESIGN.md "Floating point").
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo ${AB2_FMAD:---fmad=false} -prec-div=true -prec-sqrt=true -std=c++17 -Xcompiler -fPIC,-O2 ${AB2_EXTRA_NVCC_FLAGS}"
pids=()
for f in hydro ion api units smr; do
  if [ ! -f $f.o ] || [ $f.cu -nt $f.o ] || [ physics.cuh -nt $f.o ] || [ engine.cuh -nt $f.o ] || [ glibc_math.cuh -nt $f.o ] || [ ../../include/athena_b200.h -nt $f.o ]; then
    rm -f $f.o
    $NVCC $FLAGS ${AB2_PTXAS_V:+-Xptxas -v} -c $f.cu -o $f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done     # set -e: a failed compile stops the build here
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libathena_b200.so hydro.o ion.o api.o units.o smr.o -lcudart
echo built $(pwd)/libathena_b200.so

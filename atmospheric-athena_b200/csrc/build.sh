#!/bin/bash
# Builds libathena_b200.so (the product) and libathena_b200_units.so (function-level test hooks) for sm_100a IN-TREE
# (they travel to the GPU box with the snapshot).
# --fmad=false: the reference's gcc -O3 x86-64 build has no FMA contraction; keeping products and sums
# separately rounded is what makes the hydro path bit-identical to it (DESIGN.md "Floating point").
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${AB2_OUT:-libathena_b200.so}          # tuning variants: AB2_OUT=libvariant.so AB2_OBJ=_v AB2_EXTRA_NVCC_FLAGS=-D...
SFX=${AB2_OBJ:-}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo ${AB2_FMAD:---fmad=false} -prec-div=true -prec-sqrt=true -std=c++17 -Xcompiler -fPIC,-O2 ${AB2_EXTRA_NVCC_FLAGS}"
pids=()
for f in hydro ion api units smr comm output; do
  [ -f $f.cu ] || continue
  o=$f$SFX.o
  if [ ! -f $o ] || [ $f.cu -nt $o ] || [ physics.cuh -nt $o ] || [ engine.cuh -nt $o ] || [ glibc_math.cuh -nt $o ] || [ ../../include/athena_b200.h -nt $o ] || [ -n "$SFX" ]; then
    rm -f $o
    $NVCC $FLAGS ${AB2_PTXAS_V:+-Xptxas -v} -c $f.cu -o $o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done     # set -e: a failed compile stops the build here
objs=""
for f in hydro ion api smr comm output; do [ -f $f$SFX.o ] && objs="$objs $f$SFX.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $objs -lcudart -ldl
if [ -z "$SFX" ]; then
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libathena_b200_units.so units.o -L. -lathena_b200 -Xlinker -rpath -Xlinker '$ORIGIN' -lcudart
fi
echo built $(pwd)/$OUT

#!/bin/bash
# scripts/build_prof_lib.sh -- builds build/prof/libjz_b200.so: the library with -DJZ_GEMM_PROFILE, whose tensor-core GEMM
# prints per-role cycle counts (TMA / MMA / transform / drain waits) of CTA 5 at the end of each launch.
# Use on the GPU box as  JZ_B200_LIB=build/prof/libjz_b200.so python scripts/gemm_shapes.py 8192,32,8192
set -e
cd "$(dirname "$0")/.."
OUT=${OUT:-build/prof}
mkdir -p $OUT
FLAGS="-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden ${PROF--DJZ_GEMM_PROFILE} $EXTRA"
pids=()
for f in juzhen_b200/csrc/*.cu; do
  o=$OUT/$(basename "${f%.cu}").o
  nvcc $FLAGS -c "$f" -o "$o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -shared -ccbin /usr/bin/g++ -o $OUT/libjz_b200.so $OUT/*.o -Xlinker --no-undefined
echo "built $OUT/libjz_b200.so"

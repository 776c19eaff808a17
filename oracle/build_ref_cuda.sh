#!/bin/bash
# oracle/build_ref_cuda.sh -- builds the UNMODIFIED reference CUDA backend (cuBLAS/cuRAND, its own kernels) for
# sm_100a from the sources where they lie under /root/reference, as the "existing GPU implementation" bar that
# SURVEY.md 8(d) asks to report next to ours on the same B200.  Test/measurement infrastructure only: outputs go
# to oracle/_ref/cuda/ (git-ignored, travels to the GPU box); nothing in the product links or loads them.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref/cuda
[ -d "$REF" ] || { echo "reference tree absent: keeping prebuilt $OUT if any"; exit 0; }
OB=$(ls /opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs/libopenblasp-r0-*.so | head -1)
mkdir -p "$OUT"
FLAGS="-std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a --extended-lambda -ccbin /usr/bin/g++ -DCUDA -DLOGGING_OFF -w \
  -DPROJECT_DIR=\"/root/repo/build/dropin/project\" -I$REF/external/OpenBLAS/include"
for unit in launcher cumatrix cukernels; do
  [ "$OUT/$unit.o" -nt "$REF/cpp/$unit.cu" ] || nvcc $FLAGS -c "$REF/cpp/$unit.cu" -o "$OUT/$unit.o" &
done
wait
for prog in demo_gemm demo_mnist demo_classification helloworld_nn knn; do
  [ "$OUT/$prog" -nt "$OUT/cumatrix.o" ] || nvcc $FLAGS "$REF/examples/$prog.cu" "$OUT/launcher.o" "$OUT/cumatrix.o" "$OUT/cukernels.o" \
      -lcublas -lcurand "$OB" -Xlinker --disable-new-dtags -Xlinker -rpath -Xlinker "$(dirname "$OB")" -lpthread -o "$OUT/$prog" &
done
wait
# the reference's operator benchmark (tests/benchmarkCoreOps.cu: GEMM 512^3, cuDNN conv, LayerNorm, attention block, Adam)
[ "$OUT/benchmarkCoreOps" -nt "$OUT/cumatrix.o" ] || nvcc $FLAGS -DCUDNN_AVAILABLE "$REF/tests/benchmarkCoreOps.cu" "$OUT/launcher.o" "$OUT/cumatrix.o" "$OUT/cukernels.o" \
    -lcublas -lcurand -lcudnn "$OB" -Xlinker --disable-new-dtags -Xlinker -rpath -Xlinker "$(dirname "$OB")" -lpthread -o "$OUT/benchmarkCoreOps"
# this repository's demo_mnist-step benchmark (juzhen_b200/cpp/tests/bench_mnist_step.cu: the loop body of examples/demo_mnist.cu
# at a batch size from the environment) against the reference's own CUDA sources: staged by symlink so that its
# `#include "../cpp/juzhen.hpp"` resolves into the reference tree
STAGE=$OUT/stage
mkdir -p "$STAGE/tests"
ln -sfn "$REF/cpp" "$STAGE/cpp"; ln -sfn "$REF/ml" "$STAGE/ml"; ln -sfn "$REF/external" "$STAGE/external"
ln -sfn "$HERE/../juzhen_b200/cpp/tests/bench_mnist_step.cu" "$STAGE/tests/bench_mnist_step.cu"
[ "$OUT/bench_mnist_step" -nt "$HERE/../juzhen_b200/cpp/tests/bench_mnist_step.cu" ] && [ "$OUT/bench_mnist_step" -nt "$OUT/cumatrix.o" ] || \
  nvcc $FLAGS "$STAGE/tests/bench_mnist_step.cu" "$OUT/launcher.o" "$OUT/cumatrix.o" "$OUT/cukernels.o" \
      -lcublas -lcurand "$OB" -Xlinker --disable-new-dtags -Xlinker -rpath -Xlinker "$(dirname "$OB")" -lpthread -o "$OUT/bench_mnist_step"
ls -la "$OUT" | grep -v "\.o$"

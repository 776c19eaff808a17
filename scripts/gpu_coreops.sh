#!/bin/bash
mkdir -p gpurun_out
cd build/dropin/project
{
echo "=== tests/benchmarkCoreOps.cu, reference CUDA build (cuBLAS/cuDNN/its own kernels)"
JUZHEN_BENCH_ITERS=50 timeout 300 ../../../oracle/_ref/cuda/benchmarkCoreOps 2>&1 | grep -E "^operation|GEMM|Conv2D|LayerNorm|Attention|Adam"
echo "=== tests/benchmarkCoreOps.cu, unchanged, on juzhen-b200"
JUZHEN_BENCH_ITERS=50 JZ_STATS=1 timeout 300 ../bin/benchmarkCoreOps 2>&1 | grep -E "^operation|GEMM|Conv2D|LayerNorm|Attention|Adam|jz_stats|rror"
} | tee ../../../gpurun_out/coreops.log

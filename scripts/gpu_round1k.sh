#!/bin/bash
# same-box comparison: reference CUDA build (cuBLAS fp32 / NVIDIA_TF32=1) vs this backend on its own demos
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
for p in demo_gemm demo_mnist; do
  echo "=== $p : reference CUDA build (cuBLAS fp32)"
  ( time timeout 900 ../../../oracle/_ref/cuda/$p ) 2>&1 | grep -E "TFLPOS|Duration|Rate: 0.0[0-9]* *$|real|ERROR|error" | tail -6
  echo "=== $p : reference CUDA build (NVIDIA_TF32=1)"
  ( time NVIDIA_TF32=1 timeout 900 ../../../oracle/_ref/cuda/$p ) 2>&1 | grep -E "TFLPOS|Duration|real|ERROR|error" | tail -4
  echo "=== $p : juzhen-b200 (3xTF32 default)"
  ( time JZ_STATS=1 timeout 900 ../bin/$p ) 2>&1 | grep -E "TFLPOS|Duration|jz_stats|real" | tail -6
  echo "=== $p : juzhen-b200 (NVIDIA_TF32=1)"
  ( time NVIDIA_TF32=1 JZ_STATS=1 timeout 900 ../bin/$p ) 2>&1 | grep -E "TFLPOS|Duration|jz_stats|real" | tail -6
done 2>&1 | tee ../../../gpurun_out/ref_cuda_vs_ours.log

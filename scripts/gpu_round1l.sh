#!/bin/bash
mkdir -p gpurun_out
echo "=== GEMM tests, TN forced 128"
JZ_GEMM_TN=128 timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -4
echo "=== GEMM tests, auto"
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
echo "=== sweep TN=256 forced (4096, 2048)"
JZ_GEMM_TN=256 timeout 600 python scripts/gemm_sweep.py 2048 4096 2>&1 | grep -v "^{" | grep "A\*B "
echo "=== sweep TN=128 forced (1024..8192)"
JZ_GEMM_TN=128 timeout 600 python scripts/gemm_sweep.py 1024 2048 4096 8192 2>&1 | grep -v "^{" | grep "A\*B "
echo "=== sweep auto"
timeout 900 python scripts/gemm_sweep.py 2>&1 | tee gpurun_out/gemm_sweep_l.log | grep -v "^{"

#!/bin/bash
# tile-width / chunk choices for the in-kernel-transform GEMM
mkdir -p gpurun_out
echo "=== GEMM tests"
timeout -k 5 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
for cfg in "JZ_GEMM_TN=256" "JZ_GEMM_TN=128"; do
  echo "=== sweep $cfg"
  env $cfg timeout -k 5 900 python scripts/gemm_sweep.py 2048 4096 8192 2>&1 | grep -v "^{" | grep "A\*B "
done
echo "=== sweep auto, all sizes"
timeout -k 5 900 python scripts/gemm_sweep.py 2>&1 | tee gpurun_out/gemm_sweep_o.log | grep -v "^{"

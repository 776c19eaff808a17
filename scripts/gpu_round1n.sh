#!/bin/bash
# GEMM without a pre-pass: operands in place in either major (MN-major = 128B swizzle / 32B atoms),
# lo parts of the 3xTF32 split computed in shared memory by the epilogue warps.
# Probe first (exact integer data), then parity, then sweep.
mkdir -p gpurun_out
for cfg in "" "JZ_GEMM_TN=128" "JZ_GEMM_CG=1"; do
  echo "=== probe 3xtf32 $cfg"
  env $cfg timeout -k 5 120 python scripts/gemm_probe.py 3xtf32 512 384 320 2>&1 | tail -7
done
echo "=== GEMM tests"
timeout -k 5 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6
echo "=== sweep auto (in-kernel transform, in-place operands)"
timeout -k 5 900 python scripts/gemm_sweep.py "$@" 2>&1 | tee gpurun_out/gemm_sweep_n.log | grep -v "^{"

#!/bin/bash
# GEMM bring-up / regression: probe (exact integer data) in every tile configuration, parity, sweep
mkdir -p gpurun_out
for cfg in "" "JZ_GEMM_TN=128" "JZ_GEMM_CG=1" "JZ_GEMM_PRESPLIT=1"; do
  for mode in 3xtf32 tf32; do
    echo "=== probe $mode $cfg"
    env $cfg timeout -k 5 120 python scripts/gemm_probe.py $mode 512 384 320 2>&1 | tail -7 | grep -v "ms/call"
  done
done
echo "=== probe many tiles per CTA (2304 x 2560 x 352: 90 tiles on 74 pair slots)"
timeout -k 5 120 python scripts/gemm_probe.py 3xtf32 2304 2560 352 2>&1 | tail -7
echo "=== GEMM tests"
timeout -k 5 900 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -6
echo "=== sweep"
timeout -k 5 900 python scripts/gemm_sweep.py "$@" 2>&1 | tee gpurun_out/gemm_sweep_n.log | grep -v "^{"

#!/bin/bash
mkdir -p gpurun_out
echo "=== parity"
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_gpu_e.log
echo "=== gemm sweep"
timeout 900 python scripts/gemm_sweep.py 2>&1 | tee gpurun_out/gemm_sweep.log | grep -v "^{" 
echo "=== compute-sanitizer memcheck (subset)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "not exhaustive and not large" 2>&1 | tail -15 | tee gpurun_out/sanitizer_memcheck.log
echo "=== ncu full (updated kernels)"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'chain_v4|map1_v4|transpose64|gemm_tcgen05|prep_' -c 20 \
    -f -o gpurun_out/prof_r01e python scripts/ncu_ops.py 28 4096 > gpurun_out/ncu_full_e.log 2>&1
tail -2 gpurun_out/ncu_full_e.log

#!/bin/bash
# compute-sanitizer over the kernels added or rewritten in this session (memcheck; racecheck on the shared-memory folds)
mkdir -p gpurun_out
echo "=== memcheck: attention helpers, batched / small GEMM, reductions (cluster row reduce), tensor GEMM at small sizes"
timeout -k 5 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_parity_gpu.py \
   -m gpu -q -x -k "not exhaustive and not large and not 2_to_32 and not 4096 and not 1700 and not 1024-2" 2>&1 | tail -8 | tee gpurun_out/sanitizer_memcheck_z.log
echo "=== racecheck: shared-memory folds (attention helpers, small GEMM k-split, row/col reductions)"
timeout -k 5 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_parity_gpu.py \
   -m gpu -q -x -k "(softmax_rows or layernorm or small_products or sum_max or reductions_vs) and not 1700 and not 1024-2 and not 4096-4096 and not 100003 and not 18" 2>&1 | tail -8 | tee gpurun_out/sanitizer_racecheck_z.log

#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -m gpu -q --tb=short -k "small or fuzz or batched or all_flags" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'gemm_small' --csv python scripts/ncu_small.py 2>/dev/null | grep gemm_small | awk -F'","' '{print $5, $(NF)}' | sed 's/void jz:://; s/(.*)//' 
bash scripts/gpu_mnist_time.sh 2>&1 | head -4

#!/bin/bash
mkdir -p gpurun_out
echo "=== small gemm bench (default routing)"
timeout -k 5 200 python scripts/small_gemm_bench.py 2>&1 | tail -10
echo "=== parity"
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_t.log
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
echo "=== test_fusion"
timeout 300 ../bin/test_fusion 2>&1 | grep -E "FAIL|launches|ALL PASSED|FAILED" | tail -40
stamp() { while IFS= read -r l; do echo "$(date +%s.%N) $l"; done; }
echo "=== demo_mnist: juzhen-b200"
JZ_STATS=1 timeout 900 ../bin/demo_mnist 2>&1 | stamp | grep -E "Rate|jz_stats" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s %s %s %s %s %s %s %s\n", $1-p, $2,$3,$4,$5,$6,$7,$8,$9,$10,$11} {p=$1}' | tail -5
echo "=== demo_classification / helloworld_nn timing: reference CUDA vs ours"
for p in demo_classification; do
  ( time timeout 600 ../../../oracle/_ref/cuda/$p > /dev/null 2>&1 ) 2>&1 | grep real
  ( time timeout 600 ../bin/$p > /dev/null 2>&1 ) 2>&1 | grep real
done

#!/bin/bash
# small-product GEMM kernel + deferred fills/draws + staged uploads: parity, then config 4 (demo_mnist) timing
mkdir -p gpurun_out
echo "=== parity"
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_s.log
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
echo "=== test_fusion"
timeout 300 ../bin/test_fusion 2>&1 | grep -E "PASS|FAIL|launches" | tail -40
stamp() { while IFS= read -r l; do echo "$(date +%s.%N) $l"; done; }
echo "=== demo_mnist: reference CUDA build (cuBLAS fp32)"
timeout 900 ../../../oracle/_ref/cuda/demo_mnist 2>&1 | stamp | grep "Rate" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s\n", $1-p, $2,$3,$4} {p=$1}' | tail -4
echo "=== demo_mnist: juzhen-b200"
JZ_STATS=1 timeout 900 ../bin/demo_mnist 2>&1 | stamp | grep -E "Rate|jz_stats" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s %s %s %s %s %s %s %s\n", $1-p, $2,$3,$4,$5,$6,$7,$8,$9,$10,$11} {p=$1}' | tail -5
echo "=== ncu launch list, this backend, launches 3000..4500"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1500 --csv \
   --log-file ../../../gpurun_out/mnist_launches_s.csv ../bin/demo_mnist > /dev/null 2>&1
wc -l ../../../gpurun_out/mnist_launches_s.csv

#!/bin/bash
# config 4 (demo_mnist, batch 32): wall time per 1000 steps for the reference CUDA build and this backend,
# and the kernel histogram of ~20 steady-state steps of this backend (ncu launch list)
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
stamp() { while IFS= read -r l; do echo "$(date +%s.%N) $l"; done; }
echo "=== reference CUDA build (cuBLAS fp32)"
timeout 900 ../../../oracle/_ref/cuda/demo_mnist 2>&1 | stamp | grep "Rate" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s\n", $1-p, $2,$3,$4} {p=$1}'
echo "=== juzhen-b200 (3xTF32 default)"
JZ_STATS=1 timeout 900 ../bin/demo_mnist 2>&1 | stamp | grep -E "Rate|jz_stats" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s %s %s %s %s %s\n", $1-p, $2,$3,$4,$5,$6,$7,$8,$9} {p=$1}'
echo "=== ncu launch list, this backend, launches 3000..4500"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 1500 --csv \
   --log-file ../../../gpurun_out/mnist_launches.csv ../bin/demo_mnist > /dev/null 2>&1
wc -l ../../../gpurun_out/mnist_launches.csv

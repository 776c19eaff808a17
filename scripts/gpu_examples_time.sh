#!/bin/bash
# whole-program wall time of the reference's examples: its own CUDA/cuBLAS build vs the same source on this backend
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
{
for p in demo_classification helloworld_nn knn demo_mnist; do
  for who in reference ours; do
    if [ $who = reference ]; then exe=../../../oracle/_ref/cuda/$p; else exe=../bin/$p; fi
    s=$(date +%s.%N); timeout 900 $exe > /tmp/out_$who.txt 2>&1; rc=$?; e=$(date +%s.%N)
    echo "$p $who rc=$rc wall=$(python3 -c "print(round($e - $s, 2))") s  $(grep -E 'k-nearest neighbour, Time|Misclassification Rate|training' /tmp/out_$who.txt | tail -1)"
  done
done
} | tee ../../../gpurun_out/examples_time.log

#!/bin/bash
mkdir -p gpurun_out
echo "=== size sweep"
timeout -k 5 900 python scripts/size_sweep.py 2>&1 | tee gpurun_out/size_sweep_v.log | grep -v "^{" | cut -c1-400
echo "=== reference CUDA build vs ours on its own demos"
bash scripts/gpu_round1k.sh 2>&1 | tail -40
cp gpurun_out/ref_cuda_vs_ours.log gpurun_out/ref_cuda_vs_ours_v.log

#!/bin/bash
# 2-GPU contact: multi-GPU parity test + torchrun bench with the sharded GEMM block
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_mg_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -40 | tee gpurun_out/mg2_test.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu --gemm-n 8192 2>gpurun_out/mg2_bench.err | tail -1 | tee gpurun_out/mg2_bench.json
tail -5 gpurun_out/mg2_bench.err

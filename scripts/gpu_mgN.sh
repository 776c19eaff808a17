#!/bin/bash
# N-GPU contact (N = $1): sharded-GEMM parity worker at world N + the torchrun bench exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "=== parity worker, world $N"
timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tests/_mg_gpu_worker.py 2>gpurun_out/mg${N}_worker.err | tail -40 | tee gpurun_out/mg${N}_parity.log
echo "=== bench, world $N"
timeout -k 5 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/mg${N}_bench.err | tail -1 | tee gpurun_out/mg${N}_bench.json | cut -c1-200
tail -5 gpurun_out/mg${N}_bench.err

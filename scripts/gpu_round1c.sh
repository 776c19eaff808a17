#!/bin/bash
# third contact: drop-in acceptance binaries + full parity suite + per-op bench after the chain/d_tanh fixes
mkdir -p gpurun_out
nvidia-smi -L
echo "=== parity + drop-in tests"
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -80 | tee gpurun_out/pytest_gpu_c.log
echo "=== drop-in outputs"
for p in demo_mnist knn; do
  echo "--- $p"; (cd build/dropin/project && timeout 600 ../bin/$p 2>&1 | tail -25)
done 2>&1 | tee gpurun_out/dropin_c.log
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_c.err | tail -1 | tee gpurun_out/bench_c.json

#!/bin/bash
mkdir -p gpurun_out
echo "=== parity (full)"
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_u.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench"
timeout -k 5 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_u.err | tail -1 > gpurun_out/bench_u.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_u.json').read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); print(d['gemm']); print(d['mnist_step']); print(d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/bench_u.err

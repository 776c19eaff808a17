#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_j.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_j.err | tail -1 | tee gpurun_out/bench_j.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); print({k:v['frac'] for k,v in d['ops'].items()}); print(d['gemm']); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline'])"

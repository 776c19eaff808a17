#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 "$@" 2>gpurun_out/bench_last.err | tail -1 | tee gpurun_out/bench_last.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); print({k:v['frac'] for k,v in d['ops'].items()}); print(d['gemm']); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'], d['gpu_launches'])"
tail -3 gpurun_out/bench_last.err

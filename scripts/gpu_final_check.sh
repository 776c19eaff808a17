#!/bin/bash
# last check of the tree as committed: full GPU test-suite, smoke(), one bench line
mkdir -p gpurun_out
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -k 5 900 python bench.py 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read()); print('value',d['value'],'frac',d['frac_of_hbm_peak'],'e2e',d['e2e']['value'],'launches',d['gpu_launches']); print({k:v['frac'] for k,v in d['ops'].items()}); print({k:v.get('TFLOP/s') for k,v in d['gemm'].items()}); print(d['mnist_step']['ours'], d['mnist_step']['speedup_vs_reference_cuda'])"

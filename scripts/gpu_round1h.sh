#!/bin/bash
mkdir -p gpurun_out
timeout 600 build/tools/tune_stream 2>&1 | grep -i "transpose\|memcpy" | tee gpurun_out/tune_stream2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_h.err | tail -1 | tee gpurun_out/bench_h.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); [print(k,v) for k,v in d['ops'].items()]; print(d['gemm'])"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -5

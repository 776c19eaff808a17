#!/bin/bash
# full validation after the no-pre-pass GEMM: parity, drop-in, bench, ncu launch list of the bench command,
# ncu --set full of every hot kernel (one launch each at bench size; GEMM at 4096 and 8192)
mkdir -p gpurun_out
echo "=== parity"
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_p.log
echo "=== bench"
timeout -k 5 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_p.err | tail -1 | tee gpurun_out/bench_p.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); print({k:v['frac'] for k,v in d['ops'].items()}); print(d['gemm']); print(d['roofline']); print(d['e2e']); print(d['cpu_baseline']); print(d['clocks'])"
echo "=== reference arm"
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_p_reference.json | cut -c1-400
echo "=== ncu launch list of the bench command"
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_p.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_p.log 2>&1
wc -l gpurun_out/ncu_launches_p.csv
echo "=== ncu full"
timeout -k 5 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'chain_v4|map1_v4|map2_v4|fill_v4|transpose64|gemm_tcgen05|colreduce|rowreduce|softmax' -c 24 \
    -f -o gpurun_out/prof_r01p python scripts/ncu_ops.py 28 4096 > gpurun_out/ncu_full_p.log 2>&1
tail -2 gpurun_out/ncu_full_p.log
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05' -c 2 \
    -f -o gpurun_out/prof_r01p_gemm8192 python scripts/ncu_ops.py 20 8192 > gpurun_out/ncu_full_p2.log 2>&1
tail -2 gpurun_out/ncu_full_p2.log
ls -la gpurun_out/*.ncu-rep

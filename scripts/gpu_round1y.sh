#!/bin/bash
# round-1 closing run: full parity, smoke, bench (+reference arm), ncu launch list of the bench command,
# ncu --set full of every hot kernel incl. the new ones (raw CSV exported on the box: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
echo "=== parity (full)"
timeout -k 5 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_final.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench"
timeout -k 5 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_final.err | tail -1 > gpurun_out/bench_final.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read()); print('value',d['value'],'frac',d['frac_of_hbm_peak']); print({k:v['frac'] for k,v in d['ops'].items()}); print(d['gemm']); print(d['mnist_step']); print(d['e2e']['value'], d['clocks'], d['gpu_launches'])"
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_final_reference.json
echo "=== ncu launch list of the bench command"
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-mnist > gpurun_out/bench_under_ncu_final.log 2>&1
wc -l gpurun_out/ncu_launches_final.csv
echo "=== ncu full: hot kernels"
timeout -k 5 1500 ncu --set full --clock-control none \
    -k regex:'chain_v4|map1_v4|map2_v4|fill_v4|transpose64|gemm_tcgen05|colreduce|rowreduce|softmax_reg' -c 24 \
    -f -o /tmp/prof_final python scripts/ncu_ops.py 28 4096 > gpurun_out/ncu_full_final.log 2>&1
tail -1 gpurun_out/ncu_full_final.log
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
echo "=== ncu full: small-product GEMM + transformer helpers"
timeout -k 5 900 ncu --set full --clock-control none -k regex:'gemm_small|softmax_rows|layernorm' -c 16 \
    -f -o /tmp/prof_final_small python scripts/ncu_small.py > gpurun_out/ncu_full_final2.log 2>&1
tail -1 gpurun_out/ncu_full_final2.log
ncu -i /tmp/prof_final_small.ncu-rep --page raw --csv > gpurun_out/prof_final_small_raw.csv 2>/dev/null
du -sh gpurun_out

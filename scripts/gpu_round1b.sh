#!/bin/bash
# second contact: parity suite + smoke + bench + ncu launch list + ncu full captures.
mkdir -p gpurun_out
nvidia-smi -L
echo "=== GEMM probe (two-level accumulation)"
for mode in tf32 3xtf32; do
  timeout 120 python scripts/gemm_probe.py $mode 4096 4096 4096 2>&1 | tail -7
done 2>&1 | tee gpurun_out/probe_4096.log
echo "=== parity tests"
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench.err | tail -1 | tee gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/bench_reference.json
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --gemm-n 4096 > gpurun_out/bench_under_ncu.log 2>&1
echo "=== ncu full"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'chain_v4|softmax_reg|transpose64|map1_v4|map2_v4|rowreduce|colreduce_warp|gemm_tcgen05|prep_' -c 40 \
    -f -o gpurun_out/prof_r01b python scripts/ncu_ops.py 28 4096 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out

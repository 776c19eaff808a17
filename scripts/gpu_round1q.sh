#!/bin/bash
# bench + ncu evidence, sized to come back through gpurun_out (64 MiB cap): raw CSV pages are produced on the box
mkdir -p gpurun_out
echo "=== bench"
timeout -k 5 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_q.err | tail -1 > gpurun_out/bench_q.json
cut -c1-300 gpurun_out/bench_q.json
timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_q_reference.json
echo "=== ncu launch list of the bench command"
timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_q.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu_q.log 2>&1
wc -l gpurun_out/ncu_launches_q.csv
echo "=== ncu full: all hot kernels, one launch each (no source import: small report)"
timeout -k 5 1500 ncu --set full --clock-control none \
    -k regex:'chain_v4|map1_v4|map2_v4|fill_v4|transpose64|gemm_tcgen05|colreduce|rowreduce|softmax' -c 24 \
    -f -o /tmp/prof_r01q python scripts/ncu_ops.py 28 4096 > gpurun_out/ncu_full_q.log 2>&1
tail -1 gpurun_out/ncu_full_q.log
ncu -i /tmp/prof_r01q.ncu-rep --page raw --csv > gpurun_out/prof_r01q_raw.csv 2>/dev/null
echo "=== ncu full with source: GEMM 8192 3xTF32 (1 launch)"
timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05' -c 1 \
    -f -o /tmp/prof_r01q_gemm8192 python scripts/ncu_ops.py 20 8192 > gpurun_out/ncu_full_q2.log 2>&1
tail -1 gpurun_out/ncu_full_q2.log
ncu -i /tmp/prof_r01q_gemm8192.ncu-rep --page raw --csv > gpurun_out/prof_r01q_gemm8192_raw.csv 2>/dev/null
ncu -i /tmp/prof_r01q_gemm8192.ncu-rep --page details > gpurun_out/prof_r01q_gemm8192_details.txt 2>/dev/null
ls -la /tmp/*.ncu-rep
for f in /tmp/prof_r01q_gemm8192.ncu-rep /tmp/prof_r01q.ncu-rep; do
  sz=$(du -sm gpurun_out | cut -f1); fs=$(du -sm $f | cut -f1)
  if [ $((sz + fs)) -lt 58 ]; then cp $f gpurun_out/; fi
done
du -sh gpurun_out

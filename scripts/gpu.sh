#!/bin/bash
# scripts/gpu.sh <stage> [<stage> ...] -- the one parameterised GPU-box script (gpurun -- 'bash scripts/gpu.sh gemmtests sweep').
# Every stage writes its log under gpurun_out/ (copied back by gpurun); copy what should be judged into profiles/.
#   gemmtests     pytest of the GEMM / attention tests (fast fail before anything expensive)
#   tests         the full GPU suite (pytest -m gpu)
#   smoke         __graft_entry__.smoke()
#   sweep[:sizes] GEMM sweep 1024..16384 x 3 layouts x 2 modes (scripts/gemm_sweep.py), e.g. sweep:1024,2048
#   shapes        small squares, skinny products, demo_mnist products at large batch vs cuBLAS (scripts/gemm_shapes.py)
#   bench         python bench.py (+ --impl reference)
#   attention     bench_attention (helper kernels + strided-batched GEMM vs the reference's kernels / cuBLAS)
#   mnist         bench_mnist_step at batch 32 / 8192 / 60000, this backend and the reference's CUDA build
#   softmax       long-column softmax, with and without the next-column prefetch
#   dropin        the reference's own programs + test_fusion against this backend (tests/test_dropin_gpu.py)
#   ncu_config1   source-level ncu of the GEMM with the config-1 chain in its epilogue
#   ncu_gemm      ncu --set full of the tensor-core GEMM at 1024 / 4096 / 8192 (both modes)
#   ncu_launches  ncu launch list of the bench command
#   sanitizer     compute-sanitizer memcheck + racecheck over the GPU tests
#   mg:N          N-GPU parity worker + torchrun bench (call gpurun with --gpus N)
#   mgquick:N     parity worker (all gather modes) + C++ sharded-dot test + a short torchrun bench with the sharded GEMM block
#   mgcpp:N       the C++ multi-process sharded-dot test over the jz_mg_* ABI
TAG=${JZ_TAG:-r02}
mkdir -p gpurun_out
nvidia-smi -L | head -8
for stage in "$@"; do
  arg=""; case "$stage" in *:*) arg="${stage#*:}"; stage="${stage%%:*}";; esac
  echo "================ stage $stage $arg"
  case "$stage" in
    gemmtests)
      timeout -k 5 1500 python -m pytest tests/test_gemm_gpu.py tests/test_attention_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_gemm.log ;;
    tests)
      timeout -k 5 3000 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log ;;
    sweep)
      timeout -k 5 1200 python scripts/gemm_sweep.py ${arg//,/ } 2>&1 | tee gpurun_out/${TAG}_gemm_sweep.log | grep -v '^{' ;;
    shapes)
      timeout -k 5 900 python scripts/gemm_shapes.py 2>&1 | tee gpurun_out/${TAG}_gemm_shapes.log ;;
    prof)
      JZ_B200_LIB=build/prof/libjz_b200.so timeout -k 5 300 python scripts/gemm_prof_one.py 8192,32,8192 8192,32,8192,1,0 128,1024,60000,0,1 \
        1024,1024,1024 1536,1536,1536 2048,2048,2048,0,0,1 4096,4096,4096 2>&1 | grep -v '^$' | tee gpurun_out/${TAG}_gemm_prof.log ;;
    bench)
      timeout -k 5 1500 python bench.py --impl reference 2>gpurun_out/${TAG}_bench_reference.err | tail -1 > gpurun_out/${TAG}_bench_reference.json
      timeout -k 5 1500 python bench.py 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench.json
      cut -c1-600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json ;;
    attention)
      timeout -k 5 600 build/dropin/bin/bench_attention 2>&1 | grep -v "^profiler\|^GPU\|^Juzhen\|^___\|^GEMM mode" | tee gpurun_out/${TAG}_attention.log ;;
    mnist)
      for b in 32 8192 60000; do
        echo "--- batch $b: juzhen-b200"; MNIST_BATCH=$b JZ_STATS=1 timeout -k 5 300 build/dropin/bin/bench_mnist_step 2>&1 | grep -E "bench_mnist_step|jz_stats"
        echo "--- batch $b: reference CUDA/cuBLAS build"; MNIST_BATCH=$b timeout -k 5 300 oracle/_ref/cuda/bench_mnist_step 2>&1 | grep -E "bench_mnist_step"
        echo "--- batch $b: reference CUDA/cuBLAS build, NVIDIA_TF32=1"; NVIDIA_TF32=1 MNIST_BATCH=$b timeout -k 5 300 oracle/_ref/cuda/bench_mnist_step 2>&1 | grep -E "bench_mnist_step"
      done 2>&1 | tee gpurun_out/${TAG}_mnist_step.log ;;
    softmax)
      { timeout -k 5 300 python scripts/long_softmax_bench.py; JZ_SOFTMAX_NO_PREFETCH=1 timeout -k 5 300 python scripts/long_softmax_bench.py; } 2>&1 | tee gpurun_out/${TAG}_long_column_softmax.log ;;
    dropin)
      timeout -k 5 1500 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest_dropin.log ;;
    ncu_config1)
      # source-level profile of the GEMM with the config-1 chain fused into its epilogue (4096^3, 3xTF32)
      timeout -k 5 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 2 -c 1 -o gpurun_out/${TAG}_config1 -f \
          python scripts/config1_probe.py > gpurun_out/${TAG}_ncu_config1.log 2>&1
      tail -3 gpurun_out/${TAG}_ncu_config1.log
      ncu -i gpurun_out/${TAG}_config1.ncu-rep --page source --csv > gpurun_out/${TAG}_config1_source.csv 2>/dev/null
      ncu -i gpurun_out/${TAG}_config1.ncu-rep --page raw --csv > gpurun_out/${TAG}_config1_raw.csv 2>/dev/null; wc -c gpurun_out/${TAG}_config1_*.csv
      [ $(stat -c %s gpurun_out/${TAG}_config1.ncu-rep) -gt 20000000 ] && rm -f gpurun_out/${TAG}_config1.ncu-rep ;;
    ncu_gemm)
      timeout -k 5 900 ncu --set full --clock-control none -k regex:gemm_tcgen05 -c 12 -o gpurun_out/${TAG}_gemm -f \
          python scripts/ncu_gemm_probe.py ${arg//,/ } > gpurun_out/${TAG}_ncu_gemm.log 2>&1
      tail -3 gpurun_out/${TAG}_ncu_gemm.log
      ncu -i gpurun_out/${TAG}_gemm.ncu-rep --page raw --csv > gpurun_out/${TAG}_gemm_raw.csv 2>/dev/null; wc -c gpurun_out/${TAG}_gemm_raw.csv
      # gpurun copies back at most 64 MiB: keep the report only when it is small
      [ $(stat -c %s gpurun_out/${TAG}_gemm.ncu-rep) -gt 20000000 ] && rm -f gpurun_out/${TAG}_gemm.ncu-rep ;;
    ncu_launches)
      timeout -k 5 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/${TAG}_ncu_launches.csv \
          python bench.py --steps 2 --warmup 1 --no-cpu --no-mnist > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
      tail -2 gpurun_out/${TAG}_ncu_launches.csv | cut -c1-300 ;;
    sanitizer)
      timeout -k 5 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_parity_gpu.py \
         -m gpu -q -x -k "not exhaustive and not large and not 2_to_32 and not 4096 and not 1700 and not 1024-2 and not wide_output and not 8192 and not 60000" 2>&1 | tail -8 | tee gpurun_out/${TAG}_sanitizer_memcheck.log
      timeout -k 5 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_attention_gpu.py tests/test_gemm_gpu.py tests/test_parity_gpu.py \
         -m gpu -q -x -k "(softmax_rows or layernorm or small_products or sum_max or reductions_vs or adam) and not 1700 and not 1024-2 and not 4096-4096 and not 100003 and not 18" 2>&1 | tail -8 | tee gpurun_out/${TAG}_sanitizer_racecheck.log ;;
    mg)
      N=${arg:-2}
      timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
          tests/_mg_gpu_worker.py 2>gpurun_out/${TAG}_mg${N}_worker.err | tail -40 | tee gpurun_out/${TAG}_mg${N}_parity.log
      timeout -k 5 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench_${N}gpu.err | tail -1 | tee gpurun_out/${TAG}_bench_${N}gpu.json | cut -c1-300
      tail -5 gpurun_out/${TAG}_bench_${N}gpu.err ;;
    mgquick)
      N=${arg:-2}
      timeout -k 5 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
          tests/_mg_gpu_worker.py 2>gpurun_out/${TAG}_mg${N}_worker.err | tail -40 | tee gpurun_out/${TAG}_mg${N}_parity.log
      tail -5 gpurun_out/${TAG}_mg${N}_worker.err
      timeout -k 5 300 build/dropin/bin/test_mg_dot $N 2>&1 | tail -20 | tee gpurun_out/${TAG}_mgcpp${N}.log
      timeout -k 5 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $N --steps 6 --warmup 3 --no-mnist --gemm-n 4096 2>gpurun_out/${TAG}_benchq_${N}gpu.err | tail -1 | tee gpurun_out/${TAG}_benchq_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'e2e',d['e2e']['value']); print(json.dumps(d['sharded_gemm'],indent=1)); print(d['sharded_colsum'])"
      tail -5 gpurun_out/${TAG}_benchq_${N}gpu.err ;;
    mgcpp)
      N=${arg:-2}
      timeout -k 5 300 build/dropin/bin/test_mg_dot $N 2>&1 | tail -20 | tee gpurun_out/${TAG}_mgcpp${N}.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done

#!/bin/bash
# first contact with the B200: probes -> parity -> bench.  Everything logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))"
echo "=== GEMM probes"
for cg in 1 2; do for mode in tf32 3xtf32; do
  echo "--- CG=$cg mode=$mode 256^3"; JZ_GEMM_CG=$cg timeout 90 python scripts/gemm_probe.py $mode 256 256 256 2>&1 | tail -25
done; done 2>&1 | tee gpurun_out/probe_small.log
for cg in 1 2; do
  echo "--- CG=$cg 3xtf32 1000x520x777"; JZ_GEMM_CG=$cg timeout 90 python scripts/gemm_probe.py 3xtf32 1000 520 777 2>&1 | tail -12
done 2>&1 | tee -a gpurun_out/probe_small.log
for cg in 1 2; do for mode in tf32 3xtf32; do
  echo "--- CG=$cg mode=$mode 4096^3"; JZ_GEMM_CG=$cg timeout 120 python scripts/gemm_probe.py $mode 4096 4096 4096 2>&1 | tail -8
done; done 2>&1 | tee gpurun_out/probe_4096.log
echo "=== parity tests (non-GEMM)"
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -80 | tee gpurun_out/parity.log
echo "=== GEMM tests"
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -60 | tee gpurun_out/gemm_tests.log
echo "=== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-log2n 22 2>&1 | tail -3 | tee gpurun_out/bench_first.json

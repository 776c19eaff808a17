#!/bin/bash
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
echo "=== fusion test (lazy)"
(cd build/dropin/project && timeout 300 ../bin/test_fusion 2>&1 | tail -40) | tee gpurun_out/fusion_lazy.log
echo "=== fusion test (eager)"
(cd build/dropin/project && JZ_EAGER=1 timeout 300 ../bin/test_fusion 2>&1 | grep -E "launches|FAIL|PASSED|FAILED") | tee gpurun_out/fusion_eager.log
echo "=== drop-in + parity tests"
timeout 2400 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_f.log
echo "=== demo timings (lazy vs eager)"
for p in demo_gemm demo_mnist; do
  for e in 0 1; do
    echo "--- $p JZ_EAGER=$e"; (cd build/dropin/project && JZ_EAGER=$e bash -c "time timeout 900 ../bin/$p" 2>&1 | grep -E "TFLPOS|Misclassification Rate: 0.05|real|Duration" | head -6)
  done
done 2>&1 | tee gpurun_out/demo_timings.log

#!/bin/bash
mkdir -p gpurun_out
python juzhen_b200/cpp/build_dropin.py --extract-datasets
echo "=== stream / transpose design probe"
timeout 600 build/tools/tune_stream 2>&1 | tee gpurun_out/tune_stream.log
echo "=== drop-in fusion test"
timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q --tb=short -k "fusion or testbasic" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_g.log

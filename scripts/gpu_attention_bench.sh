#!/bin/bash
mkdir -p gpurun_out
cd build/dropin/project
timeout 600 ../bin/bench_attention 2>&1 | grep -E "softmax|layernorm" | tee ../../../gpurun_out/attention_bench.log

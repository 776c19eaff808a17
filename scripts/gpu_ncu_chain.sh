#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain_v4|UnaryF<\(int\)1>|UnaryF<1>' -c 3 \
    -f -o gpurun_out/prof_chain python scripts/ncu_ops.py 28 1024 > gpurun_out/ncu_chain.log 2>&1
tail -2 gpurun_out/ncu_chain.log

#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --cache-control none -k regex:'gemm_small' -c 1 -f -o /tmp/prof_small_src python scripts/ncu_small.py > /dev/null 2>&1
ncu -i /tmp/prof_small_src.ncu-rep --page source --csv > gpurun_out/small_source.csv 2>/dev/null
ncu -i /tmp/prof_small_src.ncu-rep --page details 2>/dev/null | grep -E "Duration|Stall|stall|Issued Warp|Eligible|No Eligible|One or More|L1/TEX Hit|L2 Hit|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc|Mem Busy|Max Bandwidth" | head -30
wc -l gpurun_out/small_source.csv

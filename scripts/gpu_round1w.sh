#!/bin/bash
for cfg in "JZ_REDUCE_WAVES=1" "JZ_REDUCE_WAVES=4" "JZ_REDUCE_WAVES=1 JZ_REDUCE_NO_CLUSTER=1"; do
echo "=== $cfg"
env $cfg timeout -k 5 900 python scripts/size_sweep.py 22 24 26 28 2>&1 | grep -v "^{" | sed 's/ /\n/g' | grep -E "^2\^|sum_dim1" | paste -sd' ' 
done

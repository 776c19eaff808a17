#!/bin/bash
python juzhen_b200/cpp/build_dropin.py --extract-datasets
cd build/dropin/project
stamp() { while IFS= read -r l; do echo "$(date +%s.%N) $l"; done; }
for cfg in "" "JZ_SMALL_WARPS=16" "JZ_SMALL_WARPS=8"; do
echo "=== demo_mnist: juzhen-b200 $cfg"
env $cfg JZ_STATS=1 timeout 900 ../bin/demo_mnist 2>&1 | stamp | grep -E "Rate|jz_stats" | awk 'NR>1{printf "%.3f s per 1000 steps  %s %s %s %s %s %s %s %s %s %s\n", $1-p, $2,$3,$4,$5,$6,$7,$8,$9,$10,$11} {p=$1}' | tail -3
done

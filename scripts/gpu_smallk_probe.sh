#!/bin/bash
# weight-gradient products (k = batch = 32; shapes 4 and 5 of small_gemm_bench.py): tensor path, small-product kernel and the
# shared-memory SIMT kernel, warm-cache kernel durations under ncu
for cfg in "" "JZ_GEMM_FORCE_SIMT=1" "JZ_GEMM_FORCE_SIMT=1 JZ_GEMM_NO_SMALL=1"; do
  echo "=== $cfg"
  env $cfg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'gemm_tcgen05|sgemm_simt|gemm_small' --launch-skip 950 -c 500 --csv python scripts/small_gemm_bench.py 2>/dev/null \
   | grep -E "gemm_tcgen05|sgemm_simt|gemm_small" | awk -F'","' '{k=$5; sub(/\(.*/,"",k); g=$9; d=$NF; gsub(/"/,"",d); s[k" grid "g]+=d; n[k" grid "g]++} END{for (x in s) printf "%-70s n=%3d avg %.2f us\n", x, n[x], s[x]/n[x]/1000}' | sort
done

#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 6 --warmup 3 --no-cpu --no-gemm 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',d['value'],'frac',d['frac_of_hbm_peak'], {k:v['frac'] for k,v in d['ops'].items() if k in ('log','chain_softplus5','sum_dim0','sum_dim1','max_dim0','softmax_cols','transpose')})"; }
echo "default (uncapped warp kernel, 2 waves rows)"; run
for w in 4 8 16; do echo "JZ_REDUCE_WAVES=$w"; JZ_REDUCE_WAVES=$w run; done
echo "JZ_REDUCE_CAP=8 (old)"; JZ_REDUCE_CAP=8 run
echo "JZ_REDUCE_CAP=16"; JZ_REDUCE_CAP=16 run
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -x -s -k "log or chain or reduc or sum or softmax or exhaustive" 2>&1 | grep -E "ulp sweep|passed|failed" | tail -12

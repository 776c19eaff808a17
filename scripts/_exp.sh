timeout -k 5 300 python bench.py --no-gemm --no-cpu --no-mnist --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',d['value'], 'frac', d['frac_of_hbm_peak'])
for k,v in d['ops'].items(): print(f'{k:18s} {v}')"
timeout -k 5 900 python -m pytest tests/test_parity_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "packed or chain or log or unary or fused or epilogue or exhaustive" 2>&1 | grep -v "^profiler" | tail -8

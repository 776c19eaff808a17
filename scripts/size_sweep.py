"""BASELINE configs[1] across its whole size range: every op of the bench sweep at 2^20 .. 2^30 fp32 elements on one
B200, CUDA events, GB/s of ALGORITHMIC bytes and fraction of the measured copy peak.  Sizes at or below
the 126 MB L2 run out of L2 (no flush: that is what a caller sees in a loop) and are marked."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
peak = bench.peaks()["hbm_gbs"]
sizes = [int(a) for a in sys.argv[1:]] or [20, 22, 24, 26, 28, 30]
out = {}
for log2n in sizes:
    rows = cols = 1 << (log2n // 2)
    if log2n % 2:
        cols *= 2
    sw = bench.Sweep(jz, rows, cols, stream)
    res = {}
    for name, bpe in bench.SWEEP:
        fn = sw.ops[name]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 200 if log2n <= 22 else (50 if log2n <= 26 else 10)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name] = round(bpe * sw.n / (ms * 1e-3) / 1e9, 1)
    out[f"2^{log2n}"] = res
    tag = "L2-resident" if 3 * 4 * sw.n <= 126e6 else "HBM"
    print(f"2^{log2n} ({rows}x{cols}, {tag}): " + " ".join(f"{k}={v:.0f}({v/peak:.2f})" for k, v in res.items()), flush=True)
    del sw
    torch.cuda.empty_cache()
    L.jz_pool_trim()
print(json.dumps({"size_sweep_GBs": out, "peak_GBs": peak}))

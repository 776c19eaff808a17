"""column softmax over very long columns (rows > 32768): cluster kernel vs the three-pass fallback (JZ_SOFTMAX_NO_CLUSTER=1)"""
import os, sys, torch
sys.path.insert(0, ".")
import juzhen_b200 as jz
L = jz.lib(); assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream; jz.set_stream(stream)
tag = "three-pass" if os.environ.get("JZ_SOFTMAX_NO_CLUSTER") else "cluster"
for rows, cols in ((65536, 4096), (131072, 2048), (262144, 1024)):
    x, y = jz.CM.randn(rows, cols, seed=1), jz.CM.empty("y", rows, cols)
    f = lambda: L.jz_softmax_cols(y.ptr, x.ptr, rows, cols, rows, stream)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"[{tag}] softmax_cols {rows} x {cols}: {ms:.3f} ms  {8.0*rows*cols/ms/1e6:.0f} GB/s")

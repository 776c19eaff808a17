"""Large but skinny products (one dimension < 64, more than 2^26 multiply-adds): tensor path vs the fp32 SIMT kernel
(JZ_GEMM_FORCE_SIMT=1).  CUDA events."""
import os
import sys

import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
tag = "simt" if os.environ.get("JZ_GEMM_FORCE_SIMT") else "default"
for (m, n, k) in ((4096, 48, 4096), (8192, 32, 8192), (32, 8192, 8192), (16384, 16, 16384), (1024, 10000, 784)):
    a, b, c = jz.CM.randn(m, k, seed=1), jz.CM.randn(k, n, seed=2), jz.CM.empty("c", m, n)
    def run():
        assert L.jz_gemm(0, 0, m, n, k, 1.0, a.ptr, m, b.ptr, k, 0.0, c.ptr, m, 0, stream) == 0, L.jz_last_error()
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"[{tag}] {m}x{n}x{k}: path={L.jz_gemm_last_path()} {ms*1e3:8.1f} us  {2.0*m*n*k/ms/1e9:7.1f} TFLOP/s")

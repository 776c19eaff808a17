"""4096^3 3xTF32 GEMM with log(exp(x/4096)+1)/5 fused into the epilogue, a few launches (for ncu) + timing of fused vs separate"""
import sys, torch
import numpy as np
sys.path.insert(0, ".")
import juzhen_b200 as jz
L = jz.lib(); assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream; jz.set_stream(stream)
n = 4096
a, b, c = jz.CM.randn(n, n, seed=1), jz.CM.randn(n, n, seed=2), jz.CM.empty("c", n, n)
steps, ns = jz._lib.make_steps([("affine", float(np.float32(1.0 / n)), 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", 0.2, 0.0)])
def fused(): assert L.jz_gemm_chain(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, c.ptr, n, steps, ns, 0, stream) == 0
def plain(): assert L.jz_gemm(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, c.ptr, n, 0, stream) == 0
for name, f in (("fused", fused), ("plain", plain)):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms")

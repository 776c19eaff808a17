"""A launch-bound step (one demo_mnist-shaped layer forward + softmax + three reductions + a 512^3 product: 8 jz_* calls)
issued call by call through the C ABI, and replayed as ONE CUDA graph captured from the same calls."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
U = jz._lib.UNARY
m, k, n = 1024, 784, 32
rng = np.random.default_rng(0)
W = jz.CM(np.asfortranarray(rng.standard_normal((m, k)).astype(np.float32) * 0.05))
X = jz.CM(np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32)))
b = jz.CM(np.asfortranarray(rng.standard_normal((m, 1)).astype(np.float32)))
A, B = jz.CM.randn(512, 512, seed=1), jz.CM.randn(512, 512, seed=2)
H, S, rs, cs, P = jz.CM.empty("h", m, n), jz.CM.empty("s", m, n), jz.CM.empty("rs", m, 1), jz.CM.empty("cs", n, 1), jz.CM.empty("p", 512, 512)
side = torch.cuda.Stream()
s = side.cuda_stream


def step():
    L.jz_gemm(0, 0, m, n, k, 1.0, W.ptr, m, X.ptr, k, 0.0, H.ptr, m, 0, s)
    L.jz_add_bcast(H.ptr, H.ptr, m, n, b.ptr, 1, 1.0, 1.0, s)
    L.jz_unary(U["tanh"], H.ptr, H.ptr, m * n, s)
    L.jz_softmax_cols(S.ptr, H.ptr, m, n, m, s)
    L.jz_sum(rs.ptr, S.ptr, m, n, m, 1, s)
    L.jz_sum(cs.ptr, S.ptr, m, n, m, 0, s)
    L.jz_unary(U["exp"], S.ptr, S.ptr, m * n, s)
    L.jz_gemm(0, 1, 512, 512, 512, 1.0, A.ptr, 512, B.ptr, 512, 0.0, P.ptr, 512, 0, s)


with torch.cuda.stream(side):
    for _ in range(20):
        step()
side.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    step()
torch.cuda.synchronize()
for name, fn in (("call by call (ctypes)", step), ("one CUDA graph replay", g.replay)):
    with torch.cuda.stream(side):
        for _ in range(20):
            fn()
        side.synchronize()
        t0 = time.perf_counter()
        for _ in range(2000):
            fn()
        side.synchronize()
    print(f"{name:26s} {(time.perf_counter() - t0) / 2000 * 1e6:8.1f} us per 8-kernel step")

"""One launch of the latency-bound and 'next-row' kernels for ncu captures: the products of a demo_mnist step through
the small-product GEMM kernel, batched row softmax (+backward), LayerNorm forward/backward."""
import sys

import numpy as np

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
CM = jz.CM
for (ta, tb, m, n, k) in ((0, 0, 1024, 32, 784), (0, 0, 128, 32, 1024), (1, 0, 784, 32, 1024), (0, 0, 10, 32, 128)):
    a = CM.randn(k if ta else m, m if ta else k, seed=1)
    b = CM.randn(n if tb else k, k if tb else n, seed=2)
    c = CM.empty("c", m, n)
    assert L.jz_gemm(ta, tb, m, n, k, 1.0, a.ptr, a.numrow, b.ptr, b.numrow, 0.0, c.ptr, m, 0, None) == 0
S, B = 256, 256
x, y, d, g = CM.randn(S * S * B, 1, seed=3), CM.empty("y", S * S * B, 1), CM.randn(S * S * B, 1, seed=4), CM.empty("g", S * S * B, 1)
assert L.jz_softmax_rows_batched(y.ptr, x.ptr, S, B, 1, -1e9, None) == 0
assert L.jz_softmax_rows_backward(g.ptr, y.ptr, d.ptr, S, B, 0.125, None) == 0
D, N = 1024, 16384
x, ga, be = CM.randn(D, N, seed=5), CM.randn(D, 1, seed=6), CM.randn(D, 1, seed=7)
yy, xh, inv, dx = CM.empty("y", D, N), CM.empty("xh", D, N), CM.empty("inv", N, 1), CM.empty("dx", D, N)
assert L.jz_layernorm_forward(yy.ptr, xh.ptr, inv.ptr, x.ptr, ga.ptr, be.ptr, D, N, None) == 0
assert L.jz_layernorm_backward(dx.ptr, yy.ptr, ga.ptr, xh.ptr, inv.ptr, D, N, None) == 0
jz.sync()
print("ncu_small done")

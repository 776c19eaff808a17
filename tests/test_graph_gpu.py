"""CUDA-graph capture of C-ABI calls (launch-bound loops: the B200 answer is stream capture, not a tracing compiler).
After one warm-up pass (so the pool, the ticket counters and the kernel attributes exist) a sequence of jz_* calls on a
capturing stream must record into a graph -- no allocation, no synchronisation, no legacy-stream work inside -- and every
replay must reproduce the eager results bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_abi_calls_capture_into_a_cuda_graph_and_replay_bit_exact(jz):
    import torch
    L = jz.lib()
    U = jz._lib.UNARY
    m, k, n = 1024, 784, 32            # a demo_mnist layer at batch 32
    rng = np.random.default_rng(0)
    W = jz.CM(np.asfortranarray(rng.standard_normal((m, k)).astype(np.float32) * 0.05))
    X = jz.CM(np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32)))
    b = jz.CM(np.asfortranarray(rng.standard_normal((m, 1)).astype(np.float32)))
    big_a, big_b = jz.CM.randn(512, 512, seed=1), jz.CM.randn(512, 512, seed=2)
    H, S, rs, cs, P = (jz.CM.empty("h", m, n), jz.CM.empty("s", m, n), jz.CM.empty("rs", m, 1), jz.CM.empty("cs", n, 1),
                       jz.CM.empty("p", 512, 512))
    wide = jz.CM.randn(256, 4096, seed=3)
    wsum = jz.CM.empty("ws", 256, 1)
    side = torch.cuda.Stream()
    s = side.cuda_stream

    def step():
        chk = jz._lib.check
        chk(L.jz_gemm(0, 0, m, n, k, 1.0, W.ptr, m, X.ptr, k, 0.0, H.ptr, m, 0, s))                    # small-product kernel
        chk(L.jz_add_bcast(H.ptr, H.ptr, m, n, b.ptr, 1, 1.0, 1.0, s))                                  # + b * ones(1, N)
        chk(L.jz_unary(U["tanh"], H.ptr, H.ptr, m * n, s))
        chk(L.jz_softmax_cols(S.ptr, H.ptr, m, n, m, s))
        chk(L.jz_sum(rs.ptr, S.ptr, m, n, m, 1, s))
        chk(L.jz_sum(cs.ptr, S.ptr, m, n, m, 0, s))
        chk(L.jz_sum(wsum.ptr, wide.ptr, 256, 4096, 256, 1, s))                                         # cluster + ticket row reduce
        chk(L.jz_gemm(0, 1, 512, 512, 512, 1.0, big_a.ptr, 512, big_b.ptr, 512, 0.0, P.ptr, 512, 0, s))  # tcgen05 path

    with torch.cuda.stream(side):
        step()                                                                                          # warm-up (eager)
    side.synchronize()
    eager = [x.to_host().copy() for x in (H, S, rs, cs, wsum, P)]
    for x in (H, S, rs, cs, wsum, P):
        jz._lib.check(L.jz_fill(x.ptr, x.size, -7.0, None))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for want, x in zip(eager, (H, S, rs, cs, wsum, P)):
        assert np.array_equal(x.to_host().view(np.uint32), want.view(np.uint32))

"""GPU parity for the transformer helper kernels and the strided-batched GEMM, through the C ABI,
against the oracle's line-by-line restatement of the reference's kernels (ml/layer.hpp:2373-2538).
Tolerances: elementwise 2 ulp where the arithmetic is the same (exp, scale), 1e-5 relative where a reduction's
summation order differs (north_star)."""
import ctypes

import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu


def dev(jz, a):
    return jz.CM(np.asfortranarray(np.asarray(a, dtype=np.float32).reshape(-1, 1)))


def host(m):
    return m.to_host().ravel()


@pytest.mark.parametrize("S,batch", [(1, 1), (7, 3), (32, 2), (33, 5), (64, 96), (128, 64), (200, 3), (1024, 2), (1700, 1)])
@pytest.mark.parametrize("causal", [0, 1])
def test_softmax_rows_batched(jz, port, S, batch, causal):
    rng = np.random.default_rng(S * 7 + batch)
    x = (rng.standard_normal(S * S * batch) * 3).astype(np.float32)
    want = port.softmax_rows_batched(x, S, batch, bool(causal), -1e9)
    dx, dy = dev(jz, x), jz.CM.empty("y", S * S * batch, 1)
    jz._lib.check(jz.lib().jz_softmax_rows_batched(dy.ptr, dx.ptr, S, batch, causal, -1e9, None))
    got = host(dy)
    assert np.allclose(got, want, rtol=1e-5, atol=1e-30), np.abs(got - want).max()
    rows = got.reshape(batch, S, S).sum(axis=1)      # sum over keys b for each (blk, a)
    assert np.allclose(rows, 1.0, atol=1e-5)
    # in place (y aliases x), and the standalone mask
    jz._lib.check(jz.lib().jz_softmax_rows_batched(dx.ptr, dx.ptr, S, batch, causal, -1e9, None))
    assert np.array_equal(host(dx), got)
    if causal:
        dm = dev(jz, x)
        jz._lib.check(jz.lib().jz_causal_mask(dm.ptr, S, batch, -1e9, None))
        m = host(dm).reshape(batch, S, S)            # [blk][b][a]
        b, a = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
        ref = x.reshape(batch, S, S).copy()
        ref[:, b > a] = -1e9
        assert np.array_equal(m, ref)


def test_against_the_reference_cpu_formulation(jz):
    """golden vectors from the UNMODIFIED reference's CPU formulation (row_softmax, LayerNorm<float>; tests/golden/
    ref_ml_golden.npz, scripts/make_golden_ml.py): the jz_* kernels agree to 1e-5 relative"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ml_golden.npz"))
    L = jz.lib()
    for S in (7, 64, 130):
        x = jz.CM(np.asfortranarray(g[f"sm_x_{S}"]))
        y = jz.CM.empty("y", S, S)
        jz._lib.check(L.jz_softmax_rows_batched(y.ptr, x.ptr, S, 1, 0, 0.0, None))
        want = g[f"sm_y_{S}"]
        assert np.all(np.abs(y.to_host() - want) <= 1e-5 * np.abs(want) + 1e-12)
    for S in (9, 70):
        A, dA, want = g[f"smb_A_{S}"], g[f"smb_dA_{S}"], g[f"smb_dS_{S}"]
        dA_, dT_, out = jz.CM(np.asfortranarray(A)), jz.CM(np.asfortranarray(dA.T)), jz.CM.empty("dS", S, S)
        jz._lib.check(L.jz_softmax_rows_backward(out.ptr, dA_.ptr, dT_.ptr, S, 1, 0.125, None))
        assert np.abs(out.to_host() - want).max() <= 1e-5 * np.abs(want).max()
    for k in ("5x7", "64x33", "300x12"):
        dim, N = (int(v) for v in k.split("x"))
        xm, gm, bm = jz.CM(np.asfortranarray(g[f"ln_x_{k}"])), dev(jz, g[f"ln_g_{k}"]), dev(jz, g[f"ln_b_{k}"])
        y, xh, inv = jz.CM.empty("y", dim, N), jz.CM.empty("xh", dim, N), jz.CM.empty("inv", N, 1)
        jz._lib.check(L.jz_layernorm_forward(y.ptr, xh.ptr, inv.ptr, xm.ptr, gm.ptr, bm.ptr, dim, N, None))
        assert np.allclose(host(inv), g[f"ln_inv_{k}"], rtol=1e-5)
        assert np.allclose(xh.to_host(), g[f"ln_xhat_{k}"], rtol=1e-5, atol=1e-5)
        assert np.allclose(y.to_host(), g[f"ln_y_{k}"], rtol=1e-5, atol=2e-5)
        dy, xh0, inv0 = jz.CM(np.asfortranarray(g[f"ln_dy_{k}"])), jz.CM(np.asfortranarray(g[f"ln_xhat_{k}"])), dev(jz, g[f"ln_inv_{k}"])
        dx = jz.CM.empty("dx", dim, N)
        jz._lib.check(L.jz_layernorm_backward(dx.ptr, dy.ptr, gm.ptr, xh0.ptr, inv0.ptr, dim, N, None))
        assert np.allclose(dx.to_host(), g[f"ln_dx_{k}"], rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("S,batch", [(1, 1), (9, 4), (32, 3), (70, 3), (128, 64), (1000, 2), (1551, 1), (1552, 1), (2100, 2)])
def test_softmax_rows_backward(jz, port, S, batch):
    rng = np.random.default_rng(S + batch)
    A = port.softmax_rows_batched(rng.standard_normal(S * S * batch).astype(np.float32), S, batch)
    dAT = rng.standard_normal(S * S * batch).astype(np.float32)
    want = port.softmax_rows_backward(A, dAT, S, batch, 0.125)
    out = jz.CM.empty("dS", S * S * batch, 1)
    dA_, dT_ = dev(jz, A), dev(jz, dAT)    # named: a temporary's buffer would go back to the pool before the launch
    jz._lib.check(jz.lib().jz_softmax_rows_backward(out.ptr, dA_.ptr, dT_.ptr, S, batch, 0.125, None))
    got = host(out)
    scale = np.abs(A.reshape(batch, S, S)).max() * (np.abs(dAT).max() + 1) * 0.125
    assert np.abs(got - want).max() <= 1e-5 * scale


@pytest.mark.parametrize("dim,N", [(1, 1), (5, 7), (64, 33), (256, 4096), (1000, 9), (4096, 64)])
def test_layernorm(jz, port, dim, N):
    rng = np.random.default_rng(dim * 3 + N)
    x = np.asfortranarray((rng.standard_normal((dim, N)) * 2 + 1).astype(np.float32))
    gamma, beta = rng.standard_normal(dim).astype(np.float32), rng.standard_normal(dim).astype(np.float32)
    y0, xh0, inv0 = port.layernorm_forward(x, gamma, beta)
    L = jz.lib()
    dxm, dg, db = jz.CM(x), dev(jz, gamma), dev(jz, beta)
    y, xh, inv = jz.CM.empty("y", dim, N), jz.CM.empty("xh", dim, N), jz.CM.empty("inv", N, 1)
    jz._lib.check(L.jz_layernorm_forward(y.ptr, xh.ptr, inv.ptr, dxm.ptr, dg.ptr, db.ptr, dim, N, None))
    if dim > 1:
        assert np.allclose(host(inv), inv0, rtol=2e-5)
    assert np.allclose(xh.to_host(), xh0, rtol=1e-4, atol=2e-5)
    assert np.allclose(y.to_host(), y0, rtol=1e-4, atol=5e-5)
    dy = np.asfortranarray(rng.standard_normal((dim, N)).astype(np.float32))
    want = port.layernorm_backward(dy, gamma, xh0, inv0)
    dxo = jz.CM.empty("dx", dim, N)
    ddy, dxh0, dinv0 = jz.CM(dy), jz.CM(xh0), dev(jz, inv0)
    jz._lib.check(L.jz_layernorm_backward(dxo.ptr, ddy.ptr, dg.ptr, dxh0.ptr, dinv0.ptr, dim, N, None))
    assert np.abs(dxo.to_host() - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-3)


@pytest.mark.parametrize("ta,tb", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("shape", [(128, 128, 64, 24), (64, 128, 128, 7), (33, 17, 9, 5), (256, 256, 256, 3)])
def test_gemm_strided_batched(jz, port, ta, tb, shape):
    """attention-sized members in one launch (small-product kernel, batch on grid.z); 256^3 members one by one on
    the tensor path; operand strides larger than the members (padding between them)"""
    m, n, k, batch = shape
    rng = np.random.default_rng(m + n + k + batch)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    sA, sB, sC = ar * ac + 8, br * bc + 4, m * n
    A = rng.standard_normal(sA * batch).astype(np.float32)
    B = rng.standard_normal(sB * batch).astype(np.float32)
    C0 = rng.standard_normal(sC * batch).astype(np.float32)
    dA, dB, dC = dev(jz, A), dev(jz, B), dev(jz, C0)
    jz._lib.check(jz.lib().jz_gemm_strided_batched(ta, tb, m, n, k, 0.5, dA.ptr, ar, sA, dB.ptr, br, sB, -1.0, dC.ptr, m, sC,
                                                   batch, 0, None))
    got = host(dC).reshape(batch, n, m)
    for i in range(batch):
        P = np.asfortranarray(A[i * sA:i * sA + ar * ac].reshape(ar, ac, order="F"))
        Q = np.asfortranarray(B[i * sB:i * sB + br * bc].reshape(br, bc, order="F"))
        truth = 0.5 * port.gemm(P, ta, Q, tb, f64=True).astype(np.float64) - C0[i * sC:(i + 1) * sC].reshape(m, n, order="F")
        assert rel_fro(got[i].T, truth) < 1e-5, (shape, ta, tb, i)

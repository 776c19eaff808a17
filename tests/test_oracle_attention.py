"""CPU tests: the oracle's restatements of the reference's transformer helper kernels (ml/layer.hpp:2373-2538).
The reference's own tests pin those kernels only against PyTorch dumps that cannot be produced here, so the oracle is
pinned two other ways: (1) against golden vectors generated from the UNMODIFIED reference's CPU formulation of the same
operations (row_softmax, LayerNorm<float>::forward/backward in ml/layer.hpp, through oracle/ref_shim_ml.cpp;
tests/golden/ref_ml_golden.npz, scripts/make_golden_ml.py) -- equal to a few ulp, the two being different but equivalent
operator orders (the softmax backward against the expression of TransformerLayer::backward, ml/layer.hpp:3351-3352, evaluated
with the reference's operators); (2) against float64 formulas."""
import os
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def port():
    return oracle.port()


@pytest.fixture(scope="module")
def ml_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ml_golden.npz"))


@pytest.mark.parametrize("S", [7, 64, 130])
def test_row_softmax_vs_reference_cpu_formulation(port, ml_golden, S):
    x, want = ml_golden[f"sm_x_{S}"], ml_golden[f"sm_y_{S}"]
    got = port.softmax_rows_batched(x.ravel(order="F"), S, 1).reshape(S, S, order="F")
    assert np.all(np.abs(got - want) <= 2e-6 * np.abs(want) + 1e-12)


@pytest.mark.parametrize("k", ["5x7", "64x33", "300x12"])
def test_layernorm_vs_reference_cpu_formulation(port, ml_golden, k):
    g = ml_golden
    y, xh, inv = port.layernorm_forward(g[f"ln_x_{k}"], g[f"ln_g_{k}"], g[f"ln_b_{k}"])
    assert np.allclose(inv, g[f"ln_inv_{k}"], rtol=2e-6)
    assert np.allclose(xh, g[f"ln_xhat_{k}"], rtol=1e-5, atol=4e-6)
    assert np.allclose(y, g[f"ln_y_{k}"], rtol=1e-5, atol=6e-6)
    dx = port.layernorm_backward(g[f"ln_dy_{k}"], g[f"ln_g_{k}"], g[f"ln_xhat_{k}"], g[f"ln_inv_{k}"])
    assert np.allclose(dx, g[f"ln_dx_{k}"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("S", [9, 70])
def test_softmax_backward_vs_reference_cpu_expression(port, ml_golden, S):
    """dS = A .* (dA - rowsum(A .* dA)) * scale as TransformerLayer::backward spells it on the CPU (ml/layer.hpp:3351-3352);
    the CUDA kernel the oracle restates takes dA transposed per block"""
    A, dA, want = ml_golden[f"smb_A_{S}"], ml_golden[f"smb_dA_{S}"], ml_golden[f"smb_dS_{S}"]
    dAT = np.asfortranarray(dA.T)
    got = port.softmax_rows_backward(A.ravel(order="F"), dAT.ravel(order="F"), S, 1, 0.125).reshape(S, S, order="F")
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()


def test_reference_ml_shim_reproduces_the_golden_vectors(ml_golden):
    """where the reference build travelled (oracle/_ref/libjzref_ml.so) it must reproduce the committed fixture bit for bit"""
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libjzref_ml.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libjzref_ml.so not built here")
    L = ctypes.CDLL(so)
    x = np.asfortranarray(ml_golden["sm_x_64"])
    y = np.empty_like(x, order="F")
    assert L.refml_row_softmax(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(64), ctypes.c_size_t(64),
                               y.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.array_equal(y.view(np.uint32), ml_golden["sm_y_64"].view(np.uint32))


def test_adam_restatement_bit_exact_vs_reference_cpu(port, ml_golden):
    """oracle/jz_oracle.c:jzo_adam_update against three consecutive steps of the unmodified reference's
    adam_update<float> (ml/util.cuh:165-257); where libjzref_ml.so travelled, that library against the fixture too"""
    g_in = ml_golden["adam_g_in"]
    n = g_in.shape[1]
    m, v = np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in range(3):
        bc1 = np.float32(1.0 / (1.0 - np.float64(np.float32(0.9)) ** (t + 1)))
        bc2 = np.float32(1.0 / (1.0 - np.float64(np.float32(0.999)) ** (t + 1)))
        g, m, v = port.adam_update(g_in[t], m, v, 0.01, 0.9, 0.999, 1e-8, bc1, bc2)
        assert np.array_equal(g.view(np.uint32), ml_golden["adam_update"][t].view(np.uint32))
        assert np.array_equal(m.view(np.uint32), ml_golden[f"adam_m_{t + 1}"].view(np.uint32))
        assert np.array_equal(v.view(np.uint32), ml_golden[f"adam_v_{t + 1}"].view(np.uint32))
    import ctypes
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libjzref_ml.so")
    if os.path.exists(so):
        L = ctypes.CDLL(so)
        g, m, v = g_in[0].copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        assert L.refml_adam_update(P(g), P(m), P(v), ctypes.c_size_t(n), ctypes.c_float(0.01), ctypes.c_float(0.9),
                                   ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_int(1)) == 0
        assert np.array_equal(g.view(np.uint32), ml_golden["adam_update"][0].view(np.uint32))


def blocks(x, S, batch):
    """(S, S*batch) column-major flat buffer -> [batch][a][b]"""
    return x.reshape(batch, S, S).transpose(0, 2, 1)   # flat index = blk*S*S + b*S + a


@pytest.mark.parametrize("S,batch,causal", [(1, 1, False), (7, 3, False), (7, 3, True), (64, 5, True), (130, 2, False)])
def test_softmax_rows_batched(port, S, batch, causal):
    rng = np.random.default_rng(S * 10 + batch)
    x = (rng.standard_normal(S * S * batch) * 3).astype(np.float32)
    y = port.softmax_rows_batched(x, S, batch, causal, -1e9)
    X = blocks(x.astype(np.float64), S, batch).copy()
    if causal:
        a, b = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
        X[:, b > a] = -1e9
    E = np.exp(X - X.max(axis=2, keepdims=True))
    want = E / E.sum(axis=2, keepdims=True)
    got = blocks(y, S, batch)
    assert np.allclose(got, want, rtol=2e-6, atol=1e-9)
    assert np.allclose(got.sum(axis=2), 1.0, atol=1e-5)
    if causal:
        assert np.all(got[:, np.triu_indices(S, 1)[0], np.triu_indices(S, 1)[1]] == 0.0)


@pytest.mark.parametrize("S,batch", [(1, 1), (9, 4), (70, 3)])
def test_softmax_rows_backward(port, S, batch):
    rng = np.random.default_rng(S + batch)
    a_flat = port.softmax_rows_batched((rng.standard_normal(S * S * batch)).astype(np.float32), S, batch)
    dAT = rng.standard_normal(S * S * batch).astype(np.float32)
    dS = port.softmax_rows_backward(a_flat, dAT, S, batch, 0.125)
    A = blocks(a_flat.astype(np.float64), S, batch)
    dA = blocks(dAT.astype(np.float64), S, batch).transpose(0, 2, 1)    # dA[a,b] = dAT[b + a*S]
    want = A * (dA - (A * dA).sum(axis=2, keepdims=True)) * 0.125
    assert np.allclose(blocks(dS, S, batch), want, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("dim,N", [(1, 1), (5, 7), (64, 33), (1000, 4)])
def test_layernorm_forward_backward(port, dim, N):
    rng = np.random.default_rng(dim * 3 + N)
    x = np.asfortranarray((rng.standard_normal((dim, N)) * 2 + 1).astype(np.float32))
    gamma = rng.standard_normal(dim).astype(np.float32)
    beta = rng.standard_normal(dim).astype(np.float32)
    y, xhat, inv = port.layernorm_forward(x, gamma, beta)
    X = x.astype(np.float64)
    mu, var = X.mean(axis=0), X.var(axis=0)
    inv64 = 1.0 / np.sqrt(var + 1e-5)
    xh64 = (X - mu) * inv64
    assert np.allclose(inv, inv64, rtol=1e-5)
    assert np.allclose(xhat, xh64, rtol=1e-4, atol=1e-5)
    assert np.allclose(y, gamma[:, None] * xh64 + beta[:, None], rtol=1e-4, atol=1e-5)
    dy = np.asfortranarray(rng.standard_normal((dim, N)).astype(np.float32))
    dx = port.layernorm_backward(dy, gamma, xhat, inv)
    dxh = gamma[:, None].astype(np.float64) * dy
    want = inv64 * (dxh - dxh.mean(axis=0) - xh64 * (dxh * xh64).mean(axis=0))
    assert np.allclose(dx, want, rtol=1e-3, atol=1e-4)

"""CPU tests: the oracle's restatements of the reference's transformer helper kernels (ml/layer.hpp:2373-2538)
against float64 numpy formulas.  These kernels have no golden vectors in the reference's tests that can run
here (PyTorch dumps), so the oracle is checked against the mathematics instead -- "parity unpinned" for SURVEY
row 8f-3, stated in oracle/jz_oracle.c and DESIGN.md."""
import numpy as np
import pytest

import oracle


@pytest.fixture(scope="module")
def port():
    return oracle.port()


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

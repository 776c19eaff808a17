"""Generates tests/golden/ref_ml_golden.npz from the UNMODIFIED reference's own CPU formulation of the transformer helpers
(oracle/_ref/libjzref_ml.so = oracle/ref_shim_ml.cpp over ml/layer.hpp: row_softmax, LayerNorm<float>::forward/backward).
Run in the container that has /root/reference:   make -C oracle ref-ml && python scripts/make_golden_ml.py"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libjzref_ml.so"))
F = ctypes.c_void_p


def p(a):
    return a.ctypes.data_as(F)


out = {}
rng = np.random.default_rng(20261018)
for S in (7, 64, 130):
    x = np.asfortranarray((rng.standard_normal((S, S)) * 3).astype(np.float32))
    y = np.empty_like(x, order="F")
    assert L.refml_row_softmax(p(x), ctypes.c_size_t(S), ctypes.c_size_t(S), p(y)) == 0
    out[f"sm_x_{S}"], out[f"sm_y_{S}"] = x, y
for S in (9, 70):
    A = out.get(f"sm_y_{S}")
    if A is None:
        x = np.asfortranarray((rng.standard_normal((S, S)) * 2).astype(np.float32))
        A = np.empty_like(x, order="F")
        assert L.refml_row_softmax(p(x), ctypes.c_size_t(S), ctypes.c_size_t(S), p(A)) == 0
    dA = np.asfortranarray(rng.standard_normal((S, S)).astype(np.float32))
    dS = np.empty_like(dA, order="F")
    assert L.refml_softmax_backward(p(A), p(dA), ctypes.c_size_t(S), ctypes.c_float(0.125), p(dS)) == 0
    out[f"smb_A_{S}"], out[f"smb_dA_{S}"], out[f"smb_dS_{S}"] = A, dA, dS
for dim, N in ((5, 7), (64, 33), (300, 12)):
    x = np.asfortranarray((rng.standard_normal((dim, N)) * 2 + 1).astype(np.float32))
    g = rng.standard_normal(dim).astype(np.float32)
    b = rng.standard_normal(dim).astype(np.float32)
    y, xh = np.empty_like(x, order="F"), np.empty_like(x, order="F")
    inv = np.empty(N, dtype=np.float32)
    assert L.refml_layernorm_forward(p(x), p(g), p(b), ctypes.c_size_t(dim), ctypes.c_size_t(N), p(y), p(xh), p(inv)) == 0
    dy = np.asfortranarray(rng.standard_normal((dim, N)).astype(np.float32))
    dx = np.empty_like(x, order="F")
    assert L.refml_layernorm_backward(p(dy), p(g), p(xh), p(inv), ctypes.c_size_t(dim), ctypes.c_size_t(N), p(dx)) == 0
    k = f"{dim}x{N}"
    out.update({f"ln_x_{k}": x, f"ln_g_{k}": g, f"ln_b_{k}": b, f"ln_y_{k}": y, f"ln_xhat_{k}": xh, f"ln_inv_{k}": inv,
                f"ln_dy_{k}": dy, f"ln_dx_{k}": dx})
# Adam: three consecutive steps of the reference's adam_update<float> (ml/util.cuh:165-257) from its own initial state
n = 4099
g0 = (rng.standard_normal((3, n)) * np.array([[1.0], [1e-3], [30.0]])).astype(np.float32)
m = np.zeros(n, dtype=np.float32)
v = np.zeros(n, dtype=np.float32)
out["adam_g_in"] = g0.copy()
upd = np.empty_like(g0)
for t in range(3):
    g = g0[t].copy()
    assert L.refml_adam_update(p(g), p(m), p(v), ctypes.c_size_t(n), ctypes.c_float(0.01), ctypes.c_float(0.9),
                               ctypes.c_float(0.999), ctypes.c_float(1e-8), ctypes.c_int(t + 1)) == 0
    upd[t] = g
    out[f"adam_m_{t + 1}"], out[f"adam_v_{t + 1}"] = m.copy(), v.copy()
out["adam_update"] = upd
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_ml_golden.npz"), **out)
print("wrote", len(out), "arrays")

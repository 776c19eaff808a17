"""oracle -- TEST INFRASTRUCTURE ONLY (checker, never the product path).

Two CPU checkers with one numpy-facing interface:

* ``port()``  -- oracle/jz_oracle.c, the plain-C restatement (travels as source, built
  by ``oracle/Makefile port``).
* ``ref()``   -- oracle/_ref/libjzref.so, the UNMODIFIED reference ``Matrix<float>``
  path compiled from /root/reference by ``oracle/Makefile ref`` (git-ignored; present
  wherever the prebuilt file travelled).  ``ref(noblas=True)`` is the
  ``-DJUZHEN_NO_BLAS`` build with the reference's fixed-order kernels.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this package.

A matrix is passed as ``(array, trans)`` where ``array`` is the PHYSICAL buffer as a
2-D float32 Fortran-ordered (column-major) numpy array ``numrow x numcol`` and ``trans`` the
reference's lazy transpose flag (cpp/core.hpp:96-98).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_uint, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OPS = {"exp": 0, "log": 1, "tanh": 2, "dtanh": 3, "square": 4, "sqrt": 5, "relu": 6, "drelu": 7}


def _f(a):
    a = np.asarray(a)
    assert a.dtype == np.float32
    return a.ctypes.data_as(c_void_p)


def _phys(a):
    """physical buffer as F-ordered float32 2-D"""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


def _ldims(a, t):
    return (a.shape[1], a.shape[0]) if t else a.shape


class Oracle:
    """numpy interface over either checker library (prefix ``jzo_`` or ``ref_``)."""

    def __init__(self, lib, prefix, kind):
        self.lib, self.pre, self.kind = lib, prefix, kind

    def _fn(self, name, restype=c_int):
        fn = getattr(self.lib, self.pre + name)
        fn.restype = restype
        return fn

    # ---- elementwise over flat buffers
    def unary(self, op, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        rc = self._fn("unary")(c_int(OPS[op]), _f(x), _f(out), c_size_t(x.size))
        assert rc == 0
        return out

    def affine(self, x, s1, a):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._fn("affine")(_f(x), _f(out), c_size_t(x.size), c_float(s1), c_float(a))
        return out

    def div_scalar(self, x, r):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._fn("div_scalar")(_f(x), _f(out), c_size_t(x.size), c_double(r))
        return out

    def eleminv(self, x, l):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._fn("eleminv")(_f(x), _f(out), c_size_t(x.size), c_double(l))
        return out

    def chain_softplus5(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        self._fn("chain_softplus5")(_f(x), _f(out), c_size_t(x.size))
        return out

    # ---- binary, transpose-aware.  Returns logical result (F-ordered) or None on
    # the reference's std::invalid_argument.
    def _bin(self, name, A, ta, B, tb, *scalars):
        A, B = _phys(A), _phys(B)
        R, C = _ldims(A, ta)
        out = np.empty((R, C), dtype=np.float32, order="F")
        args = [_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta),
                _f(B), c_size_t(B.shape[0]), c_size_t(B.shape[1]), c_int(tb)]
        args += [c_float(s) for s in scalars]
        rc = self._fn(name)(*args, _f(out))
        if rc == 2:
            return None
        assert rc == 0
        return out

    def axpby(self, A, ta, B, tb, s1, s2):
        return self._bin("axpby", A, ta, B, tb, s1, s2)

    def hadmd(self, A, ta, B, tb):
        return self._bin("hadmd", A, ta, B, tb)

    def div(self, A, ta, B, tb):
        return self._bin("div", A, ta, B, tb)

    # ---- reductions
    def sum(self, A, ta, dim, f64=False):
        A = _phys(A)
        R, C = _ldims(A, ta)
        out = np.empty(C if dim == 0 else R, dtype=np.float32)
        name = "sum_f64" if f64 else "sum"
        self._fn(name)(_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta), c_int(dim), _f(out))
        return out

    def reduce(self, which, A, ta, dim):
        """which: 'max' (k=1) or 'stats' (serial sum + max, k=2); logical result."""
        A = _phys(A)
        R, C = _ldims(A, ta)
        k = 1 if which == "max" else 2
        shape = (k, C) if dim == 0 else (R, k)
        out = np.empty(shape, dtype=np.float32, order="F")
        a = (_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta), c_int(dim), _f(out))
        if self.pre == "jzo_":
            self._fn("reduce")(c_int(0 if which == "max" else 1), *a)
        else:
            self._fn("reduce_max" if which == "max" else "reduce_stats")(*a)
        return out

    def softmax_cols(self, X):
        X = _phys(X)
        out = np.empty_like(X, order="F")
        self._fn("softmax_cols")(_f(X), c_size_t(X.shape[0]), c_size_t(X.shape[1]), _f(out))
        return out

    def softmax_ce_grad(self, X, Y, nb):
        X, Y = _phys(X), _phys(Y)
        out = np.empty_like(X, order="F")
        self._fn("softmax_ce_grad")(_f(X), _f(Y), c_size_t(X.shape[0]), c_size_t(X.shape[1]),
                                    c_double(nb), _f(out))
        return out

    # ---- transformer helper kernels (port only; see the PARITY UNPINNED note in jz_oracle.c)
    def softmax_rows_batched(self, x, S, batch, causal=False, mask_val=-1e9):
        x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        y = np.empty_like(x)
        self._fn("softmax_rows_batched")(_f(x), _f(y), c_size_t(S), c_size_t(batch), c_int(int(causal)), c_float(mask_val))
        return y

    def softmax_rows_backward(self, A, dAT, S, batch, scale):
        A = np.ascontiguousarray(A, dtype=np.float32).ravel()
        dAT = np.ascontiguousarray(dAT, dtype=np.float32).ravel()
        dS = np.empty_like(A)
        self._fn("softmax_rows_backward")(_f(A), _f(dAT), _f(dS), c_size_t(S), c_size_t(batch), c_float(scale))
        return dS

    def layernorm_forward(self, x, gamma, beta):
        x = _phys(x)
        dim, N = x.shape
        gamma = np.ascontiguousarray(gamma, dtype=np.float32).ravel()
        beta = np.ascontiguousarray(beta, dtype=np.float32).ravel()
        y, xhat = np.empty_like(x, order="F"), np.empty_like(x, order="F")
        inv = np.empty(N, dtype=np.float32)
        self._fn("layernorm_forward")(_f(x), _f(gamma), _f(beta), _f(y), _f(xhat), _f(inv), c_size_t(dim), c_size_t(N))
        return y, xhat, inv

    def layernorm_backward(self, dy, gamma, xhat, inv_std):
        dy, xhat = _phys(dy), _phys(xhat)
        dim, N = dy.shape
        gamma = np.ascontiguousarray(gamma, dtype=np.float32).ravel()
        inv_std = np.ascontiguousarray(inv_std, dtype=np.float32).ravel()
        dx = np.empty_like(dy, order="F")
        self._fn("layernorm_backward")(_f(dy), _f(gamma), _f(xhat), _f(inv_std), _f(dx), c_size_t(dim), c_size_t(N))
        return dx

    def adam_update(self, g, m, v, alpha, beta1, beta2, eps, bc1, bc2):
        """one Adam step (ml/util.cuh:223-245); returns the updated (g, m, v) copies"""
        g, m, v = (np.array(x, dtype=np.float32, copy=True).ravel() for x in (g, m, v))
        self._fn("adam_update")(_f(g), _f(m), _f(v), c_size_t(g.size), c_float(alpha), c_float(beta1), c_float(beta2),
                                c_float(eps), c_float(bc1), c_float(bc2))
        return g, m, v

    def norm(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        return float(self._fn("norm", c_float)(_f(x), c_size_t(x.size)))

    # ---- GEMM
    def gemm(self, A, ta, B, tb, f64=False):
        A, B = _phys(A), _phys(B)
        m, k = _ldims(A, ta)
        k2, n = _ldims(B, tb)
        out = np.empty((m, n), dtype=np.float32, order="F")
        name = "gemm_f64" if f64 else "gemm"
        rc = self._fn(name)(_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta),
                            _f(B), c_size_t(B.shape[0]), c_size_t(B.shape[1]), c_int(tb), _f(out))
        if rc == 2:
            return None
        assert rc == 0
        return out

    # ---- data movement
    def materialize(self, A, ta):
        A = _phys(A)
        R, C = _ldims(A, ta)
        out = np.empty((R, C), dtype=np.float32, order="F")
        self._fn("materialize")(_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta), _f(out))
        return out

    def slice(self, A, ta, r0, r1, c0, c1):
        A = _phys(A)
        out = np.empty((r1 - r0, c1 - c0), dtype=np.float32, order="F")
        self._fn("slice")(_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta),
                          c_size_t(r0), c_size_t(r1), c_size_t(c0), c_size_t(c1), _f(out))
        return out

    def slice_set(self, D, dt, r0, r1, c0, c1, S, st):
        """returns a modified copy of D's physical buffer"""
        D, S = _phys(D).copy(order="F"), _phys(S)
        self._fn("slice_set")(_f(D), c_size_t(D.shape[0]), c_size_t(D.shape[1]), c_int(dt),
                              c_size_t(r0), c_size_t(r1), c_size_t(c0), c_size_t(c1),
                              _f(S), c_size_t(S.shape[0]), c_size_t(S.shape[1]), c_int(st))
        return D

    def stack(self, vertical, mats):
        """mats: list of (array, trans).  Returns logical result or None (invalid_argument)."""
        mats = [(_phys(a), t) for a, t in mats]
        n = len(mats)
        ptrs = (c_void_p * n)(*[a.ctypes.data for a, _ in mats])
        nr = (c_size_t * n)(*[a.shape[0] for a, _ in mats])
        nc = (c_size_t * n)(*[a.shape[1] for a, _ in mats])
        tr = (c_int * n)(*[int(t) for _, t in mats])
        total = sum(a.size for a, _ in mats)
        out = np.empty(max(total, 1), dtype=np.float32)
        orow, ocol = c_size_t(0), c_size_t(0)
        rc = self._fn("stack")(c_int(1 if vertical else 0), c_int(n), ptrs, nr, nc, tr, _f(out),
                               ctypes.byref(orow), ctypes.byref(ocol))
        if rc == 2:
            return None
        assert rc == 0
        return out[: orow.value * ocol.value].reshape((orow.value, ocol.value), order="F").copy(order="F")

    # ---- seeded inputs as the reference draws them
    def randn(self, seed, n):
        out = np.empty(n, dtype=np.float32)
        self._fn("randn")(c_uint(seed), _f(out), c_size_t(n))
        return out

    def rand(self, seed, n):
        out = np.empty(n, dtype=np.float32)
        self._fn("rand")(c_uint(seed), _f(out), c_size_t(n))
        return out

    def gemm_subblock(self, A, ta, B, tb, r0, r1, c0, c1):
        """rows [r0, r1) x columns [c0, c1) of op(A)*op(B) via the reference's rows()/columns()/dot (reference build only)"""
        A, B = _phys(A), _phys(B)
        out = np.empty((r1 - r0, c1 - c0), dtype=np.float32, order="F")
        rc = self._fn("gemm_subblock")(_f(A), c_size_t(A.shape[0]), c_size_t(A.shape[1]), c_int(ta),
                                       _f(B), c_size_t(B.shape[0]), c_size_t(B.shape[1]), c_int(tb),
                                       c_size_t(r0), c_size_t(r1), c_size_t(c0), c_size_t(c1), _f(out))
        if rc == 2:
            raise ValueError("Matrix dimensions are not compatible")
        return out

    # ---- reference-only helpers
    def testbasic_expr(self, A, B):
        A, B = _phys(A), _phys(B)
        out = np.empty((2, 3), dtype=np.float32, order="F")
        self._fn("testbasic_expr")(_f(A), _f(B), _f(out))
        return out

    def config1(self, A, B, scale):
        A, B = _phys(A), _phys(B)
        n = A.shape[0]
        out = np.empty((n, n), dtype=np.float32, order="F")
        self._fn("config1")(_f(A), _f(B), c_size_t(n), c_double(scale), _f(out))
        return out

    def blas_threads(self, n=0):
        return int(self._fn("blas_threads")(c_int(n)))

    def blas_config(self):
        return self._fn("blas_config", c_char_p)().decode()


_cache = {}


def build_port():
    """compile oracle/jz_oracle.c (gcc only; works on the GPU box too)."""
    subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)


def port() -> Oracle:
    if "port" not in _cache:
        so = os.path.join(_HERE, "libjzoracle.so")
        src = os.path.join(_HERE, "jz_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build_port()
        _cache["port"] = Oracle(ctypes.CDLL(so), "jzo_", "port")
    return _cache["port"]


def ref_available(noblas=False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libjzref_noblas.so" if noblas else "libjzref.so"))


def ref(noblas=False) -> Oracle:
    key = "ref_noblas" if noblas else "ref"
    if key not in _cache:
        so = os.path.join(_HERE, "_ref", "libjzref_noblas.so" if noblas else "libjzref.so")
        _cache[key] = Oracle(ctypes.CDLL(so), "ref_", "reference")
    return _cache[key]

"""Python mirror of the reference's ``Matrix<CUDAfloat>`` interface over the C-ABI.

Same names, argument meaning and error behaviour as cpp/cumatrix.cuh:92-276 +
cpp/operators.hpp (so the parity tests read like tests/testbasic.cu and
tests/testRocmCpuParity.cu): column-major storage with a lazy ``transpose`` flag, ``T()``
aliasing the buffer, shape errors raised as ``JzShapeError`` (the reference's
``std::invalid_argument("Matrix dimensions are not compatible")``) before any launch.

C++ distinguishes lvalue/rvalue overloads (an rvalue operand is updated in place and its
buffer returned -- pointer identity is tested by tests/testElementwiseReduceTorchDump.cu:45-48).
Python has no rvalues, so every function that has an ``&&`` overload takes ``inplace=True`` to
select it.  All device work happens in libjz_b200.so; this file only does shape logic.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_void_p

import numpy as np

from . import _lib
from ._lib import check, lib


class _Buffer:
    """pool-owned device buffer (shared_ptr<CUDAfloat[]> + Memory<CUDAfloat> deleter)."""

    __slots__ = ("ptr", "count", "stream")

    def __init__(self, count, stream=None):
        p = c_void_p()
        check(lib().jz_malloc(ctypes.byref(p), count, stream))
        self.ptr, self.count, self.stream = p.value, count, stream

    def __del__(self):
        try:
            if self.ptr:
                lib().jz_free(self.ptr, self.stream)
        except Exception:
            pass


_stream = None  # cudaStream_t as int; None = legacy default stream (what the reference uses)


def set_stream(s):
    global _stream
    _stream = s


def get_stream():
    return _stream


def sync():
    check(lib().jz_sync(_stream))


class CM:
    """``Matrix<CUDAfloat>``: fp32, column-major numrow x numcol, lazy transpose flag."""

    def __init__(self, host=None, name="cu_M", *, _raw=None):
        if _raw is not None:
            self.name, self.numrow, self.numcol, self.transpose, self.buf = _raw
            return
        if host is None:  # default ctor: zeroed 2 x 2 (cpp/cumatrix.cuh:124)
            host = np.zeros((2, 2), dtype=np.float32)
        a = np.asarray(host, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        a = np.asfortranarray(a)
        self.name, self.numrow, self.numcol, self.transpose = name, a.shape[0], a.shape[1], False
        self.buf = _Buffer(a.size, _stream)
        if a.size:
            check(lib().jz_memcpy_h2d(self.buf.ptr, a.ctypes.data, a.size, _stream))
            sync()  # the reference's upload is a synchronous cudaMemcpy (cpp/cumatrix.cu:42)

    # ---- construction helpers
    @staticmethod
    def empty(name, numrow, numcol, trans=False):
        """internal temporary: NOT zero-filled (the reference zero-fills every temporary)"""
        return CM(_raw=(name, numrow, numcol, bool(trans), _Buffer(numrow * numcol, _stream)))

    @staticmethod
    def named(name, numrow, numcol):
        """public ctor Matrix(name, r, c): observably zero-filled (cpp/cumatrix.cu:50-63)"""
        m = CM.empty(name, numrow, numcol)
        m.zeros()
        return m

    @staticmethod
    def from_physical(buf, trans, name="cu_M"):
        m = CM(np.asfortranarray(buf, dtype=np.float32), name)
        m.transpose = bool(trans)
        return m

    @staticmethod
    def zeros_(m, n):
        return CM.named("zeros", m, n)

    @staticmethod
    def ones_(m, n):
        r = CM.empty("ones", m, n)
        r.ones()
        return r

    # `offset` is the stream id of the counter-based generator (jz_rng.cu).  When it is not given, successive draws
    # take successive streams, like GPUSampler::offset in the C++ shell: two default-argument calls never return the
    # same matrix.  Pass seed AND offset for a reproducible draw.
    _next_stream = 0

    @staticmethod
    def _stream_id(offset):
        if offset is not None:
            return int(offset)
        CM._next_stream += 1
        return CM._next_stream - 1

    @staticmethod
    def randn(m, n, seed=0, offset=None):
        r = CM.empty("randn", m, n)
        check(lib().jz_rand_normal(r.ptr, m * n, seed, CM._stream_id(offset), _stream))
        return r

    @staticmethod
    def rand(m, n, seed=0, offset=None):
        r = CM.empty("rand", m, n)
        check(lib().jz_rand_uniform(r.ptr, m * n, seed, CM._stream_id(offset), _stream))
        return r

    # ---- info
    @property
    def ptr(self):
        return self.buf.ptr

    def data(self):
        return self.buf.ptr

    def num_row(self):
        return self.numcol if self.transpose else self.numrow

    def num_col(self):
        return self.numrow if self.transpose else self.numcol

    def get_transpose(self):
        return int(self.transpose)

    def get_name(self):
        return self.name

    @property
    def size(self):
        return self.numrow * self.numcol

    def copy(self):
        r = CM.empty("copy of" + self.name, self.numrow, self.numcol, self.transpose)
        check(lib().jz_memcpy_d2d(r.ptr, self.ptr, self.size, _stream))
        return r

    # ---- fillers
    def zeros(self):
        check(lib().jz_fill(self.ptr, self.size, 0.0, _stream))

    def ones(self):
        check(lib().jz_fill(self.ptr, self.size, 1.0, _stream))

    # ---- host transfer
    def to_host_physical(self):
        out = np.empty((self.numrow, self.numcol), dtype=np.float32, order="F")
        if out.size:
            check(lib().jz_memcpy_d2h(out.ctypes.data, self.ptr, out.size, _stream))
        return out

    def to_host(self):
        """logical matrix (what Matrix<float>::elem(i,j) shows after to_host())"""
        p = self.to_host_physical()
        return np.asfortranarray(p.T) if self.transpose else p

    # ---- lazy transpose: shares the buffer, flips the flag (cpp/cumatrix.cu:305-310)
    def T(self):
        return CM(_raw=(self.name + "_T", self.numrow, self.numcol, not self.transpose, self.buf))

    # ---- GEMM (cpp/cumatrix.cu:177-197)
    def dot(self, B, mode=-1):
        if self.num_col() != B.num_row():
            raise _lib.JzShapeError("Matrix dimensions are not compatible")
        m, n, k = self.num_row(), B.num_col(), self.num_col()
        C = CM.empty("dot", m, n)
        check(lib().jz_gemm(int(self.transpose), int(B.transpose), m, n, k, 1.0, self.ptr, self.numrow,
                            B.ptr, B.numrow, 0.0, C.ptr, max(m, 1), mode, _stream))
        return C

    # ---- s1*M + s2*B (cpp/cumatrix.cu:227-260)
    def add(self, B, s1=None, s2=None, inplace=False):
        if not isinstance(B, CM):  # add(float a, float s1): s1*M + a (cpp/cumatrix.cu:199-215)
            a, sc = float(B), float(s1)
            out = self if inplace else CM.empty("add", self.numrow, self.numcol, self.transpose)
            check(lib().jz_affine(out.ptr, self.ptr, self.size, sc, a, _stream))
            return out
        if self.num_row() != B.num_row() or self.num_col() != B.num_col():
            raise _lib.JzShapeError("Matrix dimensions are not compatible")
        if inplace:
            # result keeps `this` layout; B is read transposed iff the flags differ
            bt = int(self.transpose != B.transpose)
            if not bt:
                check(lib().jz_axpby(self.ptr, self.ptr, B.ptr, self.size, s1, s2, _stream))
            else:
                check(lib().jz_axpby2d(self.ptr, self.numrow, self.numrow, self.numcol, self.ptr, self.numrow, 0,
                                       B.ptr, B.numrow, 1, s1, s2, _stream))
            return self
        R, Cc = self.num_row(), self.num_col()
        C = CM.empty("add", R, Cc)
        if not self.transpose and not B.transpose:
            check(lib().jz_axpby(C.ptr, self.ptr, B.ptr, self.size, s1, s2, _stream))
        else:
            check(lib().jz_axpby2d(C.ptr, R, R, Cc, self.ptr, self.numrow, int(self.transpose),
                                   B.ptr, B.numrow, int(B.transpose), s1, s2, _stream))
        return C

    def scale(self, s1, inplace=False):
        # const: add(0, s1); in place the reference calls cublasSscal (cpp/cumatrix.cu:218-224)
        return self.add(0.0, s1, inplace=inplace)

    def eleminv(self, l, inplace=False):
        out = self if inplace else CM.empty("elem_rec", self.numrow, self.numcol, self.transpose)
        check(lib().jz_eleminv(out.ptr, self.ptr, self.size, float(l), _stream))
        return out

    def norm(self):
        r = c_float(0)
        check(lib().jz_nrm2(self.ptr, self.size, ctypes.byref(r), _stream))
        return float(r.value)

    # ---- slicing (cpp/cumatrix.cuh:188-237): window is in LOGICAL coordinates
    def slice(self, rstart, rend, cstart, cend, M=None):
        if self.transpose:
            rstart, cstart, rend, cend = cstart, rstart, cend, rend
        rows, cols = rend - rstart, cend - cstart
        off = cstart * self.numrow + rstart
        if M is None:
            out = CM.empty("submatrix", rows, cols, self.transpose)
            check(lib().jz_copy2d(out.ptr, max(rows, 1), self.ptr + 4 * off, self.numrow, rows, cols, 0, _stream))
            return out
        # assignment in LOGICAL coordinates (the CPU oracle's semantics, cpp/core.hpp:365-373); the
        # reference's copyKernel ignores M's flag, which agrees with this whenever the flags match
        check(lib().jz_copy2d(self.ptr + 4 * off, self.numrow, M.ptr, M.numrow, rows, cols,
                              int(self.transpose != M.transpose), _stream))
        return None

    def rows(self, rstart, rend, M=None):
        return self.slice(rstart, rend, 0, self.num_col(), M)

    def columns(self, cstart, cend, M=None):
        return self.slice(0, self.num_row(), cstart, cend, M)

    # ---- operators (cpp/operators.hpp:109-290; scalars arrive as double)
    def __matmul__(self, B):
        return self.dot(B)

    def __mul__(self, r):
        if isinstance(r, CM):
            return self.dot(r)
        return self.scale(float(r))

    def __rmul__(self, l):
        return self.scale(float(l))

    def __add__(self, r):
        if isinstance(r, CM):
            return self.add(r, 1.0, 1.0)
        return self.add(float(r), 1.0)

    __radd__ = __add__

    def __sub__(self, r):
        if isinstance(r, CM):
            return self.add(r, 1.0, -1.0)
        return self.add(-float(r), 1.0)

    def __rsub__(self, l):
        return self.add(float(l), -1.0)

    def __neg__(self):
        return self.add(0.0, -1.0)

    def __iadd__(self, r):
        if isinstance(r, CM):
            return self.add(r, 1.0, 1.0, inplace=True)
        return self.add(float(r), 1.0, inplace=True)

    def __isub__(self, r):
        if isinstance(r, CM):
            return self.add(r, 1.0, -1.0, inplace=True)
        return self.add(-float(r), 1.0, inplace=True)

    def __truediv__(self, r):
        if isinstance(r, CM):  # A / B = hadmd(A, B.eleminv(1.0)); shapes are checked before any launch
            if self.num_row() != r.num_row() or self.num_col() != r.num_col():
                raise _lib.JzShapeError("Matrix dimensions are not compatible")
            return hadmd(self, r.eleminv(1.0), inplace_rhs=True)
        return self.scale(float(np.float32(1.0 / float(r))))  # multiply by (float)(1.0/r)

    def __rtruediv__(self, l):
        return self.eleminv(float(l))


# ---------------------------------------------------------------- free functions
def _unary(op, M, inplace):
    out = M if inplace else CM.empty(op + "M", M.numrow, M.numcol, M.transpose)
    check(lib().jz_unary(_lib.UNARY[op], out.ptr, M.ptr, M.size, _stream))
    return out


def exp(M, inplace=False):
    return _unary("exp", M, inplace)


def log(M):  # the reference has no rvalue overload (cpp/matrix.hpp:269)
    return _unary("log", M, False)


def tanh(M, inplace=False):
    return _unary("tanh", M, inplace)


def d_tanh(M, inplace=False):
    return _unary("dtanh", M, inplace)


def square(M, inplace=False):
    return _unary("square", M, inplace)


def sqrt(M):
    return _unary("sqrt", M, False)


def relu(M, inplace=True):
    return _unary("relu", M, inplace)


def d_relu(M, inplace=True):
    return _unary("drelu", M, inplace)


def chain(M, steps, inplace=False):
    """fused elementwise chain, e.g. [('exp',), ('affine', 1, 1), ('log',), ('affine', 0.2, 0)]"""
    out = M if inplace else CM.empty("chain", M.numrow, M.numcol, M.transpose)
    arr, n = _lib.make_steps(steps)
    check(lib().jz_chain(out.ptr, M.ptr, M.size, arr, n, _stream))
    return out


def hadmd(M1, M2, inplace_lhs=False, inplace_rhs=False):
    """M1 .* M2, result in M1's layout (cpp/cukernels.cu:326-400)"""
    if M1.num_row() != M2.num_row() or M1.num_col() != M2.num_col():
        raise _lib.JzShapeError("Matrix dimensions are not compatible")
    same = M1.transpose == M2.transpose
    if same:
        out = M1 if inplace_lhs else (M2 if inplace_rhs else CM.empty("hadmd", M1.numrow, M1.numcol, M1.transpose))
        check(lib().jz_hadamard(out.ptr, M1.ptr, M2.ptr, M1.size, _stream))
        return out
    out = CM.empty("hadmd", M1.numrow, M1.numcol, M1.transpose)
    check(lib().jz_hadamard2d(out.ptr, M1.numrow, M1.numrow, M1.numcol, M1.ptr, M1.numrow, 0,
                              M2.ptr, M2.numrow, 1, _stream))
    return out


def _reduce(fn, M, dim):
    # physical reduction direction: logical dim 0 on a transposed matrix is physical dim 1
    pdim = dim if not M.transpose else 1 - dim
    n_out = M.numcol if pdim == 0 else M.numrow
    out = CM.empty("sumM", n_out, 1)
    check(fn(out.ptr, M.ptr, M.numrow, M.numcol, M.numrow, pdim, _stream))
    if dim == 0:
        out.transpose = True  # 1 x ncols, stored as ncols x 1 transposed (cpp/cumatrix.cu:332-336)
    return out


def sum(M, dim):  # noqa: A001 - mirrors the reference name
    return _reduce(lib().jz_sum, M, dim)


def colmax(M, dim=0):
    """reduce() with the LogisticLayer max functor (ml/layer.hpp:254-259)"""
    return _reduce(lib().jz_max, M, dim)


def softmax_cols(M):
    assert not M.transpose
    out = CM.empty("softmax", M.numrow, M.numcol)
    check(lib().jz_softmax_cols(out.ptr, M.ptr, M.numrow, M.numcol, M.numrow, _stream))
    return out


def softmax_ce_grad(X, Y, nb):
    assert not X.transpose and not Y.transpose
    out = CM.empty("cegrad", X.numrow, X.numcol)
    check(lib().jz_softmax_ce_grad(out.ptr, X.ptr, Y.ptr, X.numrow, X.numcol, float(nb), _stream))
    return out


def fill(M, a):
    check(lib().jz_fill(M.ptr, M.size, float(a), _stream))
    return M


def hstack(mats):
    mats = [m for m in mats if m.num_row() != 0 and m.num_col() != 0]  # empties dropped (cukernels.cu:260-264)
    if not mats:
        raise _lib.JzShapeError("hstack: input list is empty or contains only empty matrices")
    R = mats[0].num_row()
    for m in mats:
        if m.num_row() != R:
            raise _lib.JzShapeError("hstack: all matrices must have the same row count")
    C = builtins_sum(m.num_col() for m in mats)
    out = CM.empty("hstack", R, C)
    col = 0
    for m in mats:
        check(lib().jz_copy2d(out.ptr + 4 * col * R, R, m.ptr, m.numrow, R, m.num_col(), int(m.transpose), _stream))
        col += m.num_col()
    return out


def vstack(mats):
    """materialised, non-transposed result (tests/testStackOps.cu:248-270)"""
    mats = [m for m in mats if m.num_row() != 0 and m.num_col() != 0]
    if not mats:
        raise _lib.JzShapeError("vstack: input list is empty or contains only empty matrices")
    C = mats[0].num_col()
    for m in mats:
        if m.num_col() != C:
            raise _lib.JzShapeError("vstack: all matrices must have the same column count")
    R = builtins_sum(m.num_row() for m in mats)
    out = CM.empty("vstack", R, C)
    row = 0
    for m in mats:
        check(lib().jz_copy2d(out.ptr + 4 * row, R, m.ptr, m.numrow, m.num_row(), C, int(m.transpose), _stream))
        row += m.num_row()
    return out


import builtins as _b  # noqa: E402

builtins_sum = _b.sum

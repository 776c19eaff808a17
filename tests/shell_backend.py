"""ctypes binding of build/dropin/lib/libjz_shell_capi.so (juzhen_b200/cpp/tests/shell_capi.cu): the C++ shell --
``Matrix<CUDAfloat>`` as a C++ user of the reference compiles against it -- behind a Python surface shaped like
``juzhen_b200.matrix`` so that the same expression can be run on both and compared bit for bit.  Test infrastructure."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "build", "dropin", "lib", "libjz_shell_capi.so")
_lib = None


class ShellShapeError(ValueError):
    """std::invalid_argument("Matrix dimensions are not compatible") thrown by the shell"""


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        sig = {
            "shell_init": (c_int, [c_int]), "shell_last_error": (c_char_p, []), "shell_last_error_is_shape": (c_int, []),
            "shell_from_host": (c_void_p, [c_void_p, c_size_t, c_size_t]), "shell_named": (c_void_p, [c_size_t, c_size_t]),
            "shell_static": (c_void_p, [c_char_p, c_size_t, c_size_t]), "shell_free": (None, [c_void_p]),
            "shell_info": (None, [c_void_p, ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t), ctypes.POINTER(c_int),
                                  ctypes.POINTER(c_void_p)]),
            "shell_to_host": (c_int, [c_void_p, c_void_p]), "shell_norm": (c_float, [c_void_p]),
            "shell_unary": (c_void_p, [c_char_p, c_void_p, c_int]),
            "shell_scalar": (c_void_p, [c_char_p, c_void_p, c_double, c_int]),
            "shell_binary": (c_void_p, [c_char_p, c_void_p, c_void_p, c_int, c_int]),
            "shell_inplace": (c_int, [c_char_p, c_void_p, c_void_p, c_double]),
            "shell_slice": (c_void_p, [c_void_p, c_size_t, c_size_t, c_size_t, c_size_t]),
            "shell_slice_set": (c_int, [c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p]),
            "shell_stack": (c_void_p, [c_int, ctypes.POINTER(c_void_p), c_int]),
            "shell_expr_softplus5": (c_void_p, [c_void_p, c_void_p, c_double]),
            "shell_expr_testbasic": (c_void_p, [c_void_p, c_void_p]),
            "shell_logistic_grad": (c_void_p, [c_void_p, c_void_p]), "shell_logistic_loss": (c_void_p, [c_void_p, c_void_p]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        if L.shell_init(0) != 0:
            raise RuntimeError("shell_init: " + L.shell_last_error().decode())
        _lib = L
    return _lib


def _wrap(h):
    if not h:
        L = lib()
        msg = L.shell_last_error().decode()
        raise (ShellShapeError if L.shell_last_error_is_shape() else RuntimeError)(msg)
    return SM(_handle=h)


def _check(rc):
    if rc:
        L = lib()
        msg = L.shell_last_error().decode()
        raise (ShellShapeError if L.shell_last_error_is_shape() else RuntimeError)(msg)


class SM:
    """a heap ``Matrix<CUDAfloat>`` of the C++ shell"""

    def __init__(self, host=None, *, _handle=None):
        if _handle is not None:
            self.h = _handle
            return
        a = np.asfortranarray(np.asarray(host, dtype=np.float32))
        if a.ndim == 1:
            a = np.asfortranarray(a.reshape(-1, 1))
        self.h = None
        self.h = _wrap(lib().shell_from_host(a.ctypes.data, a.shape[0], a.shape[1])).release()

    def release(self):
        h, self.h = self.h, None
        return h

    def __del__(self):
        try:
            if self.h:
                lib().shell_free(self.h)
        except Exception:  # noqa: BLE001
            pass

    # ---- construction
    @staticmethod
    def named(r, c):
        return _wrap(lib().shell_named(r, c))

    @staticmethod
    def ones_(r, c):
        return _wrap(lib().shell_static(b"ones", r, c))

    @staticmethod
    def zeros_(r, c):
        return _wrap(lib().shell_static(b"zeros", r, c))

    # ---- info
    def info(self):
        r, c, t, p = c_size_t(), c_size_t(), c_int(), c_void_p()
        lib().shell_info(self.h, ctypes.byref(r), ctypes.byref(c), ctypes.byref(t), ctypes.byref(p))
        return r.value, c.value, bool(t.value), p.value

    def num_row(self):
        return self.info()[0]

    def num_col(self):
        return self.info()[1]

    def get_transpose(self):
        return self.info()[2]

    def data(self):
        return self.info()[3]

    def to_host(self):
        r, c, _, _ = self.info()
        out = np.empty((r, c), dtype=np.float32, order="F")
        _check(lib().shell_to_host(self.h, out.ctypes.data))
        return out

    def norm(self):
        return float(lib().shell_norm(self.h))

    # ---- members / operators (lvalue forms; rvalue forms through the module functions' move=True)
    def T(self):
        return unary("T", self)

    def copy(self):
        return unary("copy", self)

    def dot(self, B):
        return binary("mul", self, B)

    def slice(self, r0, r1, c0, c1, M=None):
        if M is None:
            return _wrap(lib().shell_slice(self.h, r0, r1, c0, c1))
        _check(lib().shell_slice_set(self.h, r0, r1, c0, c1, M.h))
        return None

    def rows(self, r0, r1, M=None):
        return self.slice(r0, r1, 0, self.num_col(), M)

    def columns(self, c0, c1, M=None):
        return self.slice(0, self.num_row(), c0, c1, M)

    def zeros(self):
        _check(lib().shell_inplace(b"zeros", self.h, None, 0.0))

    def ones(self):
        _check(lib().shell_inplace(b"ones", self.h, None, 0.0))

    def __mul__(self, r):
        return binary("mul", self, r) if isinstance(r, SM) else scalar("mul", self, r)

    def __rmul__(self, l):
        return scalar("rmul", self, l)

    def __add__(self, r):
        return binary("add", self, r) if isinstance(r, SM) else scalar("add", self, r)

    def __radd__(self, l):
        return scalar("radd", self, l)

    def __sub__(self, r):
        return binary("sub", self, r) if isinstance(r, SM) else scalar("sub", self, r)

    def __rsub__(self, l):
        return scalar("rsub", self, l)

    def __neg__(self):
        return unary("neg", self)

    def __truediv__(self, r):
        return binary("div", self, r) if isinstance(r, SM) else scalar("div", self, r)

    def __rtruediv__(self, l):
        return scalar("rdiv", self, l)

    def __iadd__(self, r):
        if isinstance(r, SM):
            _check(lib().shell_inplace(b"iadd", self.h, r.h, 0.0))
        else:
            _check(lib().shell_inplace(b"iadds", self.h, None, float(r)))
        return self

    def __isub__(self, r):
        if isinstance(r, SM):
            _check(lib().shell_inplace(b"isub", self.h, r.h, 0.0))
        else:
            _check(lib().shell_inplace(b"isubs", self.h, None, float(r)))
        return self


def unary(op, M, move=False):
    return _wrap(lib().shell_unary(op.encode(), M.h, int(move)))


def scalar(op, M, s, move=False):
    return _wrap(lib().shell_scalar(op.encode(), M.h, float(s), int(move)))


def binary(op, A, B, move_a=False, move_b=False):
    return _wrap(lib().shell_binary(op.encode(), A.h, B.h, int(move_a), int(move_b)))


def exp(M, move=False):
    return unary("exp", M, move)


def log(M):
    return unary("log", M)


def tanh(M, move=False):
    return unary("tanh", M, move)


def d_tanh(M, move=False):
    return unary("d_tanh", M, move)


def square(M, move=False):
    return unary("square", M, move)


def sqrt(M, move=False):
    return unary("sqrt", M, move)


def relu(M):
    return unary("relu", M)


def d_relu(M):
    return unary("d_relu", M)


def colmax(M):
    return unary("colmax", M)


def sum(M, dim):  # noqa: A001 - the reference's name
    return unary("sum0" if dim == 0 else "sum1", M)


def hadmd(A, B, move_a=False, move_b=False):
    return binary("hadmd", A, B, move_a, move_b)


def fill(M, a):
    _check(lib().shell_inplace(b"fill", M.h, None, float(a)))
    return M


def _stack(vertical, mats):
    arr = (c_void_p * len(mats))(*[m.h for m in mats])
    return _wrap(lib().shell_stack(vertical, arr, len(mats)))


def hstack(mats):
    return _stack(0, mats)


def vstack(mats):
    return _stack(1, mats)


def expr_softplus5(A, B, div):
    return _wrap(lib().shell_expr_softplus5(A.h, B.h, float(div)))


def expr_testbasic(A, B):
    return _wrap(lib().shell_expr_testbasic(A.h, B.h))


def logistic_grad(X, Y):
    return _wrap(lib().shell_logistic_grad(X.h, Y.h))


def logistic_loss(X, Y):
    return _wrap(lib().shell_logistic_loss(X.h, Y.h))

"""juzhen_b200 -- B200-native (sm_100a) backend for Juzhen's ``Matrix<CUDAfloat>`` compute layer.

The product is ``libjz_b200.so`` (hand-written CUDA behind the C ABI in ``include/jz_b200.h``)
plus the drop-in C++ class in ``juzhen_b200/cpp``.  This Python package is the ctypes binding
used by the tests and the benchmark; it contains no compute and no CPU fallback.
"""
from . import _lib
from ._lib import JzError, JzShapeError, lib
from .matrix import (CM, chain, colmax, d_relu, d_tanh, exp, fill, get_stream, hadmd, hstack, log, relu,
                     set_stream, softmax_ce_grad, softmax_cols, sqrt, square, sum, sync, tanh, vstack)

__all__ = ["CM", "JzError", "JzShapeError", "lib", "chain", "colmax", "d_relu", "d_tanh", "exp", "fill",
           "get_stream", "hadmd", "hstack", "log", "relu", "set_stream", "softmax_ce_grad", "softmax_cols",
           "sqrt", "square", "sum", "sync", "tanh", "vstack"]

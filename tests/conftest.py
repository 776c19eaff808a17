import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """fixtures generated from the UNMODIFIED reference by scripts/make_golden.py"""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def jz():
    """the product library; GPU tests must never silently fall back to anything else"""
    import juzhen_b200
    rc = juzhen_b200.lib().jz_init(0)
    assert rc == 0, juzhen_b200.lib().jz_last_error()
    return juzhen_b200


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def ulp_dist(a, b):
    """distance in units in the last place between two float32 arrays (NaN == NaN, +0 == -0)"""
    a = np.ascontiguousarray(a, dtype=np.float32).ravel()
    b = np.ascontiguousarray(b, dtype=np.float32).ravel()
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ka = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    kb = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    d = np.abs(ka - kb)
    both_nan = np.isnan(a) & np.isnan(b)
    d[both_nan] = 0
    return d


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

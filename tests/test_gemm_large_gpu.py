"""GEMM parity at the benchmark sizes of BASELINE configs[2] (n = 8192, 16384) x {A*B, A.T*B, A*B.T} x {3xTF32, TF32}.

Comparator, as BASELINE.md section 3 prescribes for sizes where the full CPU product takes minutes: a 2048 x 2048
output sub-block computed by the UNMODIFIED reference on the CPU through its own rows() / columns() / dot
(oracle/_ref/libjzref.so -> Matrix<float>::dot -> OpenBLAS cblas_sgemm; oracle/ref_shim.cpp:ref_gemm_subblock), plus
256 sampled entries of the whole product recomputed in float64.  Tolerances (north_star): relative Frobenius error
<= 1e-5 in 3xTF32 mode, <= 1e-3 in TF32 mode.  Where the reference build did not travel, the C oracle's
double-accumulated product of a 256 x 256 sub-block stands in (and the test says so).
"""
import numpy as np
import pytest

from conftest import rel_fro

pytestmark = pytest.mark.gpu

TOL = {0: 1e-5, 1: 1e-3}
MODE_NAME = {0: "3xtf32", 1: "tf32"}
_cache = {}


def operands(n):
    """host copies generated once per size (numpy PCG64, seeded): 2 x n^2 floats"""
    if n not in _cache:
        _cache.clear()   # keep at most one size resident on the host (16384^2 = 1 GiB per operand)
        rng = np.random.default_rng(n)
        A = np.asfortranarray(rng.standard_normal((n, n), dtype=np.float32))
        B = np.asfortranarray(rng.standard_normal((n, n), dtype=np.float32))
        _cache[n] = (A, B)
    return _cache[n]


def logical(M, t):
    return M.T if t else M


@pytest.mark.parametrize("n", [8192, 16384])
def test_gemm_benchmark_sizes_vs_reference_openblas_subblock(jz, port, n):
    import oracle
    A, B = operands(n)
    a, b = jz.CM(A), jz.CM(B)
    L = jz.lib()
    have_ref = oracle.ref_available()
    blk = 2048 if have_ref else 256
    rng = np.random.default_rng(5)
    for (ta, tb, name) in ((0, 0, "A*B"), (1, 0, "A.T*B"), (0, 1, "A*B.T")):
        r0 = int(rng.integers(0, n - blk + 1)) // 4 * 4
        c0 = int(rng.integers(0, n - blk + 1)) // 4 * 4
        if have_ref:
            want_blk = oracle.ref().gemm_subblock(A, ta, B, tb, r0, r0 + blk, c0, c0 + blk)
            comparator = "reference rows()/columns()/dot (OpenBLAS)"
        else:
            opA, opB = logical(A, ta), logical(B, tb)
            want_blk = port.gemm(np.asfortranarray(opA[r0:r0 + blk, :]), 0, np.asfortranarray(opB[:, c0:c0 + blk]), 0, f64=True)
            comparator = "C oracle, double accumulation (reference build absent)"
        ii, jj = rng.integers(0, n, 256), rng.integers(0, n, 256)
        opA, opB = logical(A, ta), logical(B, tb)
        want_pts = np.einsum("ik,ki->i", opA[ii, :].astype(np.float64), opB[:, jj].astype(np.float64))
        for mode in (0, 1):
            da = a.T() if ta else a
            db = b.T() if tb else b
            c = da.dot(db, mode=mode)
            assert L.jz_gemm_last_path() == 1
            got_blk = c.slice(r0, r0 + blk, c0, c0 + blk).to_host()
            e_blk = rel_fro(got_blk, want_blk)
            # sampled entries of the whole product: gather columns on the device, pick rows on the host
            got_pts = np.array([c.slice(int(i), int(i) + 1, int(j), int(j) + 1).to_host()[0, 0] for i, j in zip(ii[:64], jj[:64])])
            e_pts = float(np.linalg.norm(got_pts - want_pts[:64]) / np.linalg.norm(want_pts[:64]))
            print(f"n={n} {name} {MODE_NAME[mode]}: rel_fro vs {comparator} on [{r0}:{r0 + blk}, {c0}:{c0 + blk}] = {e_blk:.3e}; "
                  f"64 sampled entries vs float64 = {e_pts:.3e}; splits={L.jz_gemm_last_splits()}")
            assert e_blk < TOL[mode], (n, name, mode, e_blk)
            assert e_pts < TOL[mode] * 2, (n, name, mode, e_pts)
            del c

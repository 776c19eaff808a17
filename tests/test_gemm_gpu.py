"""GPU parity tests for the GEMM behind operator* (SURVEY 8a row a17), through the C-ABI.

Tolerances (north_star): relative Frobenius error <= 1e-5 in 3xTF32 mode (the default) and
<= 1e-3 in TF32 mode, against the reference's CPU result (OpenBLAS fixture) and against the
oracle's double-accumulated product.  All four op(A), op(B) combinations, ragged tiles,
odd leading dimensions (the n = 1001 case of tests/testEigen.cu), alpha/beta, fused epilogue.
"""
import os

import numpy as np
import pytest

from conftest import bits, rel_fro

pytestmark = pytest.mark.gpu

TOL = {"3xtf32": 1e-5, "tf32": 1e-3, "fp32": 1e-5}


def F(a):
    return np.asfortranarray(a, dtype=np.float32)


def operands(jz, P, Q, ta, tb):
    a = jz.CM(F(P.T)).T() if ta else jz.CM(P)
    b = jz.CM(F(Q.T)).T() if tb else jz.CM(Q)
    return a, b


@pytest.mark.parametrize("name", ["g1", "g2", "g3", "g4"])
def test_gemm_vs_reference_fixture(jz, golden, name):
    P, Q = golden[f"gemm_{name}_A"], golden[f"gemm_{name}_B"]
    for ta in (0, 1):
        for tb in (0, 1):
            a, b = operands(jz, P, Q, ta, tb)
            got = (a * b).to_host()
            assert rel_fro(got, golden[f"gemm_{name}_blas_{ta}{tb}"]) < 1e-5, (name, ta, tb)


def test_gemm_odd_leading_dimension(jz, golden):
    P, Q = golden["gemm_odd_A"], golden["gemm_odd_B"]
    got = (jz.CM(P) * jz.CM(F(Q.T)).T()).to_host()
    assert rel_fro(got, golden["gemm_odd_blas_01"]) < 1e-5


def test_gemm_shape_error(jz):
    with pytest.raises(ValueError):
        jz.CM.ones_(3, 4) * jz.CM.ones_(3, 4)


@pytest.mark.parametrize("mode", ["3xtf32", "tf32", "fp32"])
@pytest.mark.parametrize("shape", [(256, 256, 256), (384, 300, 520), (130, 1000, 70), (640, 320, 384), (1024, 512, 768),
                                   (515, 2049, 257), (2048, 2048, 48), (20, 3000, 3000), (3000, 40, 3000)])
def test_gemm_all_flags_vs_oracle(jz, port, mode, shape):
    m, k, n = shape
    rng = np.random.default_rng(m * 7 + k * 3 + n)
    P, Q = F(rng.standard_normal((m, k))), F(rng.standard_normal((k, n)))
    truth = port.gemm(P, 0, Q, 0, f64=True)
    for ta in (0, 1):
        for tb in (0, 1):
            a, b = operands(jz, P, Q, ta, tb)
            got = a.dot(b, mode=jz._lib.GEMM_MODES[mode]).to_host()
            err = rel_fro(got, truth)
            path = jz.lib().jz_gemm_last_path()
            print(f"gemm {mode} {shape} ta={ta} tb={tb}: rel_fro={err:.3e} path={path}")
            assert err < TOL[mode], (mode, shape, ta, tb, err)
            if mode != "fp32" and ((min(m, n) >= 64 and m * n * k >= (1 << 22)) or m * n * k > (1 << 26)):
                assert path == 1, "expected the tcgen05 kernel"
            elif m * n * k <= (1 << 26):
                assert path == 4, "expected the small-product kernel"


@pytest.mark.parametrize("shape", [(1024, 784, 32), (128, 1024, 32), (10, 128, 32), (1024, 32, 784), (784, 1024, 32),
                                   (128, 10, 32), (33, 257, 5), (1000, 31, 1000), (2, 4096, 500), (7, 3, 2)])
def test_gemm_small_products_training_step_shapes(jz, port, shape):
    """the products of one demo_mnist step at batch 32 and other latency-bound shapes: all four
    flag combinations, ragged edges, leading dimensions that defeat 128-bit loads, alpha/beta"""
    m, k, n = shape
    rng = np.random.default_rng(m + 3 * k + 7 * n)
    P, Q = F(rng.standard_normal((m, k))), F(rng.standard_normal((k, n)))
    truth = port.gemm(P, 0, Q, 0, f64=True)
    L = jz.lib()
    for ta in (0, 1):
        for tb in (0, 1):
            a, b = operands(jz, P, Q, ta, tb)
            got = a.dot(b, mode=0).to_host()
            assert L.jz_gemm_last_path() == (1 if min(m, n) >= 64 and k >= 32 and m * n * k >= (1 << 22) else 4)
            assert rel_fro(got, truth) < 1e-5, (shape, ta, tb)
    C0 = F(rng.standard_normal((m, n)))
    a, b, c = jz.CM(P), jz.CM(Q), jz.CM(C0)
    jz._lib.check(L.jz_gemm(0, 0, m, n, k, 0.75, a.ptr, m, b.ptr, k, -0.5, c.ptr, m, 0, None))
    assert rel_fro(c.to_host(), 0.75 * truth.astype(np.float64) - 0.5 * C0) < 1e-5


def test_gemm_tf32_is_actually_tf32_and_3x_is_fp32_grade(jz, port):
    """the two modes must differ the way the README says (~1e-3 vs fp32-grade)"""
    rng = np.random.default_rng(3)
    P, Q = F(rng.standard_normal((512, 512))), F(rng.standard_normal((512, 512)))
    truth = port.gemm(P, 0, Q, 0, f64=True)
    a, b = jz.CM(P), jz.CM(Q)
    e3 = rel_fro(a.dot(b, mode=0).to_host(), truth)
    e1 = rel_fro(a.dot(b, mode=1).to_host(), truth)
    ef = rel_fro(a.dot(b, mode=2).to_host(), truth)
    print(f"rel_fro: 3xtf32={e3:.3e} tf32={e1:.3e} fp32-simt={ef:.3e}")
    assert e3 < 2e-6 and 1e-5 < e1 < 1e-3 and ef < 1e-6


def test_gemm_alpha_beta(jz, port):
    rng = np.random.default_rng(9)
    m, k, n = 300, 260, 280
    P, Q, C0 = F(rng.standard_normal((m, k))), F(rng.standard_normal((k, n))), F(rng.standard_normal((m, n)))
    truth = 0.75 * port.gemm(P, 0, Q, 0, f64=True).astype(np.float64) - 0.5 * C0
    L = jz.lib()
    for mode in (0, 2):
        a, b, c = jz.CM(P), jz.CM(Q), jz.CM(C0)
        jz._lib.check(L.jz_gemm(0, 0, m, n, k, 0.75, a.ptr, m, b.ptr, k, -0.5, c.ptr, m, mode, None))
        assert rel_fro(c.to_host(), truth) < 1e-5


def test_gemm_fused_epilogue_equals_separate_kernels(jz, golden):
    """config 1: log(exp(A*B/n)+1)/5 as a GEMM epilogue == GEMM then the elementwise chain,
    and both match the reference's CPU result."""
    A, B = golden["c1_A"], golden["c1_B"]
    n = A.shape[0]
    a, b = jz.CM(A), jz.CM(B)
    sep = jz.log(jz.exp((a * b) / float(n)) + 1.0) / 5.0
    steps = [("affine", float(np.float32(1.0 / n)), 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",),
             ("affine", float(np.float32(1.0 / 5.0)), 0.0)]
    arr, ns = jz._lib.make_steps(steps)
    fused = jz.CM.empty("f", n, n)
    jz._lib.check(jz.lib().jz_gemm_chain(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, fused.ptr, n, arr, ns, -1, None))
    assert np.array_equal(bits(sep.to_host()), bits(fused.to_host()))
    assert rel_fro(fused.to_host(), golden["c1_out"]) < 1e-5


def test_gemm_4096_sampled_block_vs_oracle(jz, port):
    """benchmark-size product: a 4096 x 64 column block recomputed by the oracle in double"""
    n = 4096
    a, b = jz.CM.randn(n, n, seed=1), jz.CM.randn(n, n, seed=2)
    for mode, tol in ((0, 1e-5), (1, 1e-3)):
        c = a.dot(b, mode=mode)
        assert jz.lib().jz_gemm_last_path() == 1
        A = a.to_host()
        Bc = b.columns(1000, 1064).to_host()
        truth = port.gemm(A, 0, Bc, 0, f64=True)
        assert rel_fro(c.columns(1000, 1064).to_host(), truth) < tol
        # A^T path on the same data: (A^T)^T B == A B
        at = jz.CM.empty("at", n, n)
        jz._lib.check(jz.lib().jz_copy2d(at.ptr, n, a.ptr, n, n, n, 1, None))
        c2 = at.T().dot(b, mode=mode)
        assert rel_fro(c2.columns(1000, 1064).to_host(), truth) < tol


@pytest.mark.parametrize("seed", range(6))
def test_gemm_fuzz_shapes_offsets_strides(jz, seed):
    """dispatch boundaries (small-product / tensor / SIMT / rank-1 / k = 0), all flag combinations, leading dimensions
    with padding, base pointers that are not 16-byte aligned, alpha/beta: 25 random cases per seed against float64"""
    rng = np.random.default_rng(1000 + seed)
    L = jz.lib()
    for case in range(25):
        big = rng.random() < 0.3
        hi = 700 if big else 90
        m, n = int(rng.integers(1, hi)), int(rng.integers(1, hi))
        k = int(rng.integers(0, hi))
        ta, tb = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        pa, pb, pc = (int(rng.integers(0, 6)) for _ in range(3))         # ld padding
        oa, ob, oc = (int(rng.integers(0, 4)) for _ in range(3))         # base offsets in floats
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        lda, ldb, ldc = max(ar, 1) + pa, max(br, 1) + pb, m + pc
        A = rng.standard_normal(oa + lda * max(ac, 1)).astype(np.float32)
        B = rng.standard_normal(ob + ldb * max(bc, 1)).astype(np.float32)
        C0 = rng.standard_normal(oc + ldc * n).astype(np.float32)
        alpha = float(rng.choice([1.0, -0.5, 2.0]))
        beta = float(rng.choice([0.0, 0.0, 1.0, -0.25]))
        dA = jz.CM(np.asfortranarray(A.reshape(-1, 1)))
        dB = jz.CM(np.asfortranarray(B.reshape(-1, 1)))
        dC = jz.CM(np.asfortranarray(C0.reshape(-1, 1)))
        mode = int(rng.choice([0, 0, 1, 2]))
        jz._lib.check(L.jz_gemm(ta, tb, m, n, k, alpha, dA.ptr + 4 * oa, lda, dB.ptr + 4 * ob, ldb, beta, dC.ptr + 4 * oc, ldc, mode, None))
        got = dC.to_host().ravel()
        Am = A[oa:oa + lda * ac].reshape(lda, ac, order="F")[:ar, :].astype(np.float64) if ac else np.zeros((ar, 0))
        Bm = B[ob:ob + ldb * bc].reshape(ldb, bc, order="F")[:br, :].astype(np.float64) if bc else np.zeros((br, 0))
        opA = Am.T if ta else Am
        opB = Bm.T if tb else Bm
        Cm = C0[oc:oc + ldc * n].reshape(ldc, n, order="F").astype(np.float64)
        want = Cm.copy()
        want[:m, :] = alpha * (opA @ opB) + beta * Cm[:m, :]
        gotm = got[oc:oc + ldc * n].reshape(ldc, n, order="F")
        tol = 1e-3 if mode == 1 else 1e-5
        scale = np.abs(alpha) * (np.abs(opA) @ np.abs(opB)) + np.abs(beta * Cm[:m, :]) + 1e-30
        desc = (seed, case, m, n, k, ta, tb, lda, ldb, ldc, oa, ob, oc, alpha, beta, mode, L.jz_gemm_last_path())
        assert np.all(np.abs(gotm[:m, :] - want[:m, :]) <= tol * scale * 4 + 1e-6), desc
        # nothing outside the m x n window moved: padding rows, and the floats in front of the base pointer
        assert np.array_equal(gotm[m:, :], Cm[m:, :].astype(np.float32)), desc
        assert np.array_equal(got[:oc], C0[:oc]), desc


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
@pytest.mark.parametrize("shape", [(1024, 1024, 1024), (512, 4096, 512), (4096, 2048, 32), (300, 5000, 200),
                                   (1024, 8192, 784), (2304, 1024, 2304)])
def test_gemm_split_k_units(jz, port, mode, shape):
    """shapes whose tile grid leaves a partial last wave (or less than one wave): the tail tiles are split along k
    into units that meet through workspace (jz_gemm_tc.cuh).  All flag combinations, alpha/beta, fused chain,
    bitwise run-to-run determinism whatever the arrival order of the units."""
    m, k, n = shape
    rng = np.random.default_rng(m + 5 * k + 11 * n)
    P, Q = F(rng.standard_normal((m, k))), F(rng.standard_normal((k, n)))
    truth = port.gemm(P, 0, Q, 0, f64=True).astype(np.float64)
    L = jz.lib()
    md = jz._lib.GEMM_MODES[mode]
    for ta in (0, 1):
        for tb in (0, 1):
            a, b = operands(jz, P, Q, ta, tb)
            got = a.dot(b, mode=md).to_host()
            assert L.jz_gemm_last_path() == 1
            if mode == "3xtf32":   # (in TF32 mode a k-block is 3x cheaper and the cost model may keep a small tail whole)
                assert L.jz_gemm_last_splits() > 1, ("expected split-K units", shape, L.jz_gemm_last_splits())
            err = rel_fro(got, truth)
            print(f"split-K {mode} {shape} ta={ta} tb={tb} splits={L.jz_gemm_last_splits()} rel_fro={err:.3e}")
            assert err < TOL[mode], (mode, shape, ta, tb, err)
            again = a.dot(b, mode=md).to_host()
            assert np.array_equal(bits(got), bits(again)), "split-K result must not depend on unit arrival order"
    C0 = F(rng.standard_normal((m, n)))
    a, b, c = jz.CM(P), jz.CM(Q), jz.CM(C0)
    jz._lib.check(L.jz_gemm(0, 0, m, n, k, 0.75, a.ptr, m, b.ptr, k, -0.5, c.ptr, m, md, None))
    assert rel_fro(c.to_host(), 0.75 * truth - 0.5 * C0) < TOL[mode]
    # fused chain on the split path == chain applied afterwards (3xTF32 only: same accumulators, same roundings)
    if mode == "3xtf32":
        steps = [("affine", float(np.float32(1.0 / k)), 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",)]
        arr, ns = jz._lib.make_steps(steps)
        fused = jz.CM.empty("f", m, n)
        jz._lib.check(L.jz_gemm_chain(0, 0, m, n, k, 1.0, a.ptr, m, b.ptr, k, fused.ptr, m, arr, ns, md, None))
        x = truth / np.float32(1.0 * k)
        want = np.log(np.exp(x) + 1.0)
        assert rel_fro(fused.to_host(), want) < 1e-5


@pytest.mark.parametrize("shape", [(1024, 1024, 1024, 0, 0, ""), (1024, 1024, 1024, 1, 1, ""), (1024, 1024, 1024, 0, 0, "JZ_GEMM_TS=0"),
                                   (1024, 1024, 1024, 1, 0, "JZ_GEMM_TS=0"), (1000, 1100, 900, 0, 1, ""), (1000, 1100, 900, 0, 1, "JZ_GEMM_TS=0"),
                                   (1280, 2048, 768, 0, 0, ""), (8192, 8192, 32, 0, 0, ""), (8192, 4100, 40, 1, 0, ""),
                                   (1536, 1536, 1536, 0, 0, "JZ_GEMM_TS=0"), (1024, 60000, 784, 0, 1, "")])
def test_gemm_cluster_split_same_bits_as_workspace_form(shape, tmp_path):
    """Products of few tiles split every tile along k; the units of a tile then form one thread-block cluster and exchange
    their partial tiles through distributed shared memory (jz_gemm_tc.cuh, cluster_split_send / _reduce): clusters of 2
    single CTAs (TMEM-A tiles), of 2 CTA pairs, of 4 CTA pairs.  The partials are added in split order either way, so the
    result must have the SAME BITS as the workspace + ticket form (JZ_GEMM_CLUSTER_SPLIT=0, read once per process: both
    runs are fresh processes), with alpha / beta and with a fused program."""
    import subprocess
    import sys
    m, k, n, ta, tb, extra = shape
    extra_env = dict(kv.split("=") for kv in extra.split()) if extra else {}
    here = os.path.dirname(os.path.abspath(__file__))
    res = {}
    for name, env in (("cluster", extra_env), ("workspace", {**extra_env, "JZ_GEMM_CLUSTER_SPLIT": "0"})):
        out = str(tmp_path / f"{name}.npz")
        r = subprocess.run([sys.executable, os.path.join(here, "_gemm_dump.py"), *map(str, (m, k, n, ta, tb, 0, 7)), out],
                           env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res[name] = np.load(out)
    assert int(res["workspace"]["cluster_split"]) == 0
    assert int(res["workspace"]["splits"]) > 1, "the shape is meant to split along k"
    print(f"cluster split {shape}: splits={int(res['cluster']['splits'])} cluster_form={int(res['cluster']['cluster_split'])}")
    if (m, k, n) in ((1024, 1024, 1024), (8192, 8192, 32)) and not extra:   # (16 clusters of 8 full-SM CTAs are not all resident on 148 SMs)
        assert int(res["cluster"]["cluster_split"]) == 1, ("expected the cluster form", shape, int(res["cluster"]["splits"]))
    for key in ("axpby", "chain"):
        assert np.array_equal(bits(res["cluster"][key]), bits(res["workspace"][key])), (shape, key)
    rng = np.random.default_rng(7)
    P = rng.standard_normal((k, m) if ta else (m, k)).astype(np.float32)
    Q = rng.standard_normal((n, k) if tb else (k, n)).astype(np.float32)
    C0 = rng.standard_normal((m, n)).astype(np.float32)
    if m * n * k <= 1 << 32:
        truth = (P.T if ta else P).astype(np.float64) @ (Q.T if tb else Q).astype(np.float64)
        assert rel_fro(res["cluster"]["axpby"], 0.75 * truth - 0.5 * C0) < 1e-5


@pytest.mark.parametrize("shape", [(8192, 1024, 32), (4096, 520, 48), (2048, 777, 64), (128, 4096, 1024), (100, 3000, 300),
                                   (96, 2100, 2000), (16384, 40, 64), (128, 64, 8192), (16384, 96, 48), (1280, 1280, 1280)])
def test_gemm_narrow_or_short_output_tmem_a_variant(jz, port, shape):
    """3xTF32 products with n <= 64 or m <= 128 run on single-CTA tiles whose A operand (hi and lo parts) is staged in
    TENSOR MEMORY by the transform warps (jz_gemm_tc.cuh, MODE_XFORM_TS): both source layouts of A (the transform reads the
    swizzled tile per row), k tails, split-K, alpha/beta, fused chain, inf propagation, and equality of the result's
    accuracy class with the shared-memory form."""
    m, k, n = shape
    rng = np.random.default_rng(3 * m + 7 * k + 13 * n)
    P, Q = F(rng.standard_normal((m, k))), F(rng.standard_normal((k, n)))
    truth = port.gemm(P, 0, Q, 0, f64=True).astype(np.float64)
    L = jz.lib()
    for ta in (0, 1):
        for tb in (0, 1):
            a, b = operands(jz, P, Q, ta, tb)
            got = a.dot(b, mode=0).to_host()
            assert L.jz_gemm_last_path() == 1
            err = rel_fro(got, truth)
            print(f"TMEM-A {shape} ta={ta} tb={tb} splits={L.jz_gemm_last_splits()} rel_fro={err:.3e}")
            assert err < 3e-6, (shape, ta, tb, err)     # fp32 grade, well inside the 1e-5 bound
            again = a.dot(b, mode=0).to_host()
            assert np.array_equal(bits(got), bits(again))
    C0 = F(rng.standard_normal((m, n)))
    a, b, c = jz.CM(P), jz.CM(Q), jz.CM(C0)
    jz._lib.check(L.jz_gemm(0, 0, m, n, k, 0.75, a.ptr, m, b.ptr, k, -0.5, c.ptr, m, 0, None))
    assert rel_fro(c.to_host(), 0.75 * truth - 0.5 * C0) < 1e-5
    steps = [("affine", float(np.float32(1.0 / k)), 0.0), ("tanh",)]
    arr, ns = jz._lib.make_steps(steps)
    fused = jz.CM.empty("f", m, n)
    jz._lib.check(L.jz_gemm_chain(0, 0, m, n, k, 1.0, a.ptr, m, b.ptr, k, fused.ptr, m, arr, ns, 0, None))
    assert rel_fro(fused.to_host(), np.tanh(truth / np.float32(1.0 * k))) < 1e-5
    # an infinite entry of A gives +-inf / nan exactly where the fp32 product does (hi carries it, lo is forced to 0)
    P2 = P.copy()
    P2[m // 2, k // 3] = np.inf
    # (columns positive with a non-zero low part: inf * lo(B) is nan when lo(B) is exactly 0, one fp32 value in 8192 --
    # the known corner of every split-precision product, cuBLAS's 3xTF32 included)
    Qp = ((np.abs(Q) + 0.5).astype(np.float32).view(np.uint32) | np.uint32(0x800)).view(np.float32)
    got = jz.CM(P2).dot(jz.CM(F(Qp)), mode=0).to_host()
    assert np.all(np.isposinf(got[m // 2, :])), "row with +inf times positive columns must be +inf"
    assert np.all(np.isfinite(np.delete(got, m // 2, axis=0)))


def test_gemm_wide_output_tiny_k(jz):
    """products with a very wide output and tiny m, k (w^T X, ones(1, d) * X ...) must not hit a grid.y
    limit on the small-product / SIMT kernels"""
    rng = np.random.default_rng(77)
    L = jz.lib()
    for (m, k, n) in [(1, 64, 600_000), (16, 2, 1_000_000), (3, 40, 4_300_000)]:
        W = F(rng.standard_normal((m, k)))
        X = F(rng.standard_normal((k, n)))
        for mode in (0, 2):
            got = jz.CM(W).dot(jz.CM(X), mode=mode).to_host()
            want = W.astype(np.float64) @ X.astype(np.float64)
            assert rel_fro(got, want) < 1e-5, (m, k, n, mode, L.jz_gemm_last_path())
    # the transposed problem: very tall output
    X = F(rng.standard_normal((2_200_000, 3)))
    W = F(rng.standard_normal((3, 5)))
    got = jz.CM(X).dot(jz.CM(W), mode=0).to_host()
    assert rel_fro(got, X.astype(np.float64) @ W.astype(np.float64)) < 1e-5


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
@pytest.mark.parametrize("seq,dh,heads,batch", [(64, 128, 2, 9), (128, 128, 4, 5), (256, 64, 2, 3), (512, 128, 1, 2), (200, 96, 2, 3)])
def test_gemm_strided_batched_attention_shapes_on_tensor_cores(jz, mode, seq, dh, heads, batch):
    """the two strided-batched products of TransformerLayer's attention (ml/layer.hpp:2896-2926) exactly as the
    reference calls cublasSgemmStridedBatched: scores_h = scale * Q_h^T K_h per head (operands are d_h-row slices of
    the (d_k x seq) per-sample matrices, lda = d_k) and H_h = V_h A_h^T written into head h's rows of H."""
    L = jz.lib()
    md = jz._lib.GEMM_MODES[mode]
    tol = TOL[mode]
    dk = dh * heads
    rng = np.random.default_rng(seq + dh + batch)
    Qm = rng.standard_normal((batch, seq, dk)).astype(np.float32)    # [b][token][feature] == column-major (d_k x seq) per sample
    Km = rng.standard_normal((batch, seq, dk)).astype(np.float32)
    Vm = rng.standard_normal((batch, seq, dk)).astype(np.float32)
    dQ, dK, dV = (jz.CM(np.asfortranarray(x.reshape(-1, 1))) for x in (Qm, Km, Vm))
    scores = jz.CM.empty("s", seq * seq * batch * heads, 1)
    H = jz.CM.empty("h", dk * seq * batch, 1)
    scale = float(1.0 / np.sqrt(dh))
    stride_qkv, stride_attn = dk * seq, seq * seq
    head_attn = stride_attn * batch
    for h in range(heads):
        jz._lib.check(L.jz_gemm_strided_batched(1, 0, seq, seq, dh, scale, dQ.ptr + 4 * h * dh, dk, stride_qkv,
                                                dK.ptr + 4 * h * dh, dk, stride_qkv, 0.0,
                                                scores.ptr + 4 * h * head_attn, seq, stride_attn, batch, md, None))
        if seq * seq * dh >= (1 << 21):
            assert L.jz_gemm_last_path() == 1, "attention members of >= 2^21 multiply-adds must run on the tcgen05 kernel"
    S = scores.to_host().ravel().reshape(heads, batch, seq, seq)      # [h][b][key j][query i] (column-major seq x seq)
    for h in range(heads):
        for b in range(batch):
            q = Qm[b, :, h * dh:(h + 1) * dh].astype(np.float64)      # (seq, dh)
            kk = Km[b, :, h * dh:(h + 1) * dh].astype(np.float64)
            want = scale * (q @ kk.T)                                  # (i, j)
            assert rel_fro(S[h, b].T, want) < tol, ("scores", h, b)
    # H_h = V_h (d_h x seq) * A_h^T, A_h stored (seq x seq) column-major; use the scores themselves as A
    for h in range(heads):
        jz._lib.check(L.jz_gemm_strided_batched(0, 1, dh, seq, seq, 1.0, dV.ptr + 4 * h * dh, dk, stride_qkv,
                                                scores.ptr + 4 * h * head_attn, seq, stride_attn, 0.0,
                                                H.ptr + 4 * h * dh, dk, stride_qkv, batch, md, None))
    Hh = H.to_host().ravel().reshape(batch, seq, dk)                  # [b][token][feature]
    for h in range(heads):
        for b in range(batch):
            v = Vm[b, :, h * dh:(h + 1) * dh].astype(np.float64)      # (seq, dh): V_h^T
            A = S[h, b].T.astype(np.float64)                           # A(i, j) logical
            want = (v.T @ A.T).T                                       # H_h(d, i) = sum_j V(d, j) A(i, j) -> stored [token i][d]
            assert rel_fro(Hh[b, :, h * dh:(h + 1) * dh], want) < tol, ("H", h, b)


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
@pytest.mark.parametrize("shape", [(64, 64, 128, 700, 1, 0), (100, 72, 96, 500, 0, 0), (128, 128, 64, 400, 0, 1), (256, 256, 128, 200, 1, 0),
                                   (300, 260, 64, 100, 0, 0), (512, 512, 128, 40, 1, 0), (128, 512, 512, 100, 0, 1), (64, 520, 32, 300, 1, 1)])
def test_gemm_strided_batched_walked_units(jz, mode, shape):
    """strided batches with more (member, tile) units than SMs are WALKED: one CTA (pair) per SM goes through the units with
    its barriers, tensor memory and TMA pipeline alive across them (jz_gemm_tc.cuh, WALK).  Single-CTA TMEM-A tiles and CTA
    pairs, ragged members, every operand layout, alpha / beta (beta != 0 takes the general epilogue, beta == 0 the compact
    one), bitwise run-to-run determinism, and the same bits as the one-CTA-per-unit launch of a batch too small to walk."""
    m, n, k, batch, ta, tb = shape
    md = jz._lib.GEMM_MODES[mode]
    L = jz.lib()
    rng = np.random.default_rng(m + 3 * n + 7 * k + batch)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = rng.standard_normal((batch, ac, ar)).astype(np.float32)      # member i = column-major (ar x ac)
    B = rng.standard_normal((batch, bc, br)).astype(np.float32)
    C0 = rng.standard_normal((batch, n, m)).astype(np.float32)
    dA, dB = (jz.CM(np.asfortranarray(x.reshape(-1, 1))) for x in (A, B))

    def run(alpha, beta, nb):
        dC = jz.CM(np.asfortranarray(C0.reshape(-1, 1)))
        jz._lib.check(L.jz_gemm_strided_batched(ta, tb, m, n, k, alpha, dA.ptr, ar, ar * ac, dB.ptr, br, br * bc, beta,
                                                dC.ptr, m, m * n, nb, md, None))
        return dC.to_host().ravel().reshape(batch, n, m), L.jz_gemm_last_walk(), L.jz_gemm_last_path()

    for alpha, beta in ((0.7, 0.0), (1.0, -0.5)):
        got, walked, path = run(alpha, beta, batch)
        assert path == 1 and walked == 1, (shape, path, walked)
        for i in list(range(0, batch, max(1, batch // 7))) + [batch - 1]:
            a = A[i].T.astype(np.float64)            # (ar, ac) logical
            b = B[i].T.astype(np.float64)
            want = alpha * ((a.T if ta else a) @ (b.T if tb else b)) + beta * C0[i].T
            assert rel_fro(got[i].T, want) < TOL[mode], (shape, alpha, beta, i)
        again, _, _ = run(alpha, beta, batch)
        assert np.array_equal(bits(got), bits(again))
        few, walked_few, path_few = run(alpha, beta, 3)     # 3 members: too few units to walk -> one CTA per unit, same arithmetic
        if path_few == 1:                                   # (tiny members of a small batch go to the small-product kernel)
            assert walked_few == 0
            assert np.array_equal(bits(got[:3]), bits(few[:3])), (shape, "walked vs one CTA per unit")

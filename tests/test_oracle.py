"""CPU tests (-m "not gpu"): the oracle is pinned before anything trusts it.

1. oracle/jz_oracle.c (the plain-C restatement) against the fixtures that the UNMODIFIED
   reference produced (tests/golden/ref_golden.npz, scripts/make_golden.py) -- bit-exact for
   elementwise, data movement, fixed-order reductions/GEMM; 1e-5 relative against the OpenBLAS
   summation order.
2. the reference's own golden vector tests/basic.testdata (values restated below).
3. when oracle/_ref/libjzref.so is present (build container), the live reference against the
   same fixtures, so a stale fixture file cannot hide.
"""
import numpy as np
import pytest

import oracle
from conftest import bits, rel_fro, ulp_dist

# /root/reference/tests/basic.testdata: header (2, 3, 0) + col-major payload
BASIC_TESTDATA = np.array([-26.99790382385254, -21.0, -23.0, -14.0, -17.0, -5.0], dtype=np.float32)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def F(a):
    return np.asfortranarray(a, dtype=np.float32)


def test_reference_golden_vector(golden):
    got = golden["basic_expr"].ravel(order="F")
    assert np.linalg.norm(got - BASIC_TESTDATA) < 1e-5          # the bar tests/testbasic.cu:16 uses
    assert same_bits(got, BASIC_TESTDATA)


def test_port_reproduces_golden_vector(port, golden):
    A, B = golden["basic_A"], golden["basic_B"]
    # log(exp(-A/B) + exp(hadmd(B,A))) - (A.T()*B).rows(0,2), restated op by op
    negA = port.affine(A.ravel(order="F"), -1.0, 0.0).reshape(2, 3, order="F")
    q = port.div(negA, 0, B, 0)
    e1 = port.unary("exp", q.ravel(order="F"))
    e2 = port.unary("exp", port.hadmd(B, 0, A, 0).ravel(order="F"))
    s = port.axpby(e1.reshape(2, 3, order="F"), 0, e2.reshape(2, 3, order="F"), 0, 1.0, 1.0)
    lg = port.unary("log", s.ravel(order="F")).reshape(2, 3, order="F")
    prod = port.gemm(A, 1, B, 0)                 # A.T() * B : 3 x 3
    rows = port.slice(prod, 0, 0, 2, 0, 3)
    out = port.axpby(lg, 0, rows, 0, 1.0, -1.0)
    assert same_bits(out.ravel(order="F"), BASIC_TESTDATA)


@pytest.mark.parametrize("op", ["exp", "tanh", "dtanh", "square", "relu", "drelu"])
def test_port_unary_bitexact(port, golden, op):
    assert same_bits(port.unary(op, golden["ew_x"]), golden["ew_" + op])


@pytest.mark.parametrize("op", ["log", "sqrt"])
def test_port_unary_positive_bitexact(port, golden, op):
    assert same_bits(port.unary(op, golden["ew_xp"]), golden["ew_" + op])


def test_port_scalar_ops_bitexact(port, golden):
    x, xp = golden["ew_x"], golden["ew_xp"]
    assert same_bits(port.affine(x, 1.7, -0.3), golden["ew_affine"])
    assert same_bits(port.affine(x, -1.0, 0.0), golden["ew_neg"])
    assert same_bits(port.div_scalar(x, 5.0), golden["ew_div5"])
    assert same_bits(port.div_scalar(x, 4096.0), golden["ew_div4096"])
    assert same_bits(port.eleminv(xp, 1.0), golden["ew_eleminv1"])
    assert same_bits(port.eleminv(xp, 3.0), golden["ew_eleminv3"])
    assert same_bits(port.chain_softplus5(golden["ew_xs"]), golden["ew_chain"])


@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_port_binary_bitexact(port, golden, ta, tb):
    A, B = golden["bin_A"], golden["bin_B"]
    a = F(A.T) if ta else A
    b = F(B.T) if tb else B
    assert same_bits(port.axpby(a, ta, b, tb, 1.5, -2.0), golden[f"bin_axpby_{ta}{tb}"])
    assert same_bits(port.hadmd(a, ta, b, tb), golden[f"bin_hadmd_{ta}{tb}"])
    assert same_bits(port.div(a, ta, b, tb), golden[f"bin_div_{ta}{tb}"])


def test_port_shape_errors(port, golden):
    A = golden["bin_A"]
    assert port.axpby(A, 0, A, 1, 1, 1) is None       # 37x53 vs 53x37 -> invalid_argument
    assert port.hadmd(A, 0, A, 1) is None
    assert port.gemm(A, 0, A, 0) is None
    assert port.stack(0, [(A, 0), (A, 1)]) is None


@pytest.mark.parametrize("name", ["r1", "r2", "r3", "r4"])
def test_port_reductions(port, golden, name):
    M = golden[f"red_{name}"]
    for ta in (0, 1):
        for dim in (0, 1):
            fixed = golden[f"red_{name}_sum_fixed_t{ta}d{dim}"]
            blas = golden[f"red_{name}_sum_blas_t{ta}d{dim}"]
            got = port.sum(M, ta, dim)
            assert same_bits(got, fixed)                                    # fixed order: bit-exact
            truth = port.sum(M, ta, dim, f64=True)
            scale = np.abs(M).sum(axis=(0 if (dim == 0) != bool(ta) else 1))
            assert np.all(np.abs(got - blas) <= 1e-5 * np.maximum(scale, 1e-30))   # OpenBLAS order
            assert np.all(np.abs(truth - blas) <= 1e-5 * np.maximum(scale, 1e-30))
            assert same_bits(port.reduce("max", M, ta, dim), golden[f"red_{name}_max_t{ta}d{dim}"])
            assert same_bits(port.reduce("stats", M, ta, dim), golden[f"red_{name}_stats_t{ta}d{dim}"])


def test_port_torchdump_fixture(port, golden):
    T = golden["torchdump_in"]
    for dim in (0, 1):
        assert same_bits(port.reduce("stats", T, 0, dim), golden[f"torchdump_sum_d{dim}"])
    # and against plain numpy, the role torch plays in tests/testElementwiseReduceTorch.py (tol 1e-6)
    assert np.allclose(port.reduce("stats", T, 0, 0)[0], T.sum(axis=0), atol=1e-6)
    assert np.allclose(port.reduce("stats", T, 0, 1)[:, 1], T.max(axis=1), atol=1e-6)


@pytest.mark.parametrize("name", ["g1", "g2", "g3", "g4"])
def test_port_gemm(port, golden, name):
    P, Q = golden[f"gemm_{name}_A"], golden[f"gemm_{name}_B"]
    for ta in (0, 1):
        for tb in (0, 1):
            a = F(P.T) if ta else P
            b = F(Q.T) if tb else Q
            got = port.gemm(a, ta, b, tb)
            assert same_bits(got, golden[f"gemm_{name}_fixed_{ta}{tb}"])      # BLAS-free order: bit-exact
            assert rel_fro(got, golden[f"gemm_{name}_blas_{ta}{tb}"]) < 1e-5  # OpenBLAS order
            assert rel_fro(port.gemm(a, ta, b, tb, f64=True), golden[f"gemm_{name}_blas_{ta}{tb}"]) < 1e-5


def test_port_data_movement_bitexact(port, golden):
    M, S, N1, N2 = golden["mv_M"], golden["mv_S"], golden["mv_N1"], golden["mv_N2"]
    assert same_bits(port.materialize(M, 1), golden["mv_T"])
    assert same_bits(port.slice(M, 0, 3, 20, 5, 30), golden["mv_slice"])
    assert same_bits(port.slice(M, 1, 3, 20, 5, 19), golden["mv_sliceT"])
    assert same_bits(port.slice_set(M, 0, 2, 6, 3, 9, S, 0), golden["mv_set"])
    assert same_bits(port.slice_set(M, 1, 2, 8, 3, 7, S, 1), golden["mv_setT"])
    assert same_bits(port.slice_set(M, 0, 2, 8, 3, 7, S, 1), golden["mv_set_mixed"])
    assert same_bits(port.stack(0, [(M, 0), (N1, 1), (N2, 0)]), golden["mv_hstack"])
    assert same_bits(port.stack(1, [(M, 0), (N1, 1), (golden["mv_N3"], 0)]), golden["mv_vstack"])
    # literal expectations of tests/testbasic.cu:67,73
    assert np.array_equal(port.stack(1, [(golden["t3_A"], 0), (golden["t3_B"], 0)]),
                          np.array([[1, 1, 3], [1, 1, 5], [-1, -1, -1], [9, 10, 11]], dtype=np.float32))
    assert np.array_equal(port.stack(0, [(golden["t3_A"], 0), (golden["t3_B"], 0)]),
                          np.array([[1, 1, 3, -1, -1, -1], [1, 1, 5, 9, 10, 11]], dtype=np.float32))


def test_port_softmax_head(port, golden):
    assert same_bits(port.softmax_cols(golden["sm_X"]), golden["sm_softmax"])
    assert same_bits(port.softmax_ce_grad(golden["sm_X"], golden["sm_Y"], 32), golden["sm_cegrad"])
    got = port.softmax_cols(golden["sm_X2"])
    ref2 = golden["sm_softmax2"]                                  # OpenBLAS column-sum order: 1e-5 relative
    assert np.all(np.abs(got - ref2) <= 1e-5 * np.abs(ref2) + 1e-30)
    assert np.allclose(got.sum(axis=0), 1.0, atol=1e-6)


def test_port_config1(port, golden):
    A, B = golden["c1_A"], golden["c1_B"]
    n = A.shape[0]
    prod = port.gemm(A, 0, B, 0)
    x = port.div_scalar(prod.ravel(order="F"), float(n))
    out = port.chain_softplus5(x).reshape(n, n, order="F")
    assert rel_fro(out, golden["c1_out"]) < 1e-6


def test_port_seeded_streams_bitexact(port, golden):
    assert same_bits(port.randn(0, 1001), golden["rng_randn_0"])
    assert same_bits(port.rand(7, 1001), golden["rng_rand_7"])


def test_port_norm(port, golden):
    assert port.norm(golden["ew_xs"]) == pytest.approx(float(golden["norm_x"]), rel=0, abs=0)


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (no /root/reference here)")
def test_live_reference_matches_fixtures(golden):
    R = oracle.ref()
    assert same_bits(R.testbasic_expr(golden["basic_A"], golden["basic_B"]).ravel(order="F"), BASIC_TESTDATA)
    assert same_bits(R.unary("exp", golden["ew_x"]), golden["ew_exp"])
    assert same_bits(R.chain_softplus5(golden["ew_xs"]), golden["ew_chain"])
    assert same_bits(R.gemm(golden["gemm_g1_A"], 0, golden["gemm_g1_B"], 0), golden["gemm_g1_blas_00"])
    assert same_bits(R.softmax_ce_grad(golden["sm_X"], golden["sm_Y"], 32), golden["sm_cegrad"])
    assert same_bits(R.randn(0, 1001), golden["rng_randn_0"])

"""GPU parity tests (-m gpu): libjz_b200.so through the C-ABI (ctypes) against
 (1) the fixtures the UNMODIFIED reference produced (tests/golden/ref_golden.npz),
 (2) the pinned CPU oracle (oracle/jz_oracle.c) on fresh seeded inputs,
 (3) size-independent properties at benchmark sizes.

Tolerances (BASELINE.json north_star): data movement / indexing bit-exact; elementwise <= 2 ulp;
reductions <= 1e-5 relative; GEMM is in tests/test_gemm_gpu.py.  The arithmetic maps
(affine, axpby, hadamard, A/B, eleminv, square, relu) are in fact held to BIT-EXACT because the
kernels use the same unfused fp32 operations as the x86-64 reference.
"""
import os

import numpy as np
import pytest

from conftest import bits, ulp_dist

pytestmark = pytest.mark.gpu


def F(a):
    return np.asfortranarray(a, dtype=np.float32)


def same_bits(a, b):
    return np.array_equal(bits(np.asfortranarray(a)), bits(np.asfortranarray(b)))


def flat(jz, x):
    return jz.CM(np.asarray(x, dtype=np.float32).reshape(-1, 1))


# ------------------------------------------------------------------ the reference's golden vector
def test_testbasic_golden_vector(jz, golden):
    """tests/testbasic.cu:28-55 (test2) through the Python mirror of the operators."""
    A, B = jz.CM(golden["basic_A"]), jz.CM(golden["basic_B"])
    C = jz.log(jz.exp(-A / B) + jz.exp(jz.hadmd(B, A))) - (A.T() * B).rows(0, 2)
    bench = jz.CM(golden["basic_expr"])
    assert (C - bench).norm() < 1e-5
    assert np.max(ulp_dist(C.to_host(), golden["basic_expr"])) <= 2


def test_shape_errors_raise_before_launch(jz):
    """tests/testbasic.cu:114-249: every incompatible pair is std::invalid_argument."""
    A, B, S = jz.CM.ones_(3, 4), jz.CM.ones_(3, 3), jz.CM.ones_(2, 2)
    before = jz.lib().jz_launch_count()
    for fn in (lambda: A + B, lambda: A - B, lambda: A * B, lambda: A / B, lambda: jz.hadmd(A, B),
               lambda: A.T() + B, lambda: A * B.T(), lambda: jz.hadmd(A.T(), B),
               lambda: jz.hstack([A, S]), lambda: jz.vstack([A, B]),
               lambda: jz.hstack([]), lambda: jz.vstack([])):
        with pytest.raises(ValueError):
            fn()
    assert jz.lib().jz_launch_count() == before


# ------------------------------------------------------------------ elementwise
@pytest.mark.parametrize("op,fn", [("exp", "exp"), ("tanh", "tanh"), ("dtanh", "d_tanh"), ("square", "square"),
                                   ("relu", "relu"), ("drelu", "d_relu")])
def test_unary_vs_reference_fixture(jz, golden, op, fn):
    x = golden["ew_x"]
    want = golden["ew_" + op]
    got = getattr(jz, fn)(flat(jz, x), inplace=False).to_host().ravel()
    d = ulp_dist(got, want)
    if op == "dtanh":
        # beyond |x| ~ 9 the reference's own `1 - tanh(x)^2` in double has cancelled to a multiple
        # of 2^-53 (its absolute error there is ~1e-16 >> 1 fp32 ulp of the true value); compare
        # within 2 ulp where the reference is itself accurate and within its quantum elsewhere.
        accurate = np.abs(x) <= 9.0
        assert np.max(d[accurate]) <= 2
        assert np.all(np.abs(got[~accurate].astype(np.float64) - want[~accurate]) <= 2.0 ** -51 + 2.4e-7 * np.abs(want[~accurate]))
    elif op in ("square", "relu", "drelu"):
        assert np.max(d) == 0
    else:
        assert np.max(d) <= 2, f"{op}: {np.max(d)} ulp at x={x[np.argmax(d)]}"


@pytest.mark.parametrize("op,fn", [("log", "log"), ("sqrt", "sqrt")])
def test_unary_positive_vs_reference_fixture(jz, golden, op, fn):
    got = getattr(jz, fn)(flat(jz, golden["ew_xp"])).to_host().ravel()
    d = ulp_dist(got, golden["ew_" + op])
    assert np.max(d) <= (0 if op == "sqrt" else 2)


def test_scalar_maps_bitexact_vs_reference_fixture(jz, golden):
    x, xp = flat(jz, golden["ew_x"]), flat(jz, golden["ew_xp"])
    assert same_bits(x.add(-0.3, 1.7).to_host().ravel(), golden["ew_affine"])
    assert same_bits((-x).to_host().ravel(), golden["ew_neg"])
    assert same_bits((x / 5.0).to_host().ravel(), golden["ew_div5"])
    assert same_bits((x / 4096.0).to_host().ravel(), golden["ew_div4096"])
    assert same_bits((1.0 / xp).to_host().ravel(), golden["ew_eleminv1"])
    assert same_bits((3.0 / xp).to_host().ravel(), golden["ew_eleminv3"])


def test_inplace_overloads_reuse_the_buffer(jz, golden):
    """rvalue overloads must update in place and hand back the same pointer
    (tests/testElementwiseReduceTorchDump.cu:45-48,69); lvalue overloads must not mutate."""
    x = golden["ew_xs"]
    m = flat(jz, x)
    p = m.ptr
    out = jz.exp(m, inplace=False)
    assert out.ptr != p and same_bits(m.to_host().ravel(), x)
    out = jz.exp(m, inplace=True)
    assert out.ptr == p


def test_chain_matches_stepwise_and_reference(jz, golden, port):
    """log(exp(x)+1)/5 : the fused one-pass chain is bit-identical to the four separate kernels
    and within 2 ulp of the reference's four CPU passes."""
    xs = golden["ew_xs"]
    m = flat(jz, xs)
    stepwise = jz.log(jz.exp(m) + 1.0) / 5.0
    fused = jz.chain(m, [("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", float(np.float32(1.0 / 5.0)), 0.0)])
    assert same_bits(stepwise.to_host(), fused.to_host())
    # end to end the chain is NOT a 2-ulp map: log(1+e) amplifies a 1-2 ulp difference in exp(x) when
    # e << 1.  Bound = 2 ulp per stage propagated through the chain: |d| <= 2^-22 * (|y| + 1/5).
    got, want = fused.to_host().ravel().astype(np.float64), golden["ew_chain"].astype(np.float64)
    assert np.all(np.abs(got - want) <= 2.0 ** -22 * (np.abs(want) + 0.2))
    # stage by stage (each map on the reference's own intermediate) every step is within 2 ulp
    e_ref = port.unary("exp", xs)
    assert np.max(ulp_dist(jz.exp(m).to_host().ravel(), e_ref)) <= 2
    s_ref = port.affine(e_ref, 1.0, 1.0)
    assert np.max(ulp_dist(jz.log(flat(jz, s_ref)).to_host().ravel(), port.unary("log", s_ref))) <= 2


@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_binary_all_flag_combinations_bitexact(jz, golden, ta, tb):
    A, B = golden["bin_A"], golden["bin_B"]
    a = jz.CM(F(A.T)).T() if ta else jz.CM(A)
    b = jz.CM(F(B.T)).T() if tb else jz.CM(B)
    assert same_bits(a.add(b, 1.5, -2.0).to_host(), golden[f"bin_axpby_{ta}{tb}"])
    assert same_bits(jz.hadmd(a, b).to_host(), golden[f"bin_hadmd_{ta}{tb}"])
    assert same_bits((a / b).to_host(), golden[f"bin_div_{ta}{tb}"])
    # in-place form keeps `this` layout (cpp/cumatrix.cu:246-260)
    c = a.copy()
    c.add(b, 1.5, -2.0, inplace=True)
    assert c.get_transpose() == ta and same_bits(c.to_host(), golden[f"bin_axpby_{ta}{tb}"])


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 1024, 1025, 4099, (1 << 20) + 7])
def test_ragged_sizes_and_unaligned_views(jz, port, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n + 3) * 2).astype(np.float32)
    d = flat(jz, x)
    L = jz.lib()
    for off in (0, 1, 3):  # 4- and 12-byte misaligned views exercise the scalar kernels
        out = jz.CM.empty("o", n + 3, 1)
        jz.fill(out, 7.0)
        jz._lib.check(L.jz_affine(out.ptr + 4 * off, d.ptr + 4 * off, n, 1.25, -0.5, None))
        got = out.to_host().ravel()
        assert same_bits(got[off:off + n], port.affine(x[off:off + n], 1.25, -0.5))
        assert np.all(got[:off] == 7.0) and np.all(got[off + n:] == 7.0)   # no out-of-range writes
        jz._lib.check(L.jz_unary(0, out.ptr + 4 * off, d.ptr + 4 * off, n, None))
        if n:
            assert np.max(ulp_dist(out.to_host().ravel()[off:off + n], port.unary("exp", x[off:off + n]))) <= 2


def test_packed_fp32_tile_paths_have_the_bits_of_the_scalar_paths(jz):
    """The vector kernels evaluate log, d_tanh and the `x + a` steps of a chain two elements per instruction (sm_100's
    FFMA2 / FADD2 / FMUL2, jz_math.cuh: log_main2, dtanh_acc2); a 4-byte-misaligned view of the same data takes the scalar
    kernels (map1_s, chain_s).  Every lane of a packed instruction rounds like the scalar instruction, so the two paths
    must agree BIT FOR BIT -- over every binade, the special values, and both sides of the d_tanh range switch."""
    rng = np.random.default_rng(2024)
    n = 1 << 20
    mags = np.exp2(rng.uniform(-126, 127, n)).astype(np.float32)
    x = (mags * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    x[:16] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 40.0, -40.0, 40.000004, 39.999996, 1e-45, -1e-45, 88.0, -88.0, 3.4e38]
    small = (rng.standard_normal(n) * 4).astype(np.float32)
    L = jz.lib()
    for name, data in (("wide", x), ("activations", small), ("positive", np.abs(x))):
        src = np.concatenate([np.zeros(1, np.float32), data])     # element 1.. = the same values at a 4-byte-misaligned address
        d_al, d_mis = flat(jz, data), flat(jz, src)
        for op in ("log", "dtanh", "exp", "tanh"):
            o_al, o_mis = jz.CM.empty("a", n, 1), jz.CM.empty("m", n + 1, 1)
            jz._lib.check(L.jz_unary(jz._lib.UNARY[op], o_al.ptr, d_al.ptr, n, None))
            jz._lib.check(L.jz_unary(jz._lib.UNARY[op], o_mis.ptr + 4, d_mis.ptr + 4, n, None))
            a, m = o_al.to_host().ravel(), o_mis.to_host().ravel()[1:]
            both_nan = np.isnan(a) & np.isnan(m)
            assert np.array_equal(bits(a)[~both_nan], bits(m)[~both_nan]), (name, op, int(np.sum(bits(a) != bits(m))))
        steps = [("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", float(np.float32(0.2)), 0.0), ("dtanh",), ("affine", 1.0, -0.25)]
        arr, ns = jz._lib.make_steps(steps)
        o_al, o_mis = jz.CM.empty("a", n, 1), jz.CM.empty("m", n + 1, 1)
        jz._lib.check(L.jz_chain(o_al.ptr, d_al.ptr, n, arr, ns, None))
        jz._lib.check(L.jz_chain(o_mis.ptr + 4, d_mis.ptr + 4, n, arr, ns, None))
        a, m = o_al.to_host().ravel(), o_mis.to_host().ravel()[1:]
        both_nan = np.isnan(a) & np.isnan(m)
        assert np.array_equal(bits(a)[~both_nan], bits(m)[~both_nan]), (name, "chain")


# ------------------------------------------------------------------ reductions
@pytest.mark.parametrize("name", ["r1", "r2", "r3", "r4"])
def test_reductions_vs_reference_fixture(jz, golden, name):
    M = golden[f"red_{name}"]
    for ta in (0, 1):
        m = jz.CM(M).T() if ta else jz.CM(M)
        for dim in (0, 1):
            blas = golden[f"red_{name}_sum_blas_t{ta}d{dim}"]
            got = jz.sum(m, dim).to_host().ravel()
            logical = M.T if ta else M
            scale = np.abs(logical).sum(axis=dim)
            assert np.all(np.abs(got - blas) <= 1e-5 * np.maximum(scale, 1e-30))
            mx = jz.colmax(m, dim).to_host().ravel()
            assert same_bits(mx, golden[f"red_{name}_max_t{ta}d{dim}"].ravel())   # max is order-free: bit-exact
            s = jz.sum(m, dim)
            assert (s.num_row(), s.num_col()) == ((1, logical.shape[1]) if dim == 0 else (logical.shape[0], 1))


@pytest.mark.parametrize("shape", [(1, 1), (1, 777), (777, 1), (31, 33), (32, 4096), (33, 4096), (4096, 33),
                                   (5, 100003), (100003, 5), (2048, 2048), (1001, 1001), (16, 1 << 18), (1 << 18, 16),
                                   (70000, 3), (40000, 40), (4096, 4096), (8200, 1000), (260, 30000)])
def test_sum_max_all_kernel_variants(jz, port, shape):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    M = F(rng.standard_normal(shape))
    m = jz.CM(M)
    for dim in (0, 1):
        truth = port.sum(M, 0, dim, f64=True)
        scale = np.abs(M).sum(axis=dim)
        got = jz.sum(m, dim).to_host().ravel()
        assert np.all(np.abs(got - truth) <= 1e-5 * np.maximum(scale, 1e-30)), (shape, dim)
        assert same_bits(jz.colmax(m, dim).to_host().ravel(), M.max(axis=dim))
        # no float atomics anywhere (cluster + ticket fold included): the same call gives the same bits
        assert same_bits(jz.sum(m, dim).to_host().ravel(), got)


def test_empty_reductions(jz):
    m = jz.CM.empty("e", 0, 5)
    assert np.array_equal(jz.sum(m, 0).to_host().ravel(), np.zeros(5, dtype=np.float32))
    assert jz.sum(m, 1).to_host().size == 0


def test_norm(jz, golden):
    x = golden["ew_xs"]
    got = flat(jz, x).norm()
    assert abs(got - np.sqrt(np.sum(x.astype(np.float64) ** 2))) <= 1e-6 * got
    assert abs(got - float(golden["norm_x"])) <= 1e-5 * got     # the reference's serial fp32 norm


def test_softmax_head_vs_reference_fixture(jz, golden):
    X, Y = golden["sm_X"], golden["sm_Y"]
    got = jz.softmax_cols(jz.CM(X)).to_host()
    ref = golden["sm_softmax"]
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-30)
    g = jz.softmax_ce_grad(jz.CM(X), jz.CM(Y), 32).to_host()
    refg = golden["sm_cegrad"]
    assert np.all(np.abs(g - refg) <= 1e-5 * np.abs(refg) + 1e-9)
    got2 = jz.softmax_cols(jz.CM(golden["sm_X2"])).to_host()       # long-column (warp) variant
    ref2 = golden["sm_softmax2"]
    assert np.all(np.abs(got2 - ref2) <= 1e-5 * np.abs(ref2) + 1e-30)
    assert np.allclose(got2.sum(axis=0), 1.0, atol=1e-5)


@pytest.mark.parametrize("rows", [16388, 20000, 32768, 40000, 70000, 131076, 200000, 262144, 300000])
def test_softmax_long_columns_cluster(jz, port, rows):
    """columns shared by a thread-block cluster (2 / 4 / 8 CTAs exchanging max and sum through distributed shared
    memory) up to 262144 rows; beyond that the three-pass fallback; both modes (plain softmax, CE gradient)"""
    rng = np.random.default_rng(rows)
    cols = 37
    X = F(rng.standard_normal((rows, cols)) * 3)
    # truth in float64: the oracle's sequential fp32 sum over > 16384 terms is itself ~1e-5 off
    X64 = X.astype(np.float64)
    E = np.exp(X64 - X64.max(axis=0, keepdims=True))
    ref = E / E.sum(axis=0, keepdims=True)
    got = jz.softmax_cols(jz.CM(X)).to_host()
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-30)
    assert np.all(np.abs(port.softmax_cols(X) - ref) <= 5e-3 * np.abs(ref) + 1e-30)      # the oracle (sequential fp32 sum, as the reference) agrees to ITS accuracy
    assert np.allclose(got.sum(axis=0, dtype=np.float64), 1.0, atol=1e-5)
    assert same_bits(jz.softmax_cols(jz.CM(X)).to_host(), got)            # deterministic
    if rows in (20000, 70000):
        Y = F((rng.random((rows, cols)) < 1e-4).astype(np.float32))
        g = jz.softmax_ce_grad(jz.CM(X), jz.CM(Y), 32).to_host()
        refg = -(Y.astype(np.float64) - ref) / 32.0
        assert np.all(np.abs(g - refg) <= 1e-5 * np.abs(refg) + 1e-9)


def test_softmax_very_long_columns_chunked(jz):
    """rows > 32768: a column is split into register-resident chunks, one CTA each, that meet through a ticket in global
    memory (jz_reduce.cu: softmax_chunks_kernel); many more column-chunks than resident CTAs, a ragged last chunk"""
    rng = np.random.default_rng(99)
    rows, cols = 40004, 600
    X = F(rng.standard_normal((rows, cols)) * 2)
    X64 = X.astype(np.float64)
    E = np.exp(X64 - X64.max(axis=0, keepdims=True))
    ref = E / E.sum(axis=0, keepdims=True)
    got = jz.softmax_cols(jz.CM(X)).to_host()
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-30)
    assert same_bits(jz.softmax_cols(jz.CM(X)).to_host(), got)


def test_softmax_composite_through_operators(jz, golden):
    """the reference's own formulation (ml/layer.hpp:254-262) through the mirrored operators,
    including the rank-1 GEMM broadcasts, equals the fused kernel."""
    X = jz.CM(golden["sm_X"])
    K = X.num_row()
    one = jz.CM.ones_(K, 1)
    mx = jz.colmax(X, 0)
    shifted = X - one * mx
    E = jz.exp(shifted, inplace=True)
    Z = one * jz.sum(E, 0)
    S = E / Z
    ref = golden["sm_softmax"]
    got = S.to_host()
    assert np.all(np.abs(got - ref) <= 1e-5 * np.abs(ref) + 1e-30)


# ------------------------------------------------------------------ data movement: bit-exact
def test_data_movement_vs_reference_fixture(jz, golden):
    M, S, N1, N2, N3 = (golden[k] for k in ("mv_M", "mv_S", "mv_N1", "mv_N2", "mv_N3"))
    m = jz.CM(M)
    assert same_bits(m.T().to_host(), golden["mv_T"])
    assert same_bits(m.slice(3, 20, 5, 30).to_host(), golden["mv_slice"])
    assert same_bits(m.T().slice(3, 20, 5, 19).to_host(), golden["mv_sliceT"])
    d = jz.CM(M); d.slice(2, 6, 3, 9, jz.CM(S))
    assert same_bits(d.to_host_physical(), golden["mv_set"])
    d = jz.CM(M).T(); d.slice(2, 8, 3, 7, jz.CM(S).T())
    assert same_bits(d.to_host_physical(), golden["mv_setT"])
    d = jz.CM(M); d.slice(2, 8, 3, 7, jz.CM(S).T())
    assert same_bits(d.to_host_physical(), golden["mv_set_mixed"])
    assert same_bits(jz.hstack([m, jz.CM(N1).T(), jz.CM(N2)]).to_host(), golden["mv_hstack"])
    v = jz.vstack([m, jz.CM(N1).T(), jz.CM(N3)])
    assert v.get_transpose() == 0 and same_bits(v.to_host(), golden["mv_vstack"])
    # tests/testbasic.cu:84-112 (test4)
    A = jz.CM(F([[1, 2, 3], [3, 4, 5]])); A.columns(0, 2, jz.CM(F([[1, 1], [1, 1]])))
    B = jz.CM(F([[6, 7, 8], [9, 10, 11]])); B.rows(0, 1, jz.CM(F([[-1, -1, -1]])))
    assert np.array_equal(jz.vstack([A, B]).to_host(), golden["t3_vstack"])
    assert np.array_equal(jz.hstack([A, B]).to_host(), golden["t3_hstack"])


@pytest.mark.parametrize("shape", [(1, 1), (1, 100), (100, 1), (31, 33), (32, 32), (64, 96), (1001, 1001),
                                   (4096, 33), (33, 4096), (2048, 1024)])
def test_transpose_roundtrip_bitexact(jz, shape):
    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    M = F(rng.standard_normal(shape))
    L = jz.lib()
    src = jz.CM(M)
    t = jz.CM.empty("t", shape[1], shape[0])
    jz._lib.check(L.jz_copy2d(t.ptr, shape[1], src.ptr, shape[0], shape[1], shape[0], 1, None))
    assert same_bits(t.to_host(), M.T)
    back = jz.CM.empty("b", shape[0], shape[1])
    jz._lib.check(L.jz_copy2d(back.ptr, shape[0], t.ptr, shape[1], shape[0], shape[1], 1, None))
    assert same_bits(back.to_host(), M)


def test_broadcast_idioms(jz, port):
    """b*ones(1,N) and ones(m,1)*v as first-class ops equal the rank-1 GEMM the reference runs"""
    rng = np.random.default_rng(5)
    A = F(rng.standard_normal((128, 50))); b = rng.standard_normal(128).astype(np.float32)
    v = rng.standard_normal(50).astype(np.float32)
    L = jz.lib()
    a_d, b_d, v_d = jz.CM(A), jz.CM(b.reshape(-1, 1)), jz.CM(v.reshape(-1, 1))
    out = jz.CM.empty("o", 128, 50)
    jz._lib.check(L.jz_add_bcast(out.ptr, a_d.ptr, 128, 50, b_d.ptr, 1, 1.0, 1.0, None))
    assert same_bits(out.to_host(), A + b[:, None])
    jz._lib.check(L.jz_add_bcast(out.ptr, a_d.ptr, 128, 50, v_d.ptr, 0, 1.0, -1.0, None))
    assert same_bits(out.to_host(), A - v[None, :])
    outer = b_d * v_d.T()                    # k == 1 GEMM -> rank-1 kernel
    assert L.jz_gemm_last_path() == 3
    assert same_bits(outer.to_host(), port.gemm(b.reshape(-1, 1), 0, v.reshape(1, -1), 0))


# ------------------------------------------------------------------ pool + rng + adam
@pytest.mark.parametrize("count", [1, 1000, (1 << 17) - 3, 1 << 17, (1 << 17) + 5, 1 << 20])
def test_staged_upload_is_bit_exact_and_source_is_free_on_return(jz, count):
    """jz_upload (pinned staging ring, 16 slots of 512 KiB; larger copies take the plain path): more uploads than slots,
    the pageable source scribbled over right after each call, bit-exact contents on the device"""
    L = jz.lib()
    rng = np.random.default_rng(count)
    keep, devs = [], []
    src = np.empty(count, dtype=np.float32)
    for i in range(40):
        data = rng.standard_normal(count).astype(np.float32)
        src[:] = data
        d = jz.CM.empty("u", count, 1)
        jz._lib.check(L.jz_upload(d.ptr, src.ctypes.data, count, None))
        src[:] = -1.0            # the caller's buffer is free again as soon as the call returns
        keep.append(data)
        devs.append(d)
    for data, d in zip(keep, devs):
        assert same_bits(d.to_host().ravel(), data)


def test_pool_reuses_exact_size_blocks(jz):
    import ctypes
    L = jz.lib()
    st = [ctypes.c_size_t() for _ in range(4)]
    a = jz.CM.empty("a", 1000, 1000); p = a.ptr; del a
    L.jz_pool_stats(*[ctypes.byref(s) for s in st]); allocs0, hits0 = st[2].value, st[3].value
    b = jz.CM.empty("b", 1000, 1000)
    L.jz_pool_stats(*[ctypes.byref(s) for s in st])
    assert b.ptr == p and st[2].value == allocs0 and st[3].value == hits0 + 1   # no cudaMalloc on the hot path
    z = jz.CM.empty("z", 0, 7)          # size 0 -> 1 element (cpp/cumatrix.cuh:71)
    assert z.ptr != 0


def test_rng_moments_and_determinism(jz):
    a = jz.CM.randn(1000, 1001, seed=1, offset=0).to_host()      # odd count: no scratch buffer needed
    b = jz.CM.randn(1000, 1001, seed=1, offset=0).to_host()
    c = jz.CM.randn(1000, 1001, seed=2, offset=0).to_host()
    d1, d2 = jz.CM.randn(64, 64).to_host(), jz.CM.randn(64, 64).to_host()   # default arguments: successive streams
    assert not np.array_equal(d1, d2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert abs(a.mean()) < 5e-3 and abs(a.std() - 1) < 5e-3 and np.isfinite(a).all()
    u = jz.CM.rand(1001, 999, seed=3).to_host()
    assert u.min() >= 0 and u.max() < 1 and abs(u.mean() - 0.5) < 2e-3


def adam_bc(beta, t):
    """bias correction as the reference's CPU path computes it: double pow, double reciprocal, float cast (ml/util.cuh:226-227)"""
    return np.float32(1.0 / (1.0 - np.float64(np.float32(beta)) ** t))


def test_adam_steps_bit_exact_vs_reference_cpu_adam_update(jz, port):
    """three consecutive jz_adam_update steps against golden vectors produced by the UNMODIFIED reference's
    adam_update<float> (ml/util.cuh:165-257, through oracle/ref_shim_ml.cpp; scripts/make_golden_ml.py): g, m and v
    bit for bit, and the same against the C oracle's restatement on a larger, unaligned, odd-length case"""
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_ml_golden.npz"))
    g_in = G["adam_g_in"]
    n = g_in.shape[1]
    L = jz.lib()
    md, vd = flat(jz, np.zeros(n, np.float32)), flat(jz, np.zeros(n, np.float32))
    for t in range(3):
        gd = flat(jz, g_in[t])
        jz._lib.check(L.jz_adam_update(gd.ptr, md.ptr, vd.ptr, n, 0.01, 0.9, 0.999, 1e-8, adam_bc(0.9, t + 1), adam_bc(0.999, t + 1), None))
        assert same_bits(gd.to_host().ravel(), G["adam_update"][t]), f"update, step {t + 1}"
        assert same_bits(md.to_host().ravel(), G[f"adam_m_{t + 1}"]) and same_bits(vd.to_host().ravel(), G[f"adam_v_{t + 1}"])
    rng = np.random.default_rng(11)
    n = 1_000_003
    g = rng.standard_normal(n + 1).astype(np.float32); m = (rng.standard_normal(n + 1) * 0.1).astype(np.float32)
    v = (np.abs(rng.standard_normal(n + 1)) * 0.01).astype(np.float32)
    for off in (0, 1):          # off = 1: pointers 4 bytes past a 16-byte boundary -> scalar kernel
        gd, md, vd = flat(jz, g), flat(jz, m), flat(jz, v)
        bc1, bc2 = adam_bc(0.9, 7), adam_bc(0.999, 7)
        jz._lib.check(L.jz_adam_update(gd.ptr + 4 * off, md.ptr + 4 * off, vd.ptr + 4 * off, n, 0.003, 0.9, 0.999, 1e-8, bc1, bc2, None))
        wg, wm, wv = port.adam_update(g[off:off + n], m[off:off + n], v[off:off + n], 0.003, 0.9, 0.999, 1e-8, bc1, bc2)
        assert same_bits(gd.to_host().ravel()[off:off + n], wg) and same_bits(md.to_host().ravel()[off:off + n], wm)
        assert same_bits(vd.to_host().ravel()[off:off + n], wv)
        if off:   # the element in front of the window did not move
            assert gd.to_host().ravel()[0] == g[0] and md.to_host().ravel()[0] == m[0]


# ------------------------------------------------------------------ exhaustive accuracy sweeps (every fp32 input)
@pytest.mark.parametrize("op,limit", [("exp", 2), ("log", 2), ("tanh", 2), ("dtanh", 2), ("sqrt", 0), ("square", 0)])
def test_exhaustive_ulp_sweep(jz, op, limit):
    """all 2^32 bit patterns against an fp64 evaluation rounded once (what the g++ reference
    computes for exp/log/tanh; for d_tanh the cancellation-free sech^2)."""
    import ctypes
    L = jz.lib()
    mx, worst = ctypes.c_uint32(), ctypes.c_uint32()
    ranges = [(0x00000000, 0x7F800000), (0x80000000, 0xFF800000)]
    if op in ("log", "sqrt"):
        ranges = ranges[:1]
    tot = 0
    for lo, hi in ranges:
        jz._lib.check(L.jz_unary_ulp_sweep(jz._lib.UNARY[op], lo, hi, ctypes.byref(mx), ctypes.byref(worst), None))
        tot = max(tot, mx.value)
        x = np.array([worst.value], dtype=np.uint32).view(np.float32)[0]
        print(f"ulp sweep {op} [{lo:#x},{hi:#x}): max {mx.value} ulp at x={x!r}")
    assert tot <= limit


# ------------------------------------------------------------------ properties at benchmark sizes
@pytest.mark.parametrize("log2n", [24, 28])
def test_large_flat_properties(jz, log2n):
    """sizes the CPU oracle cannot cover in seconds: linearity / idempotence / checksums."""
    n = 1 << log2n
    x = jz.CM.randn(n, 1, seed=log2n)
    s1 = float(jz.sum(x, 0).to_host()[0, 0])
    y = x.add(0.5, 2.0)                                # 2x + 0.5
    s2 = float(jz.sum(y, 0).to_host()[0, 0])
    assert abs(s2 - (2 * s1 + 0.5 * n)) <= 1e-5 * (abs(s1) * 2 + 0.5 * n)
    r = jz.relu(y.copy(), inplace=True)
    r2 = jz.relu(r.copy(), inplace=True)
    assert float((r - r2).norm()) == 0.0               # idempotent
    e = jz.log(jz.exp(x))                              # log(exp(x)) ~ x within a few ulp of |x|
    assert float((e - x).norm()) <= 4e-7 * float(x.norm()) + 1e-3
    nx = x.norm()
    assert abs(nx - np.sqrt(n)) < 0.01 * np.sqrt(n)
    # transpose twice over a 2-D view is the identity, bit for bit
    rows = 1 << (log2n // 2)
    cols = n // rows
    t = jz.CM.empty("t", cols, rows)
    L = jz.lib()
    jz._lib.check(L.jz_copy2d(t.ptr, cols, x.ptr, rows, cols, rows, 1, None))
    b = jz.CM.empty("b", rows, cols)
    jz._lib.check(L.jz_copy2d(b.ptr, rows, t.ptr, cols, rows, cols, 1, None))
    xx = jz.CM(_raw=("v", n, 1, False, b.buf))
    assert float((xx - x).norm()) == 0.0
    # row sums and column sums agree on the grand total
    v = jz.CM(_raw=("v", rows, cols, False, x.buf))
    tot0 = float(jz.sum(jz.sum(v, 0), 1).to_host()[0, 0])
    tot1 = float(jz.sum(jz.sum(v, 1), 0).to_host()[0, 0])
    assert abs(tot0 - tot1) <= 1e-5 * np.sqrt(n) * 4 and abs(tot0 - s1) <= 1e-5 * np.sqrt(n) * 4


def test_more_than_2_to_32_elements(jz):
    """The reference casts element counts to unsigned int (cpp/cumatrix.cuh:59-62) and window indices to int
    (cpp/cukernels.cu:79-83); here every index is size_t.  65536 x 65552 = 2^32 + 2^20 elements (16 GiB): fill, in-place
    affine map, column and row sums with exactly representable results, a marker written past element 2^32, and a
    transposing copy of the far corner."""
    import torch
    if torch.cuda.mem_get_info()[0] < 40 * (1 << 30):
        pytest.skip("needs ~17 GiB of free device memory")
    L = jz.lib()
    rows, cols = 65536, 65552
    n = rows * cols
    assert n > (1 << 32)
    x = jz.CM.empty("big", rows, cols)
    jz._lib.check(L.jz_fill(x.ptr, n, 1.5, None))
    jz._lib.check(L.jz_affine(x.ptr, x.ptr, n, 2.0, 1.0, None))                     # 4.0 everywhere
    jz._lib.check(L.jz_fill(x.ptr + 4 * (n - 3), 3, 7.0, None))                       # marker in the last column
    c = jz.CM.empty("c", cols, 1)
    r = jz.CM.empty("r", rows, 1)
    jz._lib.check(L.jz_sum(c.ptr, x.ptr, rows, cols, rows, 0, None))
    jz._lib.check(L.jz_sum(r.ptr, x.ptr, rows, cols, rows, 1, None))
    cs, rs = c.to_host().ravel(), r.to_host().ravel()
    assert np.all(cs[:-1] == 4.0 * rows) and cs[-1] == 4.0 * rows + 3 * 3.0
    assert np.all(rs[:-3] == 4.0 * cols) and np.all(rs[-3:] == 4.0 * cols + 3.0)
    # far corner through the 2-D window path: last 4 rows x last 2 columns, transposed
    w = jz.CM.empty("w", 2, 4)
    off = (cols - 2) * rows + (rows - 4)
    jz._lib.check(L.jz_copy2d(w.ptr, 2, x.ptr + 4 * off, rows, 2, 4, 1, None))
    assert np.array_equal(w.to_host(), np.array([[4, 4, 4, 4], [4, 7, 7, 7]], dtype=np.float32))

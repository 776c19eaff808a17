"""GPU parity tests of the C++ SHELL (-m gpu): ``Matrix<CUDAfloat>`` as juzhen_b200/cpp/cumatrix.cuh + the reference's own
operators.hpp / juzhen.hpp / ml/layer.hpp define it -- the class a C++ user of the reference compiles against -- driven
through the flat C wrapper juzhen_b200/cpp/tests/shell_capi.cu (build/dropin/lib/libjz_shell_capi.so) against
 (1) the fixtures the UNMODIFIED reference produced (tests/golden/ref_golden.npz), with the tolerances of
     tests/test_parity_gpu.py (data movement and arithmetic maps bit-exact, exp/log/tanh <= 2 ulp, sums 1e-5, GEMM 1e-5);
 (2) the eager C-ABI mirror (juzhen_b200/matrix.py) on random operator programs, BIT FOR BIT: two independent
     implementations of the operator dispatch (C++ with deferred fills / products / elementwise chains / fused heads, and
     Python issuing one jz_* call per operator) must produce the same bits.
The wrapper library is built in the build container (it needs /root/reference's headers) and travels to the GPU box."""
import numpy as np
import pytest

from conftest import bits, rel_fro, ulp_dist

import shell_backend as sh

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not sh.available(), reason="libjz_shell_capi.so not built (build_dropin.py)")]


def F(a):
    return np.asfortranarray(a, dtype=np.float32)


def same_bits(a, b):
    return np.array_equal(bits(np.asfortranarray(a)), bits(np.asfortranarray(b)))


def flat(x):
    return sh.SM(np.asarray(x, dtype=np.float32).reshape(-1, 1))


def test_shell_testbasic_golden_vector(golden):
    """tests/testbasic.cu:28-55 written in C++ exactly as the reference writes it (shell_expr_testbasic)"""
    C = sh.expr_testbasic(sh.SM(golden["basic_A"]), sh.SM(golden["basic_B"]))
    assert np.max(ulp_dist(C.to_host(), golden["basic_expr"])) <= 2
    assert (C - sh.SM(golden["basic_expr"])).norm() < 1e-5


def test_shell_shape_errors_are_invalid_argument():
    """tests/testbasic.cu:114-249: every incompatible pair throws std::invalid_argument"""
    A, B, S = sh.SM.ones_(3, 4), sh.SM.ones_(3, 3), sh.SM.ones_(2, 2)
    for fn in (lambda: A + B, lambda: A - B, lambda: A * B, lambda: A / B, lambda: sh.hadmd(A, B),
               lambda: A.T() + B, lambda: A * B.T(), lambda: sh.hadmd(A.T(), B),
               lambda: sh.hstack([A, S]), lambda: sh.vstack([A, B]), lambda: sh.hstack([]), lambda: sh.vstack([])):
        with pytest.raises(sh.ShellShapeError):
            fn()
    assert same_bits((A + A).to_host(), np.full((3, 4), 2.0))   # the shell is still usable afterwards


@pytest.mark.parametrize("op,fn", [("exp", "exp"), ("tanh", "tanh"), ("dtanh", "d_tanh"), ("square", "square"),
                                   ("relu", "relu"), ("drelu", "d_relu")])
@pytest.mark.parametrize("move", [False, True])
def test_shell_unary_vs_reference_fixture(golden, op, fn, move):
    if move and fn in ("relu", "d_relu"):
        pytest.skip("functor maps are exercised through the const& overload")
    x = golden["ew_x"]
    want = golden["ew_" + op]
    m = flat(x)
    p = m.data()
    out = getattr(sh, fn)(m, move) if fn not in ("relu", "d_relu") else getattr(sh, fn)(m)
    got = out.to_host().ravel()
    if move:   # the && overload works in place and hands the same buffer back (tests/testElementwiseReduceTorchDump.cu:45-48)
        assert out.data() == p
    else:
        assert out.data() != p and same_bits(m.to_host().ravel(), x)
    d = ulp_dist(got, want)
    if op == "dtanh":
        accurate = np.abs(x) <= 9.0
        assert np.max(d[accurate]) <= 2
        assert np.all(np.abs(got[~accurate].astype(np.float64) - want[~accurate]) <= 2.0 ** -51 + 2.4e-7 * np.abs(want[~accurate]))
    elif op in ("square", "relu", "drelu"):
        assert np.max(d) == 0
    else:
        assert np.max(d) <= 2, f"{op}: {np.max(d)} ulp"


def test_shell_log_sqrt_and_scalar_maps_vs_reference_fixture(golden):
    xp, x = flat(golden["ew_xp"]), flat(golden["ew_x"])
    assert np.max(ulp_dist(sh.log(xp).to_host().ravel(), golden["ew_log"])) <= 2
    assert np.max(ulp_dist(sh.sqrt(xp).to_host().ravel(), golden["ew_sqrt"])) == 0
    assert same_bits((x * 1.7 - 0.3).to_host().ravel(), golden["ew_affine"]) or \
        np.max(ulp_dist((x * 1.7 - 0.3).to_host().ravel(), golden["ew_affine"])) <= 1   # two roundings here, one in add(a, s1)
    assert same_bits((-x).to_host().ravel(), golden["ew_neg"])
    assert same_bits((x / 5.0).to_host().ravel(), golden["ew_div5"])
    assert same_bits((x / 4096.0).to_host().ravel(), golden["ew_div4096"])
    assert same_bits((1.0 / xp).to_host().ravel(), golden["ew_eleminv1"])
    assert same_bits((3.0 / xp).to_host().ravel(), golden["ew_eleminv3"])
    # the README chain as ONE C++ expression of rvalues: the shell defers it into a single pass
    xs = golden["ew_xs"]
    got = (sh.log(sh.exp(flat(xs)) + 1.0) / 5.0).to_host().ravel().astype(np.float64)
    want = golden["ew_chain"].astype(np.float64)
    assert np.all(np.abs(got - want) <= 2.0 ** -22 * (np.abs(want) + 0.2))


@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_shell_binary_all_flag_combinations_bitexact(golden, ta, tb):
    A, B = golden["bin_A"], golden["bin_B"]
    a = sh.SM(F(A.T)).T() if ta else sh.SM(A)
    b = sh.SM(F(B.T)).T() if tb else sh.SM(B)
    assert same_bits((a * 1.5 - b * 2.0).to_host(), golden[f"bin_axpby_{ta}{tb}"]) or \
        np.max(ulp_dist((a * 1.5 - b * 2.0).to_host(), golden[f"bin_axpby_{ta}{tb}"])) <= 1
    assert same_bits(sh.hadmd(a, b).to_host(), golden[f"bin_hadmd_{ta}{tb}"])
    assert same_bits((a / b).to_host(), golden[f"bin_div_{ta}{tb}"])
    # rvalue overloads: same values, result in the moved operand's buffer when the flags allow it
    for move_a, move_b in ((True, False), (False, True)):
        a2, b2 = a.copy(), b.copy()
        assert same_bits(sh.hadmd(a2, b2, move_a, move_b).to_host(), golden[f"bin_hadmd_{ta}{tb}"])
        a2, b2 = a.copy(), b.copy()
        assert same_bits(sh.binary("div", a2, b2, move_a, move_b).to_host(), golden[f"bin_div_{ta}{tb}"])
        a2, b2 = a.copy(), b.copy()
        got = sh.binary("sub", a2, b2, move_a, move_b).to_host()
        assert same_bits(got, (A - B).astype(np.float32))
    c = a.copy()
    c += b
    assert c.get_transpose() == bool(ta) and same_bits(c.to_host(), (A + B).astype(np.float32))
    c -= b
    c -= b
    assert same_bits(c.to_host(), ((A + B).astype(np.float32) - B - B).astype(np.float32))


@pytest.mark.parametrize("name", ["r1", "r2", "r3", "r4"])
def test_shell_reductions_vs_reference_fixture(golden, name):
    M = golden[f"red_{name}"]
    for ta in (0, 1):
        m = sh.SM(M).T() if ta else sh.SM(M)
        logical = M.T if ta else M
        for dim in (0, 1):
            blas = golden[f"red_{name}_sum_blas_t{ta}d{dim}"]
            s = sh.sum(m, dim)
            got = s.to_host().ravel()
            scale = np.abs(logical).sum(axis=dim)
            assert np.all(np.abs(got - blas) <= 1e-5 * np.maximum(scale, 1e-30))
            assert (s.num_row(), s.num_col()) == ((1, logical.shape[1]) if dim == 0 else (logical.shape[0], 1))
        mx = sh.colmax(m)   # reduce<F> with the reference's max functor, dim 0
        assert same_bits(mx.to_host().ravel(), golden[f"red_{name}_max_t{ta}d0"].ravel())
    x = golden["ew_xs"]
    assert abs(flat(x).norm() - float(golden["norm_x"])) <= 1e-5 * float(golden["norm_x"])


def test_shell_data_movement_vs_reference_fixture(golden):
    M, S, N1, N2, N3 = (golden[k] for k in ("mv_M", "mv_S", "mv_N1", "mv_N2", "mv_N3"))
    m = sh.SM(M)
    assert same_bits(m.T().to_host(), golden["mv_T"])
    assert same_bits(m.slice(3, 20, 5, 30).to_host(), golden["mv_slice"])
    assert same_bits(m.T().slice(3, 20, 5, 19).to_host(), golden["mv_sliceT"])
    d = sh.SM(M); d.slice(2, 6, 3, 9, sh.SM(S))
    assert same_bits(d.to_host(), golden["mv_set"])
    d = sh.SM(M).T(); d.slice(2, 8, 3, 7, sh.SM(S).T())
    assert same_bits(d.to_host().T, golden["mv_setT"])          # fixture = the physical buffer of the flagged matrix
    d = sh.SM(M); d.slice(2, 8, 3, 7, sh.SM(S).T())
    assert same_bits(d.to_host(), golden["mv_set_mixed"])
    assert same_bits(sh.hstack([m, sh.SM(N1).T(), sh.SM(N2)]).to_host(), golden["mv_hstack"])
    assert same_bits(sh.vstack([m, sh.SM(N1).T(), sh.SM(N3)]).to_host(), golden["mv_vstack"])
    A = sh.SM(F([[1, 2, 3], [3, 4, 5]])); A.columns(0, 2, sh.SM(F([[1, 1], [1, 1]])))
    B = sh.SM(F([[6, 7, 8], [9, 10, 11]])); B.rows(0, 1, sh.SM(F([[-1, -1, -1]])))
    assert np.array_equal(sh.vstack([A, B]).to_host(), golden["t3_vstack"])
    assert np.array_equal(sh.hstack([A, B]).to_host(), golden["t3_hstack"])
    z = sh.SM.named(5, 7)          # Matrix(name, r, c) is observably zero-filled (cpp/cumatrix.cu:50-63)
    assert np.array_equal(z.to_host(), np.zeros((5, 7), dtype=np.float32))


@pytest.mark.parametrize("g", ["g1", "g2", "g3", "g4"])
def test_shell_dot_vs_reference_fixture(golden, g):
    A, B = golden[f"gemm_{g}_A"], golden[f"gemm_{g}_B"]
    for ta in (0, 1):
        for tb in (0, 1):
            a = sh.SM(F(A.T)).T() if ta else sh.SM(A)
            b = sh.SM(F(B.T)).T() if tb else sh.SM(B)
            got = (a * b).to_host()
            assert rel_fro(got, golden[f"gemm_{g}_fixed_{ta}{tb}"]) < 1e-5
            assert rel_fro(got, golden[f"gemm_{g}_blas_{ta}{tb}"]) < 1e-5


def test_shell_config1_expression_and_softmax_head(golden):
    """config 1 in miniature as one C++ expression, and the reference's own LogisticLayer<CUDAfloat> (ml/layer.hpp,
    unchanged): grad() is recognised by the lazy layer and runs as one fused kernel, eval() runs operator by operator"""
    out = sh.expr_softplus5(sh.SM(golden["c1_A"]), sh.SM(golden["c1_B"]), 192.0).to_host()
    assert rel_fro(out, golden["c1_out"]) < 1e-5
    X, Y = golden["sm_X"], golden["sm_Y"]
    g = sh.logistic_grad(sh.SM(X), sh.SM(Y)).to_host()
    refg = golden["sm_cegrad"] * (32.0 / X.shape[1])     # fixture was generated with nb = 32; the layer divides by its batch
    assert np.all(np.abs(g - refg) <= 1e-5 * np.abs(refg) + 1e-9)
    loss = sh.logistic_loss(sh.SM(X), sh.SM(Y)).to_host()
    xs = X.astype(np.float64)
    lse = np.log(np.exp(xs - xs.max(axis=0)).sum(axis=0)) + xs.max(axis=0)
    want = -((xs * Y).sum(axis=0) - lse).sum() / X.shape[1]
    assert loss.shape == (1, 1) and abs(float(loss[0, 0]) - want) <= 1e-5 * abs(want)


# ------------------------------------------------------------------ shell vs eager mirror, bit for bit
def _random_program(rng, steps):
    """a list of operator applications over a growing pool of matrices; both backends execute the same list"""
    prog = []
    shapes = [(37, 53), (37, 53), (53, 37), (53, 53)]   # pool entries 0..3 are the inputs
    while len(prog) < steps:
        kind = rng.choice(["unary", "scalar", "binary", "dot", "T", "sum", "slice", "stack", "iadd", "move_unary", "move_binary"])
        i = int(rng.integers(0, len(shapes)))
        r, c = shapes[i]
        if kind in ("unary", "move_unary"):
            op = str(rng.choice(["tanh", "square", "d_tanh", "neg", "exp_small"]))
            prog.append((kind, op, i)); shapes.append((r, c))
        elif kind == "scalar":
            op = str(rng.choice(["add", "sub", "rsub", "mul", "div", "rdiv_safe"]))
            prog.append((kind, op, i, float(np.float32(rng.uniform(0.5, 3.0))))); shapes.append((r, c))
        elif kind in ("binary", "iadd", "move_binary"):
            js = [j for j, s in enumerate(shapes) if s == (r, c)]
            j = int(rng.choice(js))
            op = str(rng.choice(["add", "sub", "hadmd"]))
            prog.append((kind, op, i, j)); shapes.append((r, c))
        elif kind == "dot":
            js = [j for j, s in enumerate(shapes) if s[0] == c and s[1] <= 64]
            if not js or r > 64:
                continue
            j = int(rng.choice(js))
            prog.append((kind, i, j)); shapes.append((r, shapes[j][1]))
        elif kind == "T":
            prog.append((kind, i)); shapes.append((c, r))
        elif kind == "sum":
            dim = int(rng.integers(0, 2))
            prog.append((kind, i, dim)); shapes.append((1, c) if dim == 0 else (r, 1))
        elif kind == "slice":
            r0, c0 = int(rng.integers(0, r)), int(rng.integers(0, c))
            r1, c1 = int(rng.integers(r0 + 1, r + 1)), int(rng.integers(c0 + 1, c + 1))
            prog.append((kind, i, r0, r1, c0, c1)); shapes.append((r1 - r0, c1 - c0))
        elif kind == "stack":
            js = [j for j, s in enumerate(shapes) if s[0] == r and c + s[1] <= 256]
            if not js:
                continue
            j = int(rng.choice(js))
            prog.append((kind, i, j)); shapes.append((r, c + shapes[j][1]))
    return prog


def _run(prog, inputs, be):
    """be: backend adaptor with the same method names for the shell and for the mirror"""
    pool = [be.make(x) for x in inputs]
    for st in prog:
        kind = st[0]
        if kind in ("unary", "move_unary"):
            _, op, i = st
            src = be.copy(pool[i]) if kind == "move_unary" else pool[i]
            mv = kind == "move_unary"
            if op == "exp_small":
                out = be.unary("exp", be.scalar("mul", be.unary("tanh", src, False), 0.5), True)   # bounded argument
            elif op == "square":
                out = be.unary("square", be.unary("tanh", src, mv), True)                          # bounded: programs are long
            else:
                out = be.unary(op, src, mv)
        elif kind == "scalar":
            _, op, i, s = st
            if op == "rdiv_safe":
                out = be.scalar("rdiv", be.scalar("add", be.unary("square", pool[i], False), 1.0), s)
            else:
                out = be.scalar(op, pool[i], s)
        elif kind == "binary":
            _, op, i, j = st
            out = be.binary(op, pool[i], pool[j], False, False)
            if op == "hadmd":
                out = be.scalar("mul", out, 0.25)
        elif kind == "move_binary":
            _, op, i, j = st
            out = be.binary(op, be.copy(pool[i]), pool[j], True, False)
            if op == "hadmd":
                out = be.scalar("mul", out, 0.25)
        elif kind == "iadd":
            _, op, i, j = st
            out = be.copy(pool[i])
            out = be.iadd(out, pool[j]) if op != "sub" else be.isub(out, pool[j])
        elif kind == "dot":
            out = be.scalar("mul", be.binary("mul", pool[st[1]], pool[st[2]], False, False), 1.0 / 64.0)   # keep magnitudes bounded
        elif kind == "T":
            out = be.T(pool[st[1]])
        elif kind == "sum":
            out = be.scalar("mul", be.sum(pool[st[1]], st[2]), 1.0 / 16.0)
        elif kind == "slice":
            out = be.slice(pool[st[1]], *st[2:])
        elif kind == "stack":
            out = be.hstack([pool[st[1]], pool[st[2]]])
        pool.append(out)
    return [be.host(m) for m in pool]


class _ShellBE:
    make = staticmethod(lambda x: sh.SM(x))
    copy = staticmethod(lambda m: m.copy())
    unary = staticmethod(lambda op, m, mv: sh.unary(op, m, mv))
    scalar = staticmethod(lambda op, m, s: sh.scalar(op, m, s))
    binary = staticmethod(lambda op, a, b, ma, mb: sh.binary(op, a, b, ma, mb))
    T = staticmethod(lambda m: m.T())
    sum = staticmethod(lambda m, d: sh.sum(m, d))
    slice = staticmethod(lambda m, r0, r1, c0, c1: m.slice(r0, r1, c0, c1))
    hstack = staticmethod(lambda ms: sh.hstack(ms))
    host = staticmethod(lambda m: m.to_host())

    @staticmethod
    def iadd(a, b):
        a += b
        return a

    @staticmethod
    def isub(a, b):
        a -= b
        return a


def _mirror_backend(jz):
    class BE:
        make = staticmethod(lambda x: jz.CM(x))
        copy = staticmethod(lambda m: m.copy())
        T = staticmethod(lambda m: m.T())
        sum = staticmethod(lambda m, d: jz.sum(m, d))
        slice = staticmethod(lambda m, r0, r1, c0, c1: m.slice(r0, r1, c0, c1))
        hstack = staticmethod(lambda ms: jz.hstack(ms))
        host = staticmethod(lambda m: m.to_host())

        @staticmethod
        def unary(op, m, mv):
            if op == "neg":
                return -m
            return getattr(jz, op)(m, inplace=mv)

        @staticmethod
        def scalar(op, m, s):
            return {"add": lambda: m + s, "sub": lambda: m - s, "rsub": lambda: s - m, "mul": lambda: m * s,
                    "div": lambda: m / s, "rdiv": lambda: s / m}[op]()

        @staticmethod
        def binary(op, a, b, ma, mb):
            if op == "mul":
                return a * b
            if op == "hadmd":
                return jz.hadmd(a, b, inplace_lhs=ma, inplace_rhs=mb)
            if ma:   # operators.hpp: an rvalue lhs is updated in place
                return a.add(b, 1.0, 1.0 if op == "add" else -1.0, inplace=True) or a
            return a + b if op == "add" else a - b

        @staticmethod
        def iadd(a, b):
            a += b
            return a

        @staticmethod
        def isub(a, b):
            a -= b
            return a
    return BE


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_shell_equals_eager_mirror_bitwise_on_random_programs(jz, seed):
    """150-step random operator programs (maps, scalar ops, axpby / hadamard with every flag combination, products, lazy
    transposes, sums, slices, stacks, compound assignment, rvalue overloads): the C++ shell -- deferring fills, products,
    elementwise chains and broadcast adds -- and the eager Python mirror agree BIT FOR BIT on every intermediate."""
    rng = np.random.default_rng(100 + seed)
    inputs = [F(rng.standard_normal(s)) for s in ((37, 53), (37, 53), (53, 37), (53, 53))]
    prog = _random_program(rng, 150)
    a = _run(prog, inputs, _ShellBE)
    b = _run(prog, inputs, _mirror_backend(jz))
    assert len(a) == len(b)
    for k, (x, y) in enumerate(zip(a, b)):
        assert x.shape == y.shape, (seed, k, prog[k - 4] if k >= 4 else "input")
        assert np.all(np.isfinite(y)), (seed, k, prog[k - 4] if k >= 4 else "input")
        assert same_bits(x, y), (seed, k, prog[k - 4] if k >= 4 else "input", float(np.max(np.abs(x - y))))

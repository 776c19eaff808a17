"""Generate tests/golden/ref_golden.npz from the UNMODIFIED reference.

Runs in the build container only (needs oracle/_ref/libjzref*.so, i.e. /root/reference).  Every
array saved here is either a seeded input or the output the reference's own Matrix<float> CPU
code (OpenBLAS build, and the -DJUZHEN_NO_BLAS build where summation order matters) produced
for it.  tests/test_oracle.py pins oracle/jz_oracle.c against these; the GPU parity tests
compare libjz_b200.so against them on the GPU box, where /root/reference does not exist.

    python scripts/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def F(a):
    return np.asfortranarray(a, dtype=np.float32)


def main():
    R, RN = oracle.ref(), oracle.ref(noblas=True)
    g = {}
    rng = np.random.default_rng(20261017)

    # ---- the reference's own golden vector, tests/testbasic.cu:3-12 + tests/basic.testdata
    A = F([[1, 2, 3], [3, 4, 5]])
    B = F([[6, 7, 8], [9, 10, 11]])
    g["basic_A"], g["basic_B"] = A, B
    g["basic_expr"] = R.testbasic_expr(A, B)

    # ---- elementwise
    x = np.concatenate([
        (rng.standard_normal(4000) * 3).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 1e-45, 20.0, -20.0, 50.0, -50.0, 87.5, -87.5, 88.9, -104.0,
                  0.5, -0.5, 8.9, -8.9, 1e-4, -1e-4, 3.4e38, -3.4e38], dtype=np.float32),
        R.randn(123, 73),
    ]).astype(np.float32)
    g["ew_x"] = x
    for op in ["exp", "tanh", "dtanh", "square", "relu", "drelu"]:
        g["ew_" + op] = R.unary(op, x)
    xp = np.abs(x[np.isfinite(x)]) + np.float32(1e-3)
    xp = np.concatenate([xp, np.array([1.0, 1e-38, 1e-45, 3.4e38, 0.99999994, 1.0000001], dtype=np.float32)])
    g["ew_xp"] = xp
    g["ew_log"] = R.unary("log", xp)
    g["ew_sqrt"] = R.unary("sqrt", xp)
    g["ew_affine"] = R.affine(x, 1.7, -0.3)
    g["ew_neg"] = R.affine(x, -1.0, 0.0)
    g["ew_div5"] = R.div_scalar(x, 5.0)
    g["ew_div4096"] = R.div_scalar(x, 4096.0)
    g["ew_eleminv1"] = R.eleminv(xp, 1.0)
    g["ew_eleminv3"] = R.eleminv(xp, 3.0)
    xs = np.clip(x, -80, 80).astype(np.float32)
    g["ew_xs"] = xs
    g["ew_chain"] = R.chain_softplus5(xs)

    # ---- binary with every flag combination (37 x 53 logical)
    A = F(rng.standard_normal((37, 53)))
    Bs = F(rng.standard_normal((37, 53)) + 2.5)
    g["bin_A"], g["bin_B"] = A, Bs
    At, Bt = F(A.T), F(Bs.T)
    for ta in (0, 1):
        for tb in (0, 1):
            a = At if ta else A
            b = Bt if tb else Bs
            g[f"bin_axpby_{ta}{tb}"] = R.axpby(a, ta, b, tb, 1.5, -2.0)
            g[f"bin_hadmd_{ta}{tb}"] = R.hadmd(a, ta, b, tb)
            g[f"bin_div_{ta}{tb}"] = R.div(a, ta, b, tb)

    # ---- reductions (OpenBLAS order and the fixed BLAS-free order)
    for name, shape in [("r1", (37, 53)), ("r2", (10, 200)), ("r3", (300, 7)), ("r4", (129, 65))]:
        M = F(rng.standard_normal(shape) * 2)
        g[f"red_{name}"] = M
        for ta in (0, 1):
            for dim in (0, 1):
                g[f"red_{name}_sum_blas_t{ta}d{dim}"] = R.sum(M, ta, dim)
                g[f"red_{name}_sum_fixed_t{ta}d{dim}"] = RN.sum(M, ta, dim)
                g[f"red_{name}_max_t{ta}d{dim}"] = R.reduce("max", M, ta, dim)
                g[f"red_{name}_stats_t{ta}d{dim}"] = R.reduce("stats", M, ta, dim)
    g["norm_x"] = np.float32(R.norm(xs))

    # ---- the reference's own elementwise/reduce fixture input (tests/testElementwiseReduceTorchDump.cu:37)
    T = F(np.array([-3.0, -1.5, -0.25, 0.0, 0.5, 1.25, 2.0, 3.5, 4.0, -2.0, 0.75, 5.0], dtype=np.float32).reshape(4, 3).T)
    g["torchdump_in"] = T
    for dim in (0, 1):
        g[f"torchdump_sum_d{dim}"] = R.reduce("stats", T, 0, dim)

    # ---- GEMM, all four flag combinations (OpenBLAS and BLAS-free)
    for name, (m, k, n) in [("g1", (64, 48, 40)), ("g2", (130, 257, 70)), ("g3", (33, 1, 17)), ("g4", (256, 128, 256))]:
        P = F(rng.standard_normal((m, k)))
        Q = F(rng.standard_normal((k, n)))
        g[f"gemm_{name}_A"], g[f"gemm_{name}_B"] = P, Q
        for ta in (0, 1):
            for tb in (0, 1):
                a = F(P.T) if ta else P
                b = F(Q.T) if tb else Q
                g[f"gemm_{name}_blas_{ta}{tb}"] = R.gemm(a, ta, b, tb)
                g[f"gemm_{name}_fixed_{ta}{tb}"] = RN.gemm(a, ta, b, tb)
    # n = 1001-style odd leading dimension (tests/testEigen.cu) at a small size
    P = F(rng.standard_normal((101, 101)))
    Q = F(rng.standard_normal((101, 101)))
    g["gemm_odd_A"], g["gemm_odd_B"] = P, Q
    g["gemm_odd_blas_01"] = R.gemm(P, 0, F(Q.T), 1)

    # ---- config 1 at a CPU-sized n: log(exp(A*B/n)+1)/5
    n = 192
    P, Q = F(R.randn(1, n * n).reshape(n, n, order="F")), F(R.randn(2, n * n).reshape(n, n, order="F"))
    g["c1_A"], g["c1_B"] = P, Q
    g["c1_out"] = R.config1(P, Q, float(n))

    # ---- data movement
    M = F(rng.standard_normal((23, 31)))
    g["mv_M"] = M
    g["mv_T"] = R.materialize(M, 1)
    g["mv_slice"] = R.slice(M, 0, 3, 20, 5, 30)
    g["mv_sliceT"] = R.slice(M, 1, 3, 20, 5, 19)
    S = F(rng.standard_normal((4, 6)))
    g["mv_S"] = S
    g["mv_set"] = R.slice_set(M, 0, 2, 6, 3, 9, S, 0)
    g["mv_setT"] = R.slice_set(M, 1, 2, 8, 3, 7, S, 1)
    g["mv_set_mixed"] = R.slice_set(M, 0, 2, 8, 3, 7, S, 1)
    N1 = F(rng.standard_normal((31, 23)))
    N2 = F(rng.standard_normal((23, 5)))
    g["mv_N1"], g["mv_N2"] = N1, N2
    g["mv_hstack"] = R.stack(0, [(M, 0), (N1, 1), (N2, 0)])
    N3 = F(rng.standard_normal((4, 31)))
    g["mv_N3"] = N3
    g["mv_vstack"] = R.stack(1, [(M, 0), (N1, 1), (N3, 0)])
    assert g["mv_hstack"] is not None and g["mv_vstack"] is not None
    # literal expectations of tests/testbasic.cu:58-82 (test3)
    A3 = F([[1, 1, 3], [1, 1, 5]])
    B3 = F([[-1, -1, -1], [9, 10, 11]])
    g["t3_A"], g["t3_B"] = A3, B3
    g["t3_vstack"] = R.stack(1, [(A3, 0), (B3, 0)])
    g["t3_hstack"] = R.stack(0, [(A3, 0), (B3, 0)])

    # ---- softmax head (ml/layer.hpp:252-264)
    X = F(rng.standard_normal((10, 33)) * 3)
    Y = np.zeros((10, 33), dtype=np.float32, order="F")
    Y[rng.integers(0, 10, 33), np.arange(33)] = 1
    g["sm_X"], g["sm_Y"] = X, Y
    g["sm_softmax"] = R.softmax_cols(X)
    g["sm_cegrad"] = R.softmax_ce_grad(X, Y, 32)
    X2 = F(rng.standard_normal((100, 17)) * 2)
    g["sm_X2"] = X2
    g["sm_softmax2"] = R.softmax_cols(X2)

    # ---- seeded streams as the reference draws them (cpp/matrix.hpp:49-71)
    g["rng_randn_0"] = R.randn(0, 1001)
    g["rng_rand_7"] = R.rand(7, 1001)

    out = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(out, **g)
    print(f"wrote {out}: {len(g)} arrays, {os.path.getsize(out)/1024:.0f} KiB; BLAS: {R.blas_config()}")


if __name__ == "__main__":
    main()

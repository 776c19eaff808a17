"""One launch of every hot-path kernel at bench size, for `ncu` captures (never a bench number).

    ncu --set full --clock-control none --import-source on -k regex:<pat> -o gpurun_out/prof \
        python scripts/ncu_ops.py [log2n] [gemm_n]
"""
import sys

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
gn = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
L = jz.lib()
assert L.jz_init(0) == 0, L.jz_last_error()
rows = cols = 1 << (log2n // 2)
n = rows * cols
CM, U = jz.CM, jz._lib.UNARY
X, P, Y, T = CM.randn(rows, cols, seed=0), CM.rand(rows, cols, seed=1), CM.randn(rows, cols, seed=2), CM.empty("T", rows, cols)
L.jz_affine(P.ptr, P.ptr, n, 1.0, 0.5, None)
v0, v1 = CM.empty("v0", cols, 1), CM.empty("v1", rows, 1)
chain, nchain = jz._lib.make_steps([("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", 0.20000000298023224, 0.0)])
for _ in range(1):  # ncu replays each launch itself; no warm-up pass needed
    L.jz_fill(T.ptr, n, 1.0, None)
    for op in ("exp", "log", "tanh", "dtanh", "square"):
        L.jz_unary(U[op], T.ptr, P.ptr if op == "log" else X.ptr, n, None)
    L.jz_affine(T.ptr, T.ptr, n, 2.0, 1.0, None)
    L.jz_eleminv(T.ptr, P.ptr, n, 1.0, None)
    L.jz_axpby(T.ptr, X.ptr, Y.ptr, n, 1.0, -1.0, None)
    L.jz_hadamard(T.ptr, X.ptr, Y.ptr, n, None)
    L.jz_chain(T.ptr, X.ptr, n, chain, nchain, None)
    L.jz_sum(v0.ptr, X.ptr, rows, cols, rows, 0, None)
    L.jz_sum(v1.ptr, X.ptr, rows, cols, rows, 1, None)
    L.jz_max(v0.ptr, X.ptr, rows, cols, rows, 0, None)
    L.jz_softmax_cols(T.ptr, X.ptr, rows, cols, rows, None)
    L.jz_copy2d(T.ptr, cols, X.ptr, rows, cols, rows, 1, None)
    jz.sync()
del X, P, Y, T
a, b, c = CM.randn(gn, gn, seed=11), CM.randn(gn, gn, seed=12), CM.empty("c", gn, gn)
for _ in range(1):
    for mode in (0, 1):
        assert L.jz_gemm(0, 0, gn, gn, gn, 1.0, a.ptr, gn, b.ptr, gn, 0.0, c.ptr, gn, mode, None) == 0
    jz.sync()
print("ncu_ops done, launches:", L.jz_launch_count())

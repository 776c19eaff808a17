"""Arbitrary-shape GEMM timing through the C ABI (CUDA events), both tensor modes, with the unit plan of each launch
(k-splits of the tail tiles) and torch.matmul (cuBLAS, allow_tf32 on / off) on the same box as the comparison bar.

    python scripts/gemm_shapes.py [m,n,k[,ta,tb] ...]      # default: small squares, skinny, demo_mnist at large batch
"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"] / 2
DEFAULT = [(1024, 1024, 1024), (1536, 1536, 1536), (2048, 2048, 2048), (3072, 3072, 3072), (4096, 4096, 4096),
           (8192, 32, 8192), (32, 8192, 8192), (4096, 48, 4096), (16384, 16, 16384),
           # demo_mnist 784-1024-128-10 at batch N: forward W*X, backward W^T*delta, weight gradient delta*X^T
           (1024, 8192, 784), (128, 8192, 1024), (10, 8192, 128), (1024, 784, 8192, 0, 1), (128, 1024, 8192, 0, 1),
           (784, 8192, 1024, 1, 0), (1024, 60000, 784), (1024, 784, 60000, 0, 1), (128, 1024, 60000, 0, 1)]
shapes = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or DEFAULT


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rng = np.random.default_rng(0)
for sh in shapes:
    m, n, k = sh[:3]
    ta, tb = (sh[3], sh[4]) if len(sh) == 5 else (0, 0)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    a, b, c = jz.CM.randn(ar, ac, seed=1), jz.CM.randn(br, bc, seed=2), jz.CM.empty("c", m, n)
    flops = 2.0 * m * n * k
    reps = 50 if flops < 2e10 else (20 if flops < 3e11 else 5)
    # float64 spot check of 48 entries
    ii, jj = rng.integers(0, m, 48), rng.integers(0, n, 48)
    A = torch.empty(ar * ac, dtype=torch.float32, device="cuda"); L.jz_copy(A.data_ptr(), a.ptr, ar * ac, stream)
    B = torch.empty(br * bc, dtype=torch.float32, device="cuda"); L.jz_copy(B.data_ptr(), b.ptr, br * bc, stream)
    Am, Bm = A.view(ac, ar).t(), B.view(bc, br).t()          # logical (row, col) of the column-major buffers
    opA, opB = (Am.t() if ta else Am), (Bm.t() if tb else Bm)
    want = (opA[ii, :].double() * opB[:, jj].t().double()).sum(dim=1).cpu().numpy()
    line = f"{m:6d}x{n:<6d}x{k:<6d} ta={ta} tb={tb} |"
    for mode, mname in ((0, "3xtf32"), (1, "tf32")):
        def run():
            rc = L.jz_gemm(ta, tb, m, n, k, 1.0, a.ptr, ar, b.ptr, br, 0.0, c.ptr, m, mode, stream)
            assert rc == 0, L.jz_last_error()
        ms = timed(run, reps)
        Cv = torch.empty(m * n, dtype=torch.float32, device="cuda"); L.jz_copy(Cv.data_ptr(), c.ptr, m * n, stream)
        got = Cv.view(n, m).t()[ii, jj].double().cpu().numpy()
        rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
        pk = peak / (3 if mode == 0 else 1)
        line += (f" {mname}: {ms*1e3:8.1f} us {flops/ms/1e9:6.1f} TF ({flops/ms/1e9/pk:.2f}) path={L.jz_gemm_last_path()}"
                 f" splits={L.jz_gemm_last_splits()} rel={rel:.1e} |")
    old = torch.backends.cuda.matmul.allow_tf32
    for allow, key in ((True, "cublas_tf32"), (False, "cublas_fp32")):
        torch.backends.cuda.matmul.allow_tf32 = allow
        ms = timed(lambda: torch.matmul(opA, opB), reps)
        line += f" {key}: {ms*1e3:8.1f} us {flops/ms/1e9:6.1f} TF |"
    torch.backends.cuda.matmul.allow_tf32 = old
    print(line, flush=True)
    del a, b, c, A, B

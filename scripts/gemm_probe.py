"""GEMM bring-up probe (run on the GPU box, each configuration in its own process under
`timeout` so a barrier deadlock cannot take the box down):

    JZ_GEMM_CG=1 timeout 120 python scripts/gemm_probe.py tf32 256 256 256
"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
m, n, k = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (256, 256, 256)
L = jz.lib()
assert L.jz_init(0) == 0, L.jz_last_error()
rng = np.random.default_rng(0)
# small integers: exact in tf32, so any mismatch is a layout/descriptor bug, not rounding
A = np.asfortranarray(rng.integers(-4, 5, (m, k)).astype(np.float32))
B = np.asfortranarray(rng.integers(-4, 5, (k, n)).astype(np.float32))
want = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)
for ta in (1, 0):
    for tb in (0, 1):
        a = jz.CM(np.asfortranarray(A.T)).T() if ta else jz.CM(A)
        b = jz.CM(np.asfortranarray(B.T)).T() if tb else jz.CM(B)
        t0 = time.time()
        c = a.dot(b, mode=jz._lib.GEMM_MODES[mode])
        got = c.to_host()
        dt = time.time() - t0
        bad = np.argwhere(got != want)
        print(f"[{mode} m={m} n={n} k={k} ta={ta} tb={tb}] path={L.jz_gemm_last_path()} "
              f"mismatches={len(bad)}/{m*n} max_abs_err={np.abs(got-want).max():.4g} ({dt*1e3:.1f} ms)", flush=True)
        if len(bad):
            for (i, j) in bad[:6]:
                print(f"   C[{i},{j}] got {got[i,j]} want {want[i,j]}")
            rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
            print(f"   bad rows: {len(rows_bad)} (first {rows_bad[:8]}), bad cols: {len(cols_bad)} (first {cols_bad[:8]})")
# random data: rounding behaviour
A = np.asfortranarray(rng.standard_normal((m, k)).astype(np.float32))
B = np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32))
truth = A.astype(np.float64) @ B.astype(np.float64)
got = jz.CM(A).dot(jz.CM(B), mode=jz._lib.GEMM_MODES[mode]).to_host()
print(f"[{mode}] random data rel_fro = {np.linalg.norm(got-truth)/np.linalg.norm(truth):.3e}")
# timing
a, b = jz.CM(A), jz.CM(B)
for _ in range(3):
    c = a.dot(b, mode=jz._lib.GEMM_MODES[mode])
jz.sync()
t0 = time.time()
iters = 10
for _ in range(iters):
    c = a.dot(b, mode=jz._lib.GEMM_MODES[mode])
jz.sync()
dt = (time.time() - t0) / iters
print(f"[{mode}] {m}x{n}x{k}: {dt*1e3:.3f} ms/call (wall, incl. pre-pass) = {2*m*n*k/dt/1e12:.2f} TFLOP/s")

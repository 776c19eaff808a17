"""BASELINE configs[2]: fp32 GEMM sweep 1024..16384 square, A*B, A.T*B, A*B.T, 3xTF32 and TF32 modes, on 1 B200.
Times jz_gemm (pre-pass included) with CUDA events through the C ABI and checks sampled entries of every
product against a float64 recomputation.  Prints one JSON line (-> profiles/)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192, 16384]
L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
peak = json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"] / 2
out = {}
rng = np.random.default_rng(0)
for n in sizes:
    a, b, c = jz.CM.randn(n, n, seed=1), jz.CM.randn(n, n, seed=2), jz.CM.empty("c", n, n)
    # host copies of a few rows/columns for the float64 spot check
    ii = rng.integers(0, n, 64)
    jj = rng.integers(0, n, 64)
    A = torch.empty(n * n, dtype=torch.float32, device="cuda")
    B = torch.empty(n * n, dtype=torch.float32, device="cuda")
    L.jz_copy(A.data_ptr(), a.ptr, n * n, stream)
    L.jz_copy(B.data_ptr(), b.ptr, n * n, stream)
    A2, B2 = A.view(n, n).t(), B.view(n, n).t()   # logical (row, col) views of the column-major buffers
    for (ta, tb, name) in ((0, 0, "A*B"), (1, 0, "A.T*B"), (0, 1, "A*B.T")):
        opA = A2.t() if ta else A2
        opB = B2.t() if tb else B2
        want = (opA[ii, :].double() * opB[:, jj].t().double()).sum(dim=1).cpu().numpy()
        for mode, mname in ((0, "3xtf32"), (1, "tf32")):
            def run():
                rc = L.jz_gemm(ta, tb, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, c.ptr, n, mode, stream)
                assert rc == 0, L.jz_last_error()
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            reps = 20 if n <= 4096 else (8 if n <= 8192 else 4)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            Cv = torch.empty(n * n, dtype=torch.float32, device="cuda")
            L.jz_copy(Cv.data_ptr(), c.ptr, n * n, stream)
            got = Cv.view(n, n).t()[ii, jj].double().cpu().numpy()
            rel = float(np.linalg.norm(got - want) / np.linalg.norm(want))
            tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
            pk = peak / (3 if mode == 0 else 1)
            out[f"{n}:{name}:{mname}"] = {"ms": round(ms, 4), "TFLOP/s": round(tf, 1), "frac": round(tf / pk, 3),
                                           "rel_err_sampled": float(f"{rel:.3e}"), "path": L.jz_gemm_last_path()}
            print(f"{n:6d} {name:6s} {mname:7s} {ms:9.4f} ms {tf:8.1f} TFLOP/s  frac {tf/pk:.3f}  rel {rel:.2e}", flush=True)
    del a, b, c, A, B
print(json.dumps({"gemm_sweep": out, "tf32_peak_assumed": peak}))

"""Latency of the products of one demo_mnist step at batch 32 through jz_gemm, CUDA events, warm L2
(as inside the training loop).  Variants are selected with JZ_SMALL_* / JZ_GEMM_NO_SMALL in the environment."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
# (name, ta, tb, m, n, k)
SHAPES = [("fwd0  W0*x      ", 0, 0, 1024, 32, 784), ("fwd1  W1*h      ", 0, 0, 128, 32, 1024), ("fwd2  W2*h      ", 0, 0, 10, 32, 128),
          ("gW0   t*x.T     ", 0, 1, 1024, 784, 32), ("gW1   t*h.T     ", 0, 1, 128, 1024, 32), ("gW2   t*h.T     ", 0, 1, 10, 128, 32),
          ("back1 W1.T*t    ", 1, 0, 1024, 32, 128), ("back0 W0.T*t    ", 1, 0, 784, 32, 1024), ("back2 W2.T*t    ", 1, 0, 128, 32, 10)]
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("JZ_")) or "default"
tot = 0.0
rng = np.random.default_rng(0)
for name, ta, tb, m, n, k in SHAPES:
    P = np.asfortranarray(rng.standard_normal((m, k)).astype(np.float32))
    Q = np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32))
    a = jz.CM(np.asfortranarray(P.T)).T() if ta else jz.CM(P)
    b = jz.CM(np.asfortranarray(Q.T)).T() if tb else jz.CM(Q)
    c = jz.CM.empty("c", m, n)
    def run():
        rc = L.jz_gemm(ta, tb, m, n, k, 1.0, a.ptr, a.numrow, b.ptr, b.numrow, 0.0, c.ptr, m, 0, stream)
        assert rc == 0, L.jz_last_error()
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 300
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    err = np.linalg.norm(c.to_host().astype(np.float64) - P.astype(np.float64) @ Q.astype(np.float64)) / np.linalg.norm(P.astype(np.float64) @ Q)
    tot += us
    print(f"[{tag}] {name} {m:5d}x{n:4d}x{k:5d} path={L.jz_gemm_last_path()} {us:7.2f} us  rel {err:.1e}", flush=True)
print(f"[{tag}] total {tot:.1f} us")

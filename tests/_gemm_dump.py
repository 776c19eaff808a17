"""Helper of tests/test_gemm_gpu.py: one product through the C ABI in a fresh process (so that environment knobs read once
per process, e.g. JZ_GEMM_CLUSTER_SPLIT, can differ between two runs); writes C and the launch diagnostics.

    python tests/_gemm_dump.py m k n ta tb mode seed out.npz
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import juzhen_b200 as jz  # noqa: E402

m, k, n, ta, tb, mode, seed = (int(x) for x in sys.argv[1:8])
L = jz.lib()
jz._lib.check(L.jz_init(0))
rng = np.random.default_rng(seed)
P = np.asfortranarray(rng.standard_normal((k, m) if ta else (m, k)), dtype=np.float32)
Q = np.asfortranarray(rng.standard_normal((n, k) if tb else (k, n)), dtype=np.float32)
C0 = np.asfortranarray(rng.standard_normal((m, n)), dtype=np.float32)
a, b, c = jz.CM(P), jz.CM(Q), jz.CM(C0)
steps = [("affine", float(np.float32(1.0 / k)), 0.0), ("tanh",)]
arr, ns = jz._lib.make_steps(steps)
out = {}
jz._lib.check(L.jz_gemm(ta, tb, m, n, k, 0.75, a.ptr, P.shape[0], b.ptr, Q.shape[0], -0.5, c.ptr, m, mode, None))
out["axpby"] = c.to_host()
out["splits"] = np.int32(L.jz_gemm_last_splits())
out["cluster_split"] = np.int32(L.jz_gemm_last_cluster_split())
f = jz.CM.empty("f", m, n)
jz._lib.check(L.jz_gemm_chain(ta, tb, m, n, k, 1.0, a.ptr, P.shape[0], b.ptr, Q.shape[0], f.ptr, m, arr, ns, mode, None))
out["chain"] = f.to_host()
np.savez(sys.argv[8], **out)

"""One product through the instrumented library (scripts/build_prof_lib.sh): prints the per-role cycle counts of CTA 5.

    JZ_B200_LIB=build/prof/libjz_b200.so python scripts/gemm_prof_one.py m,n,k[,ta,tb[,mode]] ...
"""
import sys

import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
for arg in sys.argv[1:]:
    v = [int(x) for x in arg.split(",")]
    m, n, k = v[:3]
    ta, tb = (v[3], v[4]) if len(v) >= 5 else (0, 0)
    mode = v[5] if len(v) >= 6 else 0
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    a, b, c = jz.CM.randn(ar, ac, seed=1), jz.CM.randn(br, bc, seed=2), jz.CM.empty("c", m, n)
    # keep the GPU busy first: a single cold launch runs at idle clocks (SM and memory) and says nothing
    for rep in range(300):
        L.jz_gemm(ta, tb, m, n, k, 1.0, a.ptr, ar, b.ptr, br, 0.0, c.ptr, m, mode, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(20):
        L.jz_gemm(ta, tb, m, n, k, 1.0, a.ptr, ar, b.ptr, br, 0.0, c.ptr, m, mode, stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"--- {m}x{n}x{k} ta={ta} tb={tb} mode={mode}: {e0.elapsed_time(e1) * 1e3 / 20:.1f} us per launch; CTA 5 of the last launch:", flush=True)
    import ctypes
    buf = (ctypes.c_longlong * 32)()
    raw = ctypes.CDLL(jz._lib.LIB_PATH)
    assert raw.jz_debug_gemm_prof(buf) == 0
    g = list(buf)
    print(f"    TMA: {g[0]} k-blocks, {g[1]} clk, waiting for a free stage {g[2]}")
    print(f"    MMA: {g[3]} clk, waiting for operands {g[4]}, for a drained accumulator {g[5]}")
    for grp in (0, 1):
        q = g[8 + 8 * grp:8 + 8 * grp + 5]
        q = g[8 + 8 * grp:8 + 8 * grp + 7]
        print(f"    transform group {grp}: mainloop {q[0]} clk: waiting for TMA {q[1]}, transform {q[2]}, waiting for a chunk {q[3]}, drain {q[4]}; own iterations {q[5]}, skipped {q[6]}")
    print(f"    epilogue {g[24]} clk (split {g[25]} of {g[26]}): partial tile out {g[29]}, fence + barrier {g[30]}, ticket + waiting for the siblings {g[31]}")

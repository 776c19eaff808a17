"""a handful of GEMM launches for `ncu --set full`: n in argv (default 4096 8192) x {TF32, 3xTF32}, A*B, 2 warm-up + 1 measured each"""
import sys, torch
sys.path.insert(0, ".")
import juzhen_b200 as jz
L = jz.lib(); assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream; jz.set_stream(stream)
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
    a, b, c = jz.CM.randn(n, n, seed=1), jz.CM.randn(n, n, seed=2), jz.CM.empty("c", n, n)
    for mode in (1, 0):
        for _ in range(3):
            assert L.jz_gemm(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, c.ptr, n, mode, stream) == 0
        torch.cuda.synchronize()

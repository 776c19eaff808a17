"""column softmax over long columns (rows 8192 .. 262144, 2^28 elements): GB/s of the algorithmic 8 B/elem and fraction of
the measured HBM copy peak.  JZ_SOFTMAX_NO_PREFETCH=1 = the register-only CTA-per-column form without the cp.async
next-column prefetch (rows <= 32768); JZ_SOFTMAX_NO_CLUSTER=1 = three-pass fallback for rows > 32768.  Checks every
result against torch.softmax in float64 on a sample of columns."""
import json, os, sys, torch
sys.path.insert(0, ".")
import juzhen_b200 as jz
L = jz.lib(); assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream; jz.set_stream(stream)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
tag = "no-prefetch" if os.environ.get("JZ_SOFTMAX_NO_PREFETCH") else ("three-pass" if os.environ.get("JZ_SOFTMAX_NO_CLUSTER") else "default")
if os.environ.get("JZ_SOFTMAX_LONG_MIN"): tag += f", chunked above {os.environ['JZ_SOFTMAX_LONG_MIN']} rows"
for rows, cols in ((8192, 32768), (12288, 16384), (16384, 16384), (20480, 8192), (24576, 8192), (32768, 8192), (49152, 4096), (65536, 4096),
                   (131072, 2048), (262144, 1024)):
    x, y = jz.CM.randn(rows, cols, seed=1), jz.CM.empty("y", rows, cols)
    f = lambda: L.jz_softmax_cols(y.ptr, x.ptr, rows, cols, rows, stream)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    X = torch.empty(rows * cols, dtype=torch.float32, device="cuda"); L.jz_copy(X.data_ptr(), x.ptr, rows * cols, stream)
    Y = torch.empty(rows * cols, dtype=torch.float32, device="cuda"); L.jz_copy(Y.data_ptr(), y.ptr, rows * cols, stream)
    pick = torch.tensor([0, 1, cols // 2, cols - 1], device="cuda")
    want = torch.softmax(X.view(cols, rows)[pick].double(), dim=1)
    err = float((Y.view(cols, rows)[pick].double() - want).abs().max() / want.abs().max())
    gbs = 8.0 * rows * cols / ms / 1e6
    print(f"[{tag}] softmax_cols {rows} x {cols}: {ms:.3f} ms  {gbs:.0f} GB/s  ({gbs / peak:.2f} of the copy peak)  max rel err {err:.1e}", flush=True)
    del x, y, X, Y

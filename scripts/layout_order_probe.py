"""Is the 3xTF32 gap between A*B and the transposed layouts at 8192^3 a layout effect or the power state of the run?
Times the three layouts in several orders (each: 3 warm-up + 6 timed launches, CUDA events), printing the SM clock
nvidia-smi reports right after each timed region.   python scripts/layout_order_probe.py [n]"""
import subprocess
import sys

import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

L = jz.lib()
assert L.jz_init(0) == 0
stream = torch.cuda.current_stream().cuda_stream
jz.set_stream(stream)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a, b, c = jz.CM.randn(n, n, seed=1), jz.CM.randn(n, n, seed=2), jz.CM.empty("c", n, n)
NAMES = {(0, 0): "A*B  ", (1, 0): "A.T*B", (0, 1): "A*B.T"}


def clock():
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        return out
    except Exception:  # noqa: BLE001
        return "?"


def timed(ta, tb, mode, reps=6):
    def run():
        assert L.jz_gemm(ta, tb, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, c.ptr, n, mode, stream) == 0
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


import time
for order in ([(0, 0), (1, 0), (0, 1)], [(1, 0), (0, 1), (0, 0)], [(0, 1), (0, 0), (1, 0)]):
    time.sleep(3.0)   # let the part cool down: every order starts from the same state
    for (ta, tb) in order:
        ms = timed(ta, tb, 0)
        print(f"n={n} 3xtf32 {NAMES[(ta, tb)]} {ms:8.4f} ms {2.0 * n ** 3 / ms / 1e9:7.1f} TFLOP/s   (sm MHz, W after: {clock()})", flush=True)
    print("--")

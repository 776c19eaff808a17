"""One strided-batched product of the attention block (Q^T K: member = seq x seq, k = d_h) through the C ABI, for ncu captures
and quick timing:   python scripts/batched_probe.py [seq] [d_h] [batch] [mode] [reps]"""
import sys

import torch

sys.path.insert(0, ".")
import juzhen_b200 as jz  # noqa: E402

seq, dh, batch, mode, reps = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 512), (2, 128), (3, 128), (4, 1), (5, 20)))
L = jz.lib()
assert L.jz_init(0) == 0
s = torch.cuda.current_stream().cuda_stream
jz.set_stream(s)
q = jz.CM.randn(dh, seq * batch, seed=1)
k = jz.CM.randn(dh, seq * batch, seed=2)
out = jz.CM.empty("s", seq, seq * batch)


def run():
    rc = L.jz_gemm_strided_batched(1, 0, seq, seq, dh, 0.088, q.ptr, dh, dh * seq, k.ptr, dh, dh * seq, 0.0, out.ptr, seq, seq * seq, batch, mode, s)
    assert rc == 0, L.jz_last_error()


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
byts = 4.0 * (2 * dh * seq + seq * seq) * batch
print(f"Q^T K seq {seq} d_h {dh} batch {batch} mode {mode}: {ms * 1e3:.1f} us, {2.0 * seq * seq * dh * batch / ms / 1e9:.1f} TFLOP/s, "
      f"{byts / ms / 1e6:.0f} GB/s of operand + result traffic")

"""2-GPU worker (NCCL) for tests/test_mg_gpu.py: column-sharded GEMM + fused chain, gathered by NCCL and by the
fused P2P epilogue, checked bit-for-bit against the single-GPU product through the same C ABI."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    import juzhen_b200 as jz
    from juzhen_b200 import mg
    L = jz.lib()
    jz._lib.check(L.jz_init(local))
    stream = torch.cuda.current_stream().cuda_stream
    jz.set_stream(stream)
    ok = True
    msgs = []
    for (m, n, k) in ((1024, 1024, 1024), (640, 1000, 520)):
        rng = np.random.default_rng(5)
        A = np.asfortranarray(rng.standard_normal((m, k)).astype(np.float32))
        B = np.asfortranarray(rng.standard_normal((k, n)).astype(np.float32))
        steps = [("affine", 1.0 / k, 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",)]
        a, bfull = jz.CM(A), jz.CM(B)
        arr, ns = jz._lib.make_steps(steps)
        ref = jz.CM.empty("ref", m, n)
        jz._lib.check(L.jz_gemm_chain(0, 0, m, n, k, 1.0, a.ptr, m, bfull.ptr, k, ref.ptr, m, arr, ns, 0, stream))
        want = ref.to_host()
        j0, j1 = mg.block_range(n, world, rank)
        b = jz.CM(np.asfortranarray(B[:, j0:j1]))
        for mode in ("nccl", "fused"):
            g = mg.GpuShardedGemm(jz, m, n, k, steps=steps, gemm_mode=0, mode=mode)
            c = g.run(a.ptr, m, 0, b.ptr, k, stream)
            torch.cuda.synchronize()
            got = c.cpu().numpy().reshape(m, n, order="F")
            same = bool(np.array_equal(got, want))
            msgs.append(f"rank{rank} {m}x{n}x{k} {mode}: bit-exact={same}")
            ok &= same
            del g
    # column sums over ROW-sharded data: local jz_sum + one all-reduce == the sum over the stacked matrix
    rows, cols = 4096, 300
    rng = np.random.default_rng(100 + rank)
    Xr = np.asfortranarray(rng.standard_normal((rows, cols)).astype(np.float32))
    x = jz.CM(Xr)
    part = torch.empty(cols, dtype=torch.float32, device="cuda")
    jz._lib.check(L.jz_sum(part.data_ptr(), x.ptr, rows, cols, rows, 0, stream))
    mg.allreduce_partial_sums(part)
    want = np.zeros(cols, dtype=np.float64)
    for r in range(world):
        want += np.random.default_rng(100 + r).standard_normal((rows, cols)).astype(np.float32).sum(axis=0, dtype=np.float64)
    got = part.cpu().numpy().astype(np.float64)
    same = bool(np.all(np.abs(got - want) <= 1e-5 * np.sqrt(rows * world) * 4))
    msgs.append(f"rank{rank} row-sharded column sums + all-reduce: ok={same}")
    ok &= same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print("\n".join(msgs), flush=True)
    dist.destroy_process_group()
    if rank == 0:
        print("MG_GPU_OK" if int(flag.item()) == 1 else "MG_GPU_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

"""N-GPU worker (torchrun) for tests/test_mg_gpu.py: column-sharded GEMM + fused chain, gathered by NCCL, by the fused
P2P-store epilogue (symmetric memory), by the multicast epilogue (NVSwitch multimem.st) and through the torch-free
jz_mg_* ABI (CUDA IPC).  Every mode must give the SAME bits (same local kernel, same block), replicas must be
identical, and the result must match the single-GPU product of the same operands (which may use another split-K plan:
2e-6 relative, not bitwise)."""
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
        first = None
        for mode in ("nccl", "fused", "mcast", "ipc"):
            try:
                g = mg.GpuShardedGemm(jz, m, n, k, steps=steps, gemm_mode=0, mode=mode)
            except RuntimeError as e:
                if mode == "mcast":
                    msgs.append(f"rank{rank} {m}x{n}x{k} mcast: unavailable ({e})")
                    continue
                raise
            for _ in range(2):   # twice: the second run exercises the pre-store barrier / image reuse
                c = g.run(a.ptr, m, 0, b.ptr, k, stream)
            torch.cuda.synchronize()
            got = c.cpu().numpy().reshape(m, n, order="F")
            rel = float(np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want))
            if first is None:
                first = got
            same = bool(np.array_equal(got, first))
            chk = torch.from_numpy(got.view(np.int32).ravel()).sum(dtype=torch.int64).cuda()
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            repl = bool(lo.item() == hi.item())
            msgs.append(f"rank{rank} {m}x{n}x{k} {mode}: rel vs single-GPU {rel:.2e}, same bits as nccl={same}, replicas identical={repl}")
            ok &= same and repl and rel < 2e-6
            g.close()
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

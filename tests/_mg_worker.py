"""world_size-2 gloo worker for tests/test_mg_cpu.py: exercises the multi-GPU HOST logic of juzhen_b200/mg.py
(partition, placement, gather, partial-sum all-reduce) on CPU.  The compute is the oracle here (this is a
test); on the B200 box the same logic drives libjz_b200.so."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from juzhen_b200 import mg  # noqa: E402
import oracle  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    P = oracle.port()
    ok = True
    for (m, n, k) in ((48, 64, 40), (33, 51, 17)):  # even and ragged column splits
        A = np.asfortranarray(P.randn(7, m * k).reshape(m, k, order="F"))
        B = np.asfortranarray(P.randn(8, k * n).reshape(k, n, order="F"))
        plan = mg.ShardedGemmPlan(m, n, k, world, rank)
        c_full = torch.zeros(m * n, dtype=torch.float32)

        def compute_block(j0, j1, out):
            blk = P.gemm(A, 0, np.asfortranarray(B[:, j0:j1]), 0)          # m x (j1-j0)
            blk = P.chain_softplus5(np.ascontiguousarray(blk.ravel(order="F")))
            out.copy_(torch.from_numpy(np.ascontiguousarray(blk)))

        mg.sharded_gemm(plan, compute_block, c_full)
        want = P.chain_softplus5(np.ascontiguousarray(P.gemm(A, 0, B, 0).ravel(order="F")))
        got = c_full.numpy()
        # OpenBLAS-free port: block products are bit-identical to the corresponding columns of the full product
        ok &= bool(np.array_equal(got, want))
        # column sums over ROW-sharded X: partial vectors + all-reduce
        X = np.asfortranarray(P.randn(9, m * n).reshape(m, n, order="F"))
        r0, r1 = mg.block_range(m, world, rank)
        part = torch.from_numpy(X[r0:r1, :].sum(axis=0, dtype=np.float64).astype(np.float32))
        tot = mg.allreduce_partial_sums(part)
        ok &= bool(np.allclose(tot.numpy(), X.sum(axis=0, dtype=np.float64), rtol=1e-5, atol=1e-5))
    # flat partition covers the buffer exactly once with 16-byte aligned starts
    for count in (0, 1, 5, 1024, 1027):
        spans = [mg.flat_partition(count, world, r) for r in range(world)]
        ok &= spans[0][0] == 0 and spans[-1][1] == count
        ok &= all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        ok &= all(s[0] % 4 == 0 or s[0] == count for s in spans)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MG_WORKER_OK" if int(flag) == 1 else "MG_WORKER_FAIL")
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()

"""Multi-GPU host logic for the Matrix<CUDAfloat> path: one process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on the B200 box, gloo in the CPU tests), the C ABI for every byte of compute.

The reference has no collectives at all (grep for nccl|mpi|cudaSetDevice finds nothing);
this module is what section 8(e) adds:

  elementwise / fused chains      flat contiguous partition of the physical buffer, NO collective
  column sums of row-sharded X    local jz_sum + one all-reduce of a length-ncols vector
  row sums of column-sharded X    local jz_sum + one all-reduce of a length-nrows vector
  GEMM  C = chain(A * B)          B and C sharded by COLUMNS (contiguous in column-major storage; this is the
                                  "row-sharded" problem of north_star seen through the transpose): every rank
                                  holds A, computes C[:, j0:j1] with the chain fused in the GEMM epilogue, then
                                  either  (a) one NCCL all-gather of the finished blocks, or
                                          (b) FUSED: the epilogue stores each finished tile straight into every
                                              peer's image of C (jz_gemm_chain_bcast, P2P stores over NVLink
                                              through torch symmetric memory), followed by one barrier.
  transpose / slice / stack       replicas only (would need an all-to-all; not on the path)

Compute is injected as a callable so the partition / placement logic can be exercised on CPU with gloo
(tests/test_mg_cpu.py uses the oracle as the compute there; the product path uses libjz_b200.so).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist


def block_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """contiguous [begin, end) of rank's share of n items; the first n % world ranks get one more"""
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def flat_partition(count: int, world: int, rank: int, align: int = 4) -> tuple[int, int]:
    """elementwise work: contiguous slice of the flat physical buffer, begin aligned to `align` elements
    (16 bytes for fp32) so every shard keeps the 128-bit access path"""
    units = (count + align - 1) // align
    b, e = block_range(units, world, rank)
    return min(b * align, count), min(e * align, count)


@dataclass
class ShardedGemmPlan:
    """C (m x n) = chain(op(A)(m x k) * B(k x n)), B and C column-sharded over `world` ranks"""
    m: int
    n: int
    k: int
    world: int
    rank: int

    @property
    def cols(self) -> tuple[int, int]:
        return block_range(self.n, self.world, self.rank)

    @property
    def even(self) -> bool:
        return self.n % self.world == 0

    def c_offset(self, rank=None) -> int:
        """element offset of a rank's block inside the full column-major C"""
        j0, _ = block_range(self.n, self.world, self.rank if rank is None else rank)
        return j0 * self.m

    def c_count(self, rank=None) -> int:
        j0, j1 = block_range(self.n, self.world, self.rank if rank is None else rank)
        return (j1 - j0) * self.m


def all_gather_blocks(c_full: torch.Tensor, plan: ShardedGemmPlan, group=None) -> None:
    """gather every rank's finished column block into the full C held by each rank (in place: the local
    block already sits at its final position).  Even split = one all_gather_into_tensor on views of the
    same buffer; ragged split = one broadcast per block."""
    flat = c_full.view(-1)
    if plan.world == 1:
        return
    if plan.even:
        mine = flat.narrow(0, plan.c_offset(), plan.c_count())
        dist.all_gather_into_tensor(flat, mine, group=group)
        return
    for r in range(plan.world):
        dist.broadcast(flat.narrow(0, plan.c_offset(r), plan.c_count(r)), src=r, group=group)


def sharded_gemm(plan: ShardedGemmPlan, compute_block, c_full: torch.Tensor, group=None) -> torch.Tensor:
    """compute_block(j0, j1, out_flat_view) must write chain(A * B[:, j0:j1]) column-major into the view"""
    j0, j1 = plan.cols
    compute_block(j0, j1, c_full.view(-1).narrow(0, plan.c_offset(), plan.c_count()))
    all_gather_blocks(c_full, plan, group)
    return c_full


def allreduce_partial_sums(partial: torch.Tensor, group=None) -> torch.Tensor:
    """column sums over row-sharded data / row sums over column-sharded data: fp32 sum of per-rank partial
    vectors (tiny message; NVLS in-switch reduction when NCCL offers it)"""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


# ------------------------------------------------------------------------------------------ GPU side
class GpuShardedGemm:
    """the product path on B200s: local block through jz_gemm_chain (mode 'nccl') or fused with the gather
    through jz_gemm_chain_bcast over symmetric memory (mode 'fused')."""

    def __init__(self, jz, m, n, k, steps=(), gemm_mode=-1, mode="nccl", group=None):
        self.jz, self.L = jz, jz.lib()
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.plan = ShardedGemmPlan(m, n, k, world, rank)
        self.steps, self.nsteps = jz._lib.make_steps(list(steps))
        self.gemm_mode, self.mode = gemm_mode, mode
        dev = torch.device("cuda", torch.cuda.current_device())
        self.peer_ptrs = None
        if mode == "fused" and world > 1:
            import torch.distributed._symmetric_memory as symm
            self.c_full = symm.empty(m * n, dtype=torch.float32, device=dev)
            self.hdl = symm.rendezvous(self.c_full, group=group if group is not None else dist.group.WORLD)
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            off = 4 * self.plan.c_offset()
            others = [ptrs[r] + off for r in range(world) if r != rank]
            if len(others) > 7:
                raise ValueError("fused gather supports up to 8 GPUs (JZ_MAX_PEERS = 7)")
            self.peer_ptrs = (ctypes.c_void_p * len(others))(*others)
            self.n_peers = len(others)
        else:
            self.c_full = torch.empty(m * n, dtype=torch.float32, device=dev)

    def run(self, a_ptr, lda, trans_a, b_block_ptr, ldb, stream):
        """A: full operand (trans_a = its lazy transpose flag); b_block: this rank's k x (j1-j0) column block"""
        p = self.plan
        j0, j1 = p.cols
        c_ptr = self.c_full.data_ptr() + 4 * p.c_offset()
        if self.peer_ptrs is not None:
            rc = self.L.jz_gemm_chain_bcast(int(trans_a), 0, p.m, j1 - j0, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb,
                                            c_ptr, p.m, self.peer_ptrs, self.n_peers, self.steps, self.nsteps,
                                            self.gemm_mode, stream)
            self.jz._lib.check(rc)
            self.hdl.barrier(channel=0)  # every rank's stores have landed in every image of C
        else:
            rc = self.L.jz_gemm_chain(int(trans_a), 0, p.m, j1 - j0, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb,
                                      c_ptr, p.m, self.steps, self.nsteps, self.gemm_mode, stream)
            self.jz._lib.check(rc)
            all_gather_blocks(self.c_full, p, self.group)
        return self.c_full

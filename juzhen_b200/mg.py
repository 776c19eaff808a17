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
def _on_stream(stream):
    """torch work (barriers, NCCL collectives) issued inside this context runs on the SAME stream the jz_* kernels were
    launched on, so the two are ordered without events"""
    if stream is None or int(stream) == int(torch.cuda.current_stream().cuda_stream):
        import contextlib
        return contextlib.nullcontext()
    return torch.cuda.stream(torch.cuda.ExternalStream(int(stream)))


class GpuShardedGemm:
    """the product path on B200s.  Every mode computes the local column block with the tcgen05 GEMM + fused chain; they
    differ in how the blocks reach the other ranks:

      'nccl'   one NCCL all-gather after the local product (the baseline)
      'fused'  the GEMM epilogue stores every finished tile into each peer's image of C (P2P stores over NVLink through
               torch symmetric memory, jz_gemm_chain_bcast), then one barrier
      'mcast'  the epilogue issues ONE multimem.st per element to the NVSwitch multicast mapping of C
               (jz_gemm_chain_mcast): 1/N of the NVLink egress of 'fused'
      'ipc'    like 'fused', but through the torch-free jz_mg_* ABI: CUDA IPC handles (carried here by
               torch.distributed.all_gather_object, nothing else of torch), jz_mg_gemm_allgather, jz_mg_barrier

    C is reused by successive run() calls: every fused mode puts a barrier BEFORE the stores as well, so that no rank
    overwrites an image a peer is still reading from the previous call."""

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
        self.mc_ptr = None
        self.ipc = None
        if world == 1:
            self.mode = "local"
        if self.mode in ("fused", "mcast"):
            import torch.distributed._symmetric_memory as symm
            self.c_full = symm.empty(m * n, dtype=torch.float32, device=dev)
            self.hdl = symm.rendezvous(self.c_full, group=group if group is not None else dist.group.WORLD)
            off = 4 * self.plan.c_offset()
            if self.mode == "mcast":
                mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
                if mc == 0:
                    raise RuntimeError("this system offers no NVSwitch multicast mapping for symmetric memory")
                self.mc_ptr = mc + off
            else:
                ptrs = [int(p) for p in self.hdl.buffer_ptrs]
                others = [ptrs[r] + off for r in range(world) if r != rank]
                if len(others) > 7:
                    raise ValueError("fused gather supports up to 8 GPUs (JZ_MAX_PEERS = 7)")
                self.peer_ptrs = (ctypes.c_void_p * len(others))(*others)
                self.n_peers = len(others)
        elif self.mode == "ipc":
            self.c_full = torch.empty(m * n, dtype=torch.float32, device=dev)
            self.flags = torch.zeros(64, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            self.ipc = {"c": self._map_all(self.c_full.data_ptr(), world, rank), "f": self._map_all(self.flags.data_ptr(), world, rank)}
            self.epoch = 0
        else:
            self.c_full = torch.empty(m * n, dtype=torch.float32, device=dev)

    def _map_all(self, ptr, world, rank):
        """export `ptr`, gather every rank's handle (the only use of torch.distributed on this path), import the peers'"""
        from ._lib import MG_HANDLE_BYTES, check
        mine = ctypes.create_string_buffer(MG_HANDLE_BYTES)
        check(self.L.jz_mg_export(ptr, mine))
        allh = [None] * world
        dist.all_gather_object(allh, bytes(mine.raw), group=self.group)
        out = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                out[r] = ptr
            else:
                p = ctypes.c_void_p()
                check(self.L.jz_mg_import(ctypes.create_string_buffer(allh[r], MG_HANDLE_BYTES), ctypes.byref(p)))
                out[r] = p.value
        return out

    def close(self):
        if self.ipc is not None:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            for arr in self.ipc.values():
                for r in range(self.plan.world):
                    if r != self.plan.rank and arr[r]:
                        self.L.jz_mg_release(arr[r])
            self.ipc = None

    def _ipc_barrier(self, stream):
        self.epoch += 1
        self.jz._lib.check(self.L.jz_mg_barrier(self.ipc["f"], self.plan.world, self.plan.rank, self.epoch, stream))

    def run(self, a_ptr, lda, trans_a, b_block_ptr, ldb, stream):
        """A: full operand (trans_a = its lazy transpose flag); b_block: this rank's k x (j1-j0) column block"""
        p = self.plan
        j0, j1 = p.cols
        c_ptr = self.c_full.data_ptr() + 4 * p.c_offset()
        check = self.jz._lib.check
        if self.mode in ("fused", "mcast"):
            with _on_stream(stream):
                self.hdl.barrier(channel=1)      # every rank is done reading the previous contents of every image
            if self.mode == "mcast":
                rc = self.L.jz_gemm_chain_mcast(int(trans_a), 0, p.m, j1 - j0, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb, c_ptr, p.m,
                                                self.mc_ptr, self.steps, self.nsteps, self.gemm_mode, stream)
            else:
                rc = self.L.jz_gemm_chain_bcast(int(trans_a), 0, p.m, j1 - j0, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb, c_ptr, p.m,
                                                self.peer_ptrs, self.n_peers, self.steps, self.nsteps, self.gemm_mode, stream)
            check(rc)
            with _on_stream(stream):
                self.hdl.barrier(channel=0)      # every rank's stores have landed in every image of C
        elif self.mode == "ipc":
            self._ipc_barrier(stream)
            check(self.L.jz_mg_gemm_allgather(int(trans_a), p.m, p.n, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb, self.ipc["c"], p.world, p.rank,
                                              self.steps, self.nsteps, self.gemm_mode, stream))
            self._ipc_barrier(stream)
        else:
            rc = self.L.jz_gemm_chain(int(trans_a), 0, p.m, j1 - j0, p.k, 1.0, a_ptr, lda, b_block_ptr, ldb,
                                      c_ptr, p.m, self.steps, self.nsteps, self.gemm_mode, stream)
            check(rc)
            if p.world > 1:
                with _on_stream(stream):
                    all_gather_blocks(self.c_full, p, self.group)
        return self.c_full

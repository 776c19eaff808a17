"""CPU (gloo, world_size 2) tests of the multi-GPU host logic in juzhen_b200/mg.py."""
import os
import socket
import subprocess
import sys

from juzhen_b200 import mg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_range_properties():
    for n in (0, 1, 7, 8, 1001, 32768):
        for world in (1, 2, 3, 4, 8):
            spans = [mg.block_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_plan_offsets():
    p = mg.ShardedGemmPlan(m=10, n=7, k=3, world=2, rank=1)
    assert p.cols == (4, 7) and p.c_offset() == 40 and p.c_count() == 30 and not p.even
    assert mg.ShardedGemmPlan(10, 8, 3, 2, 0).even


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_gemm_and_partial_sums_world2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.join(ROOT, "tests", "_mg_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "MG_WORKER_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])

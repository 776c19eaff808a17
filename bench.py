#!/usr/bin/env python
"""bench.py -- headline benchmark of the Matrix<CUDAfloat> hot path (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is ONE pass of the elementwise + row/col-reduction sweep (SWEEP below: the ops of
SURVEY.md section 8a) over one batch of synthetic fp32 input of 2^28 elements
(16384 x 16384, 1 GiB per operand -- larger than the 126 MB L2, so no flush is needed between
iterations).  `value` is the algorithmic HBM bytes the step moves (SURVEY 8d per-op figures)
divided by the device time of the step, in GB/s, summed over all ranks (weak scaling: every
GPU runs the same per-GPU batch, the path shards by independent elements, no collective).
`e2e` is the same step driven through the C-ABI with HOST buffers: the inputs are copied from
pinned host memory and the reduction results are read back inside the timed region.
The other BASELINE configs ride in the same line: `gemm` = configs[2] (fp32 GEMM 1024..16384, A*B / A.T*B / A*B.T,
3xTF32 and TF32 modes, TFLOP/s and fraction of the tensor roofline), `config1` = configs[0] (4096^2 A*B then
log(exp(A*B/4096)+1)/5: fused epilogue vs separate passes, checked against the reference's CPU result),
`mnist_step` = configs[3] (demo_mnist step at batch 32 / 8192 / 60000, this backend and the reference's CUDA build),
`sharded_gemm` = configs[4] (32768^3 column-sharded GEMM + fused chain: the N = 1 anchor here, strong scaling at N > 1).

--impl reference times the reference's own CPU implementation of the same step
(oracle/_ref/libjzref.so = the unmodified Matrix<float> + OpenBLAS build; falls back to the
pinned C restatement when that file did not travel) on the SAME workload (one full 2^28 step), all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2N_DEFAULT = 28
FIFTH = 0.20000000298023224  # (float)(1.0/5.0): operator/ multiplies by the rounded reciprocal

# (name, algorithmic bytes per element) -- SURVEY.md 8d
SWEEP = [
    ("fill", 4), ("exp", 8), ("log", 8), ("tanh", 8), ("d_tanh", 8), ("square", 8), ("affine_inplace", 8),
    ("eleminv", 8), ("axpby", 12), ("hadamard", 12), ("chain_softplus5", 8), ("sum_dim0", 4), ("sum_dim1", 4),
    ("max_dim0", 4), ("softmax_cols", 8), ("transpose", 8),
]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        # bf16_tflops = burst (a kernel timed alone, SM clock at its maximum); bf16_tflops_sustained = the same GEMM back to
        # back for seconds (the power cap holds the SM clock near 1.5 GHz): the denominator for launches of milliseconds
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1590.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        busy = sm[len(sm) // 2:] if sm else []          # upper half = samples under load
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- our arm
class Sweep:
    """the step, expressed on the C-ABI (raw pointers; no torch types cross the boundary)"""

    def __init__(self, jz, rows, cols, stream):
        import ctypes
        self.jz, self.L, self.ct = jz, jz.lib(), ctypes
        self.rows, self.cols, self.n, self.s = rows, cols, rows * cols, stream
        CM = jz.CM
        self.X = CM.randn(rows, cols, seed=0)                       # randn input (exp/tanh/affine/...)
        self.P = CM.rand(rows, cols, seed=1)                        # rand + 0.5 for log / eleminv
        self.L.jz_affine(self.P.ptr, self.P.ptr, self.n, 1.0, 0.5, stream)
        self.Y = CM.randn(rows, cols, seed=2)
        self.T = CM.empty("T", rows, cols)
        self.v0 = CM.empty("v0", cols, 1)
        self.v1 = CM.empty("v1", rows, 1)
        steps = [("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", FIFTH, 0.0)]
        self.chain, self.nchain = jz._lib.make_steps(steps)
        self.ops = self.make_ops(self.X.ptr, self.Y.ptr)
        self.bytes_per_step = sum(b for _, b in SWEEP) * self.n

    def make_ops(self, X, Y):
        """the 16 calls of one step on a given pair of input buffers (raw device pointers)"""
        U = self.jz._lib.UNARY
        L, P, T, n, s, r, c = self.L, self.P.ptr, self.T.ptr, self.n, self.s, self.rows, self.cols
        return {
            "fill": lambda: L.jz_fill(T, n, 1.0, s),
            "exp": lambda: L.jz_unary(U["exp"], T, X, n, s),
            "log": lambda: L.jz_unary(U["log"], T, P, n, s),
            "tanh": lambda: L.jz_unary(U["tanh"], T, X, n, s),
            "d_tanh": lambda: L.jz_unary(U["dtanh"], T, X, n, s),
            "square": lambda: L.jz_unary(U["square"], T, X, n, s),
            "affine_inplace": lambda: L.jz_affine(T, T, n, 2.0, 1.0, s),
            "eleminv": lambda: L.jz_eleminv(T, P, n, 1.0, s),
            "axpby": lambda: L.jz_axpby(T, X, Y, n, 1.0, -1.0, s),
            "hadamard": lambda: L.jz_hadamard(T, X, Y, n, s),
            "chain_softplus5": lambda: L.jz_chain(T, X, n, self.chain, self.nchain, s),
            "sum_dim0": lambda: L.jz_sum(self.v0.ptr, X, r, c, r, 0, s),
            "sum_dim1": lambda: L.jz_sum(self.v1.ptr, X, r, c, r, 1, s),
            "max_dim0": lambda: L.jz_max(self.v0.ptr, X, r, c, r, 0, s),
            "softmax_cols": lambda: L.jz_softmax_cols(T, X, r, c, r, s),
            "transpose": lambda: L.jz_copy2d(T, c, X, r, c, r, 1, s),
        }

    def step(self, ops=None):
        ops = ops or self.ops
        for name, _ in SWEEP:
            rc = ops[name]()
            if rc != 0:
                raise RuntimeError(f"{name}: {self.L.jz_last_error().decode()}")


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa = "unchanged"
    try:   # bind this rank to the CPU cores next to its GPU: pinned staging pages and the DMA then share a NUMA node
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
        numa = f"{len(os.sched_getaffinity(0))} cores near GPU {local}"
    except Exception as e:  # noqa: BLE001 - containers may forbid it; the bench runs either way
        numa = f"unchanged ({type(e).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import juzhen_b200 as jz

    L = jz.lib()
    jz._lib.check(L.jz_init(local))
    stream = torch.cuda.current_stream().cuda_stream
    jz.set_stream(stream)
    rows = cols = 1 << (args.log2n // 2)
    if args.log2n % 2:
        cols *= 2
    sw = Sweep(jz, rows, cols, stream)
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.jz_launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = L.jz_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_step, launches = timed(sw.step, args.steps, args.warmup)
    value = sw.bytes_per_step * world / (ms_step * 1e-3) / 1e9

    # ---- per-op device times (CUDA events on the launching stream), for the roofline block
    per_op = {}
    for name, bpe in SWEEP:
        fn = sw.ops[name]
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, args.steps)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = bpe * sw.n / (ms * 1e-3) / 1e9
        per_op[name] = {"ms": round(ms, 4), "GB/s": round(gbs, 1), "frac": round(gbs / pk["hbm_gbs"], 3)}
    # dominant kernel of the step = the 1R+1W map kernel family map1_v4<F> (7 of the 16 launches, ~40% of the
    # step's device time); achieved = algorithmic bytes of those launches / their summed CUDA-event time.
    fam = ["exp", "log", "tanh", "d_tanh", "square", "affine_inplace", "eleminv"]
    fam_ms = sum(per_op[k]["ms"] for k in fam)
    fam_gbs = 8 * sw.n * len(fam) / (fam_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read+write per launch of the same kernel at the same size (ncu --set full)
        k = json.load(open(tpath))["kernels"].get("map1_v4<UnaryF<0>>")
        if k and args.log2n == 28:
            traffic = k["dram_bytes_read"] + k["dram_bytes_write"]
    slowest = min(per_op, key=lambda k: per_op[k]["frac"])
    roofline = {"bound": "hbm", "kernel": "map1_v4<F> (exp, log, tanh, d_tanh, square, affine, eleminv: 7 of 16 launches)",
                "achieved": round(fam_gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(fam_gbs / pk["hbm_gbs"], 3),
                "traffic": traffic, "peak_source": pk["source"], "bytes_per_launch": 8 * sw.n,
                "share_of_step_ms": round(fam_ms / sum(v["ms"] for v in per_op.values()), 3),
                "slowest_kernel": {"op": slowest, **per_op[slowest]}}

    # ---- e2e: same step through the C-ABI with host buffers (pinned), H2D + D2H inside the timed region
    hx = torch.empty(sw.n, dtype=torch.float32, pin_memory=True).normal_()
    hy = torch.empty(sw.n, dtype=torch.float32, pin_memory=True).normal_()
    h0 = torch.empty(cols, dtype=torch.float32, pin_memory=True)
    h1 = torch.empty(rows, dtype=torch.float32, pin_memory=True)

    # Double-buffered: the H2D copies of step i+1 run on a copy stream while step i computes (steps are
    # independent batches, so this is the throughput a streaming caller gets); both copies and the D2H reads of
    # every step are inside the timed region, ordered by events -- no host synchronisation inside a step.
    copy_stream = torch.cuda.Stream()
    cs = copy_stream.cuda_stream
    X2, Y2 = jz.CM.empty("X2", rows, cols), jz.CM.empty("Y2", rows, cols)
    sets = [(sw.X.ptr, sw.Y.ptr, sw.ops), (X2.ptr, Y2.ptr, sw.make_ops(X2.ptr, Y2.ptr))]
    loaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    main_stream = torch.cuda.current_stream()
    state = {"i": 0, "primed": False}

    def upload(k):
        copy_stream.wait_event(consumed[k])        # the step that last read this buffer set has finished
        L.jz_memcpy_h2d(sets[k][0], hx.data_ptr(), sw.n, cs)
        L.jz_memcpy_h2d(sets[k][1], hy.data_ptr(), sw.n, cs)
        loaded[k].record(copy_stream)

    def e2e_step():
        k = state["i"] & 1
        if not state["primed"]:
            consumed[0].record(main_stream); consumed[1].record(main_stream)
            upload(k)
            state["primed"] = True
        upload(k ^ 1)                               # next step's inputs, overlapping this step's compute
        main_stream.wait_event(loaded[k])
        sw.step(sets[k][2])
        consumed[k].record(main_stream)
        L.jz_memcpy_d2h(h0.data_ptr(), sw.v0.ptr, cols, stream)
        L.jz_memcpy_d2h(h1.data_ptr(), sw.v1.ptr, rows, stream)
        state["i"] += 1

    e2e_steps = max(2, min(args.steps, 5))
    ms_e2e, _ = timed(e2e_step, e2e_steps, 1)
    e2e = {"value": round(sw.bytes_per_step * world / (ms_e2e * 1e-3) / 1e9, 2), "unit": "GB/s",
           "h2d_bytes_per_step": 2 * 4 * sw.n, "d2h_bytes_per_step": 4 * (rows + cols), "ms_per_step": round(ms_e2e, 3),
           "note": "PCIe-bound: 2 GiB of pinned host input per step; uploads double-buffered against compute",
           "cpu_affinity": numa,
           "h2d_GB/s": round(2 * 4 * sw.n / (ms_e2e * 1e-3) / 1e9, 1)}

    # ---- GEMM half of the metric (BASELINE configs[2]): n x {A*B, A.T*B, A*B.T} x {3xTF32, TF32}
    gemm = {}
    if not args.no_gemm:
        for gn in args.gemm_n:
            a, b = jz.CM.randn(gn, gn, seed=11), jz.CM.randn(gn, gn, seed=12)
            c = jz.CM.empty("c", gn, gn)
            reps = max(3, args.steps // 2) if gn >= 8192 else max(10, args.steps)
            for ta, tb, suffix in ((0, 0, ""), (1, 0, "_ATB"), (0, 1, "_ABT")):
                for mode_name, mode in (("3xtf32", 0), ("tf32", 1)):
                    def g():
                        rc = L.jz_gemm(ta, tb, gn, gn, gn, 1.0, a.ptr, gn, b.ptr, gn, 0.0, c.ptr, gn, mode, stream)
                        if rc:
                            raise RuntimeError(L.jz_last_error().decode())
                    key = f"{mode_name}_{gn}{suffix}"
                    if gn >= 8192:
                        # launches of milliseconds leave the part power-capped: without a pause the configurations timed
                        # later in the list inherit a lower SM clock (8192^3 3xTF32: 293-295 TFLOP/s for every layout when each
                        # starts from idle, 243-253 for whichever comes second and third, profiles/r02h_layout_order_8192.log)
                        torch.cuda.synchronize()
                        time.sleep(1.5)
                    try:
                        ms, _ = timed(g, reps, 3)
                    except RuntimeError as e:
                        gemm[key] = {"error": str(e)}
                        continue
                    tf = 2.0 * gn ** 3 / (ms * 1e-3) / 1e12 * world
                    peak = pk["bf16_tflops"] / 2 / (3 if mode == 0 else 1)   # TF32 = bf16/2; 3xTF32 = TF32/3
                    sus = pk["bf16_tflops_sustained"] / 2 / (3 if mode == 0 else 1)
                    gemm[key] = {"TFLOP/s": round(tf, 1), "ms": round(ms, 4), "roofline_peak": round(peak * world, 1),
                                 "frac": round(tf / (peak * world), 3),
                                 # launches of milliseconds run power-capped (SM clock ~1.5 GHz): the sustained figure bounds them
                                 "frac_of_sustained_peak": round(tf / (sus * world), 3), "path": L.jz_gemm_last_path(),
                                 "k_splits_of_tail_tiles": L.jz_gemm_last_splits(),
                                 "splits_meet_in_cluster_dsmem": bool(L.jz_gemm_last_cluster_split())}
            del a, b, c
            # measurement only (never on the product path): what the vendor library reaches on this box for the same
            # product, as a cross-check of the assumed TF32 peak (MEASURED_PEAKS.json has no TF32 figure)
            try:
                old = torch.backends.cuda.matmul.allow_tf32
                ta_, tb_ = torch.randn(gn, gn, device="cuda"), torch.randn(gn, gn, device="cuda")
                for allow, key in ((True, "cublas_tf32"), (False, "cublas_fp32")):
                    torch.backends.cuda.matmul.allow_tf32 = allow
                    ms, _ = timed(lambda: torch.matmul(ta_, tb_), reps if allow else max(3, reps // 3), 2)
                    gemm[f"{key}_{gn}"] = {"TFLOP/s": round(2.0 * gn ** 3 / (ms * 1e-3) / 1e12 * world, 1), "ms": round(ms, 4),
                                           "note": "torch.matmul (cuBLAS) on the same box: comparison bar, not our kernel"}
                torch.backends.cuda.matmul.allow_tf32 = old
                del ta_, tb_
            except Exception as e:  # noqa: BLE001
                gemm[f"cublas_{gn}"] = {"error": str(e)[:120]}

    config1 = None
    if rank == 0 and world == 1 and not args.no_gemm:
        try:
            config1 = run_config1(jz, L, stream, timed, pk, no_cpu=args.no_cpu)
        except Exception as e:  # noqa: BLE001
            config1 = {"error": f"{type(e).__name__}: {e}"[:300]}

    clk = clocks.stop() if rank == 0 else None   # sampled across every timed region above (sweep, per-op, e2e, GEMM)
    sharded = None
    sharded_sum = None
    if world > 1:
        sharded_sum = run_sharded_colsum(jz, L, sw, world, stream, timed, pk)
    if not args.no_gemm:   # N = 1 runs the same product whole: the anchor of the strong-scaling curve
        sharded = run_sharded_gemm(jz, L, args, world, rank, stream, timed, pk)

    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_reference_subprocess(log2n=args.cpu_log2n, steps=1)
    mnist = None
    if rank == 0 and world == 1 and not args.no_mnist:
        try:
            mnist = mnist_step_block()
        except Exception as e:  # noqa: BLE001 - informative block only
            mnist = {"error": f"{type(e).__name__}: {e}"[:200]}

    if rank == 0:
        line = {
            "metric": "Elementwise/reduce HBM GB/s (sweep) and GEMM TFLOP/s, % of roofline", "value": round(value, 1),
            "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: elementwise + row/col-sum sweep, 2^{args.log2n} fp32 elements "
                                   f"({rows}x{cols}) per GPU, {len(SWEEP)} ops/step",
                       "l2": "inputs (1 GiB/operand) larger than L2, no flush", "ops": [n for n, _ in SWEEP],
                       "algorithmic_bytes_per_step": sw.bytes_per_step, "parallelism": f"independent shards x{world}"},
            "frac_of_hbm_peak": round(value / world / pk["hbm_gbs"], 3),
            "frac_of_8TBs_spec": round(value / world / 8000.0, 3),
            "roofline": roofline, "ops": per_op, "gemm": gemm, "config1": config1, "sharded_gemm": sharded, "sharded_colsum": sharded_sum, "mnist_step": mnist, "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clk, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_sharded_colsum(jz, L, sw, world, stream, timed, pk):
    """column sums of a matrix whose ROWS are sharded over the ranks (each rank holds rows/N x cols... here its
    own 2^28-element shard, i.e. a (N*rows) x cols matrix in total): local jz_sum over the shard + ONE all-reduce of the
    length-cols partial vector (NCCL over NVLink; NVLS in-switch reduction when available).  Weak scaling."""
    import torch
    from juzhen_b200 import mg
    part = torch.empty(sw.cols, dtype=torch.float32, device="cuda")

    def local_only():
        jz._lib.check(L.jz_sum(part.data_ptr(), sw.X.ptr, sw.rows, sw.cols, sw.rows, 0, stream))

    def with_allreduce():
        local_only()
        mg.allreduce_partial_sums(part)

    ms_local, _ = timed(local_only, 10, 3)
    ms_full, _ = timed(with_allreduce, 10, 3)
    gbs = 4.0 * sw.n * world / (ms_full * 1e-3) / 1e9
    return {"ms_local_sum": round(ms_local, 4), "ms_with_allreduce": round(ms_full, 4), "allreduce_bytes": 4 * sw.cols,
            "GB/s_aggregate": round(gbs, 1), "frac_of_hbm_peak_xN": round(gbs / (pk["hbm_gbs"] * world), 3)}


def run_config1(jz, L, stream, timed, pk, no_cpu=False):
    """BASELINE configs[0]: 4096 x 4096 fp32 A*B then log(exp(A*B/4096)+1)/5 (examples/demo_gemm.cu:41-95 style; the
    1/4096 keeps exp finite).  Three spellings on the GPU -- jz_gemm_chain (what the C++ shell
    emits for the deferred expression), GEMM + one fused elementwise pass (2 launches), GEMM + the five separate maps
    an lvalue-by-lvalue operator chain issues (6 launches) -- and the reference's own CPU result (Matrix<float>,
    OpenBLAS) on the same inputs as the comparator and the CPU bar."""
    import numpy as np
    n = 4096
    rng = np.random.default_rng(41)
    A = np.asfortranarray(rng.standard_normal((n, n), dtype=np.float32))
    B = np.asfortranarray(rng.standard_normal((n, n), dtype=np.float32))
    a, b = jz.CM(A), jz.CM(B)
    cf, cs, c5 = jz.CM.empty("cf", n, n), jz.CM.empty("cs", n, n), jz.CM.empty("c5", n, n)
    inv_n = float(np.float32(1.0 / n))
    steps, ns = jz._lib.make_steps([("affine", inv_n, 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", FIFTH, 0.0)])
    U, nn = jz._lib.UNARY, n * n

    def ck(rc):
        if rc:
            raise RuntimeError(L.jz_last_error().decode())

    def fused():
        ck(L.jz_gemm_chain(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, cf.ptr, n, steps, ns, 0, stream))

    def gemm_then_chain():
        ck(L.jz_gemm(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, cs.ptr, n, 0, stream))
        ck(L.jz_chain(cs.ptr, cs.ptr, nn, steps, ns, stream))

    def gemm_then_maps():
        ck(L.jz_gemm(0, 0, n, n, n, 1.0, a.ptr, n, b.ptr, n, 0.0, c5.ptr, n, 0, stream))
        ck(L.jz_affine(c5.ptr, c5.ptr, nn, inv_n, 0.0, stream))
        ck(L.jz_unary(U["exp"], c5.ptr, c5.ptr, nn, stream))
        ck(L.jz_affine(c5.ptr, c5.ptr, nn, 1.0, 1.0, stream))
        ck(L.jz_unary(U["log"], c5.ptr, c5.ptr, nn, stream))
        ck(L.jz_affine(c5.ptr, c5.ptr, nn, FIFTH, 0.0, stream))

    out = {"workload": "BASELINE configs[0]: 4096^2 fp32 A*B then log(exp(A*B/4096)+1)/5, 3xTF32 (fp32 accuracy)"}
    peak = pk["bf16_tflops"] / 2 / 3
    # jz_gemm_chain decides by itself: a program with transcendental steps on an output this large runs as the product
    # plus ONE streaming pass (the one-shot kernel cannot hide a 45-instruction-per-element epilogue), lighter programs
    # and multi-GPU gathers stay in the epilogue; bit-identical either way
    for name, fn in (("jz_gemm_chain", fused), ("gemm_plus_jz_chain_2_launches", gemm_then_chain),
                     ("gemm_plus_5_maps_6_launches", gemm_then_maps)):
        ms, _ = timed(fn, 10, 3)
        tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
        out[name] = {"ms": round(ms, 4), "TFLOP/s": round(tf, 1), "frac_of_3xtf32_roofline": round(tf / peak, 3)}
    hf, hs, h5 = cf.to_host(), cs.to_host(), c5.to_host()
    out["fused_equals_separate_bitwise"] = bool(np.array_equal(hf.view(np.uint32), hs.view(np.uint32)) and
                                                np.array_equal(hf.view(np.uint32), h5.view(np.uint32)))
    if not no_cpu:
        res = cpu_task("config1", {"A": A, "B": B})
        if "error" in res:
            out["reference_cpu"] = res
        else:
            outp = res.pop("out")
            want = np.load(outp)
            out["rel_fro_vs_reference_cpu"] = float(f"{np.linalg.norm(hf.astype(np.float64) - want) / np.linalg.norm(want):.3e}")
            out["reference_cpu"] = res
            _rm_tmp(outp)
    return out


def run_sharded_gemm(jz, L, args, world, rank, stream, timed, pk):
    """BASELINE configs[4]: column-sharded C = log(exp(A*B/n)+1)/5 (3xTF32) with the chain fused in the GEMM epilogue,
    STRONG scaling: the same n^3 product over N GPUs.  N = 1 runs it whole (the anchor).  N > 1 gathers the blocks by
    (a) one NCCL all-gather, (b) P2P stores from the epilogue, (c) multimem.st to the NVSwitch multicast mapping,
    (d) the torch-free jz_mg_* ABI (CUDA IPC).  Parity: replicas identical, all modes the same bits, 64 sampled entries
    per rank against float64, and on rank 0 a 256 x 256 sub-block against the reference's CPU path (oracle/_ref)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from juzhen_b200 import mg
    out = {}
    modes = ("nccl", "fused", "mcast", "ipc") if world > 1 else ("local",)
    for n in args.sharded_n:
        steps = [("affine", 1.0 / n, 0.0), ("exp",), ("affine", 1.0, 1.0), ("log",), ("affine", FIFTH, 0.0)]
        a = jz.CM.randn(n, n, seed=21, offset=0)        # replicated operands: the same bits on every rank (same seed and stream id)
        bfull = jz.CM.randn(n, n, seed=22, offset=0)
        j0, j1 = mg.block_range(n, world, rank)
        b_ptr = bfull.ptr + 4 * j0 * n         # this rank's column block: contiguous in column-major storage
        res = {"strong_scaling": True, "flops": 2.0 * n ** 3}
        sums = {}
        peak = pk["bf16_tflops"] / 2 / 3 * world
        sus = pk["bf16_tflops_sustained"] / 2 / 3 * world   # a launch of tens of milliseconds runs at the power-capped clock
        c_keep = None
        for mode in modes:
            try:
                g = mg.GpuShardedGemm(jz, n, n, n, steps=steps, gemm_mode=0, mode=mode)
                fn = lambda: g.run(a.ptr, n, 0, b_ptr, n, stream)  # noqa: E731
                ms, _ = timed(fn, max(3, args.steps // 3), 2)
            except Exception as e:  # noqa: BLE001 - report, keep the bench line alive
                res[mode] = {"error": f"{type(e).__name__}: {e}"[:300]}
                continue
            tf = 2.0 * n ** 3 / (ms * 1e-3) / 1e12
            chk = g.c_full.view(torch.int32).sum(dtype=torch.int64)
            lo, hi = chk.clone(), chk.clone()
            if world > 1:
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            sums[mode] = int(chk.item())
            res[mode] = {"ms": round(ms, 3), "TFLOP/s": round(tf, 1), "frac_of_3xtf32_peak_xN": round(tf / peak, 3),
                         "frac_of_sustained_3xtf32_peak_xN": round(tf / sus, 3),
                         "replicas_identical": bool(lo.item() == hi.item())}
            if c_keep is None:
                c_keep = g.c_full.clone() if (n <= 16384 or world == 1) else g.c_full   # for the parity checks below
            g.close()
            del g
            torch.cuda.empty_cache()
        ok_modes = [m for m in modes if m in sums]
        if len(ok_modes) > 1:
            res["all_modes_same_bits"] = len({sums[m] for m in ok_modes}) == 1
        if ok_modes:
            best = min(ok_modes, key=lambda m: res[m]["ms"])
            res["best"] = {"mode": best, **res[best]}
        # truth checks at full size on the gathered C of the first mode that ran
        try:
            gen = torch.Generator(device="cpu").manual_seed(1234 + rank)
            ii = torch.randint(0, n, (64,), generator=gen).cuda()
            jj = torch.randint(0, n, (64,), generator=gen).cuda()     # any column: other ranks' blocks included
            A2 = torch.empty(n * n, dtype=torch.float32, device="cuda")
            L.jz_copy(A2.data_ptr(), a.ptr, n * n, stream)
            B2 = torch.empty(n * n, dtype=torch.float32, device="cuda")
            L.jz_copy(B2.data_ptr(), bfull.ptr, n * n, stream)
            Arows = A2.view(n, n).t()[ii, :].double()                    # logical A[i, :] of the column-major buffer
            Bcols = B2.view(n, n)[jj, :].double()                        # logical B[:, j]
            x = (Arows * Bcols).sum(dim=1) / n
            want = torch.log(torch.exp(x) + 1.0) / 5.0
            got = c_keep.view(n, n).t()[ii, jj].double()
            rel = float((got - want).norm() / want.norm())
            worst = torch.tensor([rel], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            res["rel_err_64_sampled_entries_per_rank_vs_float64"] = float(f"{worst.item():.3e}")
            if rank == 0 and not args.no_cpu:
                # 256 x 256 sub-block straddling a block boundary when there is one, against the reference's CPU path
                blk = 256
                r0 = (n // 3) // 4 * 4
                c0 = max(0, min(n - blk, (n // world if world > 1 else n // 2) - blk // 2))
                Ah = A2.view(n, n).t()[r0:r0 + blk, :].contiguous().cpu().numpy()          # (blk, n) C-order
                Bh = B2.view(n, n)[c0:c0 + blk, :].contiguous().cpu().numpy()              # (blk, n): rows = columns of B
                sub = cpu_task("gemm_chain_block", {"Arows": np.asfortranarray(Ah), "Bcols": np.asfortranarray(Bh.T), "n": np.array([n])})
                if "error" in sub:
                    res["subblock_vs_reference_cpu"] = sub
                else:
                    outp = sub.pop("out")
                    want_blk = np.load(outp)
                    _rm_tmp(outp)
                    got_blk = c_keep.view(n, n).t()[r0:r0 + blk, c0:c0 + blk].cpu().numpy()
                    sub["rel_fro"] = float(f"{np.linalg.norm(got_blk.astype(np.float64) - want_blk) / np.linalg.norm(want_blk):.3e}")
                    sub["block"] = f"rows [{r0}, {r0 + blk}) x columns [{c0}, {c0 + blk})"
                    res["subblock_vs_reference_cpu"] = sub
            del A2, B2
        except Exception as e:  # noqa: BLE001
            res["parity_check_error"] = f"{type(e).__name__}: {e}"[:200]
        c_keep = None
        torch.cuda.empty_cache()
        out[str(n)] = res
        del a, bfull
    return out


# ---------------------------------------------------------------------------------- config 4 (launch-bound)
def mnist_step_block():
    """BASELINE configs[3]: one demo_mnist training step (784-1024-128-10 MLP, batch 32, Adam) -- the reference's own
    program, unchanged, built against this backend (build/dropin/bin/demo_mnist) and against the reference's CUDA
    sources + cuBLAS (oracle/_ref/cuda/demo_mnist), both run here on the same GPU.  Wall-clock between the
    program's own progress lines (one every 1000 steps); launch-bound, so no roofline fraction."""
    ours = os.path.join(ROOT, "build", "dropin", "bin", "demo_mnist")
    ref = os.path.join(ROOT, "oracle", "_ref", "cuda", "demo_mnist")
    proj = os.path.join(ROOT, "build", "dropin", "project")
    if not os.path.exists(ours):
        return None
    subprocess.run([sys.executable, os.path.join(ROOT, "juzhen_b200", "cpp", "build_dropin.py"), "--extract-datasets"],
                   capture_output=True, timeout=300)

    def per_1000(binary, env_extra):
        env = dict(os.environ, **env_extra)
        try:
            p = subprocess.Popen([binary], cwd=proj, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        except OSError as e:
            return {"error": str(e)}
        stamps, launches = [], None
        t_end = time.time() + 180
        for ln in p.stdout:
            if "Misclassification Rate" in ln:
                stamps.append(time.perf_counter())
            if "jz_stats" in ln and "kernel launches" in ln:
                launches = int(ln.split("ms,")[1].split("kernel")[0])
            if time.time() > t_end:
                p.kill()
                break
        p.wait()
        gaps = sorted(b - a for a, b in zip(stamps, stamps[1:]))
        if not gaps:
            return {"error": "no progress lines"}
        out = {"ms_per_step": round(gaps[len(gaps) // 2], 4), "intervals": len(gaps)}   # median s per 1000 steps = ms per step
        if launches:
            out["launches_per_step"] = round(launches / 10000.0, 1)
        return out

    blk = {"workload": "examples/demo_mnist.cu unchanged: 10000 Adam steps, batch 32, test pass every 1000 steps (inside the interval)",
           "ours": per_1000(ours, {"JZ_STATS": "1"})}
    if os.path.exists(ref):
        blk["reference_cuda_cublas"] = per_1000(ref, {})
        try:
            blk["speedup_vs_reference_cuda"] = round(blk["reference_cuda_cublas"]["ms_per_step"] / blk["ours"]["ms_per_step"], 2)
        except (KeyError, ZeroDivisionError):
            pass
    # the same training step at batch 32 / 8192 / 60000: juzhen_b200/cpp/tests/
    # bench_mnist_step.cu = the loop body of the demo on synthetic MNIST-shaped data, CUDA events, built against this
    # backend and against the reference's CUDA/cuBLAS sources (fp32 math, and NVIDIA_TF32=1)
    ours_b = os.path.join(ROOT, "build", "dropin", "bin", "bench_mnist_step")
    ref_b = os.path.join(ROOT, "oracle", "_ref", "cuda", "bench_mnist_step")

    def one(binary, batch, env_extra):
        if not os.path.exists(binary):
            return None
        env = dict(os.environ, MNIST_BATCH=str(batch), **env_extra)
        try:
            r = subprocess.run([binary], cwd=proj, env=env, capture_output=True, text=True, timeout=240)
        except (OSError, subprocess.TimeoutExpired) as e:
            return {"error": str(e)[:120]}
        res = {}
        for ln in r.stdout.splitlines():
            if ln.startswith("bench_mnist_step"):
                for tok in ln.split()[1:]:
                    k, _, v = tok.partition("=")
                    res[k] = float(v) if k != "batch" and k != "steps" else int(v)
            if "jz_stats" in ln and "kernel launches" in ln and "steps" in res:
                res["launches_per_step"] = round(int(ln.split("ms,")[1].split("kernel")[0]) / res["steps"], 1)
        return res or {"error": (r.stderr or r.stdout)[-200:]}

    batches = {}
    for batch in (32, 8192, 60000):
        row = {"ours_3xtf32": one(ours_b, batch, {"JZ_STATS": "1"}), "ours_tf32": one(ours_b, batch, {"NVIDIA_TF32": "1"}),
               "reference_cuda_fp32": one(ref_b, batch, {}), "reference_cuda_tf32": one(ref_b, batch, {"NVIDIA_TF32": "1"})}
        try:
            row["speedup_fp32_accuracy"] = round(row["reference_cuda_fp32"]["ms_per_step"] / row["ours_3xtf32"]["ms_per_step"], 2)
            row["speedup_tf32"] = round(row["reference_cuda_tf32"]["ms_per_step"] / row["ours_tf32"]["ms_per_step"], 2)
        except (KeyError, TypeError, ZeroDivisionError):
            pass
        batches[str(batch)] = row
    blk["step_by_batch"] = batches
    return blk


# ---------------------------------------------------------------------------------- reference arm (CPU)
def cpu_step_factory(log2n):
    """the same step on the reference's CPU implementation (bounded sample of the batch)"""
    import numpy as np
    import oracle
    if oracle.ref_available():
        O, kind = oracle.ref(), "reference"
    else:
        O, kind = oracle.port(), "port"
    n = 1 << log2n
    rows = cols = 1 << (log2n // 2)
    if log2n % 2:
        cols *= 2
    rng = np.random.default_rng(0)
    X = rng.standard_normal(n, dtype=np.float32)
    P = rng.random(n, dtype=np.float32) + np.float32(0.5)
    Y = rng.standard_normal(n, dtype=np.float32)
    X2 = np.asfortranarray(X.reshape(rows, cols, order="F"))
    Y2 = np.asfortranarray(Y.reshape(rows, cols, order="F"))
    ops = {
        "fill": lambda: O.affine(X, 0.0, 1.0),
        "exp": lambda: O.unary("exp", X), "log": lambda: O.unary("log", P), "tanh": lambda: O.unary("tanh", X),
        "d_tanh": lambda: O.unary("dtanh", X), "square": lambda: O.unary("square", X),
        "affine_inplace": lambda: O.affine(X, 2.0, 1.0), "eleminv": lambda: O.eleminv(P, 1.0),
        "axpby": lambda: O.axpby(X2, 0, Y2, 0, 1.0, -1.0), "hadamard": lambda: O.hadmd(X2, 0, Y2, 0),
        "chain_softplus5": lambda: O.chain_softplus5(X), "sum_dim0": lambda: O.sum(X2, 0, 0),
        "sum_dim1": lambda: O.sum(X2, 0, 1), "max_dim0": lambda: O.reduce("max", X2, 0, 0),
        "softmax_cols": lambda: O.softmax_cols(X2), "transpose": lambda: O.materialize(X2, 1),
    }
    threads = O.blas_threads(0) if kind == "reference" else 1

    def step():
        for name, _ in SWEEP:
            ops[name]()

    return step, n, kind, threads


def run_cpu(log2n, steps, warmup):
    step, n, kind, threads = cpu_step_factory(log2n)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    bytes_per_step = sum(b for _, b in SWEEP) * n
    out = {"value": round(bytes_per_step / dt / 1e9, 4), "unit": "GB/s", "cores": int(threads), "kind": kind,
           "sample": f"one pass of the same {len(SWEEP)}-op sweep on 2^{log2n} elements "
                     f"({dt:.2f} s/step; elementwise loops are single-threaded in the reference, "
                     f"sum() uses OpenBLAS sgemv on {threads} threads)", "ms_per_step": round(dt * 1e3, 1)}
    if kind == "reference" and log2n >= 20:
        # the GEMM half of the metric on the same host cores: the reference's Matrix<float>::dot (cblas_sgemm, OpenBLAS)
        import numpy as np
        import oracle
        O = oracle.ref()
        gn = 4096
        rng = np.random.default_rng(1)
        A = np.asfortranarray(rng.standard_normal((gn, gn), dtype=np.float32))
        B = np.asfortranarray(rng.standard_normal((gn, gn), dtype=np.float32))
        O.gemm(A[:512, :512], 0, B[:512, :512], 0)   # thread pool warm-up
        t0 = time.perf_counter()
        O.gemm(A, 0, B, 0)
        g = time.perf_counter() - t0
        out["gemm_4096"] = {"TFLOP/s": round(2.0 * gn ** 3 / g / 1e12, 3), "ms": round(g * 1e3, 1), "cores": int(threads),
                            "what": "reference Matrix<float>::dot -> cblas_sgemm (OpenBLAS), one 4096^3 product"}
    return out


def _cpu_env():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU bar is 'all host cores' (OpenBLAS threads)"""
    env = dict(os.environ)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "GOTO_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    return env


def cpu_reference_subprocess(log2n, steps, warmup=0):
    """run in a child so the reference's exit-time profiler printouts cannot pollute our JSON line"""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu_child", "--cpu-log2n", str(log2n),
                        "--steps", str(steps), "--warmup", str(warmup)], capture_output=True, text=True, timeout=1500, env=_cpu_env())
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            return json.loads(ln)
    return {"error": (r.stderr or r.stdout)[-400:]}


def _rm_tmp(outp):
    import shutil
    shutil.rmtree(os.path.dirname(outp), ignore_errors=True)


def cpu_task(task, arrays):
    """one-off CPU-reference computation in a child process (oracle/_ref when it travelled, else the C restatement):
    inputs through an .npz, the result through an .npy whose path comes back as res['out']"""
    import tempfile
    import numpy as np
    d = tempfile.mkdtemp(prefix="jz_cpu_")
    inp = os.path.join(d, "in.npz")
    np.savez(inp, **arrays)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--_cpu_child", "--_cpu_task", task, "--_cpu_io", inp],
                       capture_output=True, text=True, timeout=1500, env=_cpu_env())
    for ln in r.stdout.splitlines():
        if ln.startswith("{"):
            return json.loads(ln)
    return {"error": (r.stderr or r.stdout)[-400:]}


def run_cpu_task(task, inp):
    import numpy as np
    import oracle
    O, kind = (oracle.ref(), "reference") if oracle.ref_available() else (oracle.port(), "port")
    z = np.load(inp)
    outp = inp[:-4] + ".out.npy"
    threads = O.blas_threads(0) if kind == "reference" else 1
    t0 = time.perf_counter()
    if task == "config1":
        A, B = np.asfortranarray(z["A"]), np.asfortranarray(z["B"])
        n = A.shape[0]
        if kind == "reference":
            O.gemm(A[:256, :256], 0, B[:256, :256], 0)     # thread-pool warm-up
            t0 = time.perf_counter()
            res = O.config1(A, B, float(n))
        else:
            prod = O.gemm(A, 0, B, 0)
            res = O.chain_softplus5(O.div_scalar(prod.ravel(order="F"), float(n))).reshape(n, n, order="F")
        dt = time.perf_counter() - t0
        np.save(outp, res)
        return {"out": outp, "kind": kind, "cores": int(threads), "ms": round(dt * 1e3, 1), "TFLOP/s": round(2.0 * n ** 3 / dt / 1e12, 3),
                "what": "reference Matrix<float>: log(exp(A*B/4096)+1)/5, cblas_sgemm + four single-threaded scalar passes"}
    if task == "gemm_chain_block":
        Ar, Bc, n = np.asfortranarray(z["Arows"]), np.asfortranarray(z["Bcols"]), int(z["n"][0])
        prod = O.gemm(Ar, 0, Bc, 0)
        res = O.chain_softplus5(O.div_scalar(prod.ravel(order="F"), float(n))).reshape(prod.shape, order="F")
        np.save(outp, res)
        return {"out": outp, "kind": kind, "cores": int(threads), "ms": round((time.perf_counter() - t0) * 1e3, 1)}
    return {"error": f"unknown task {task}"}


def run_reference(args):
    """the reference arm: the reference's own CPU implementation of the SAME workload (one full pass of the sweep over
    2^log2n elements per step; the CPU rate is flat in size, so steps are capped to keep the run within minutes)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu_steps = 1   # one full-size step is ~25 s of CPU work; --steps K / --warmup W are echoed, the sample says what ran
    res = cpu_reference_subprocess(args.log2n, cpu_steps, 0)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rows = cols = 1 << (args.log2n // 2)
    if args.log2n % 2:
        cols *= 2
    if "error" in res:
        print(json.dumps({"impl": "reference", "unavailable": res["error"][:200]}))
        return
    line = {
        "impl": "reference", "metric": "Elementwise/reduce HBM GB/s (sweep) and GEMM TFLOP/s, % of roofline",
        "value": res["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[1]: elementwise + row/col-sum sweep, 2^{args.log2n} fp32 elements "
                               f"({rows}x{cols}) per GPU, {len(SWEEP)} ops/step",
                   "l2": "inputs (1 GiB/operand) larger than L2, no flush", "ops": [n for n, _ in SWEEP],
                   "algorithmic_bytes_per_step": sum(b for _, b in SWEEP) * (1 << args.log2n),
                   "parallelism": f"independent shards x{world}"},
        "cpu_baseline": res,
        "e2e": {"value": res["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=LOG2N_DEFAULT)
    ap.add_argument("--cpu-log2n", type=int, default=26, dest="cpu_log2n")
    ap.add_argument("--gemm-n", type=int, nargs="*", default=[1024, 2048, 4096, 8192, 16384], dest="gemm_n")
    ap.add_argument("--sharded-n", type=int, nargs="*", default=[16384, 32768], dest="sharded_n")
    ap.add_argument("--no-gemm", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mnist", action="store_true", dest="no_mnist")
    ap.add_argument("--_cpu_child", action="store_true")
    ap.add_argument("--_cpu_task", default="")
    ap.add_argument("--_cpu_io", default="")
    args = ap.parse_args()
    if args._cpu_child:
        if args._cpu_task:
            print(json.dumps(run_cpu_task(args._cpu_task, args._cpu_io)))
        else:
            print(json.dumps(run_cpu(args.cpu_log2n, args.steps, args.warmup)))
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

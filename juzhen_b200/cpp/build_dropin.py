"""Build the reference's OWN programs, unchanged, against the B200 backend (drop-in acceptance).

    python juzhen_b200/cpp/build_dropin.py [--only name ...] [--force]

The reference's sources use relative includes (`#include "../cpp/juzhen.hpp"`), so which backend a
program sees is decided by directory layout.  This script lays out a staging tree under
build/dropin/stage/ (git-ignored) made of FILE symlinks:

    stage/cpp/{core,matrix,operators,helper,juzhen,cpulinalg}.hpp -> /root/reference/cpp/...   (untouched)
    stage/cpp/{cumatrix.cuh,memory.hpp}                          -> juzhen_b200/cpp/...      (this repo)
    stage/ml/*, stage/examples/*.cu, stage/tests/*.cu            -> /root/reference/...       (untouched)

and compiles each acceptance program with nvcc for sm_100a, linking juzhen_b200/cpp/{cumatrix,launcher}.cu
and libjz_b200.so.  No reference file is copied into the repository; only the built binaries (git-ignored)
travel to the GPU box.  Without /root/reference (e.g. on the GPU box) the script does nothing.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("JZ_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "build", "dropin")
STAGE, OBJ, BIN, PROJECT = (os.path.join(OUT, d) for d in ("stage", "obj", "bin", "project"))
LIBDIR = os.path.join(OUT, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
PYSITE = "/opt/prime-rl/.venv/lib/python3.12/site-packages"

# (binary name, source relative to the reference, extra flags)
PROGRAMS = [
    ("testbasic", "tests/testbasic.cu", []),
    ("testStackOps", "tests/testStackOps.cu", []),
    ("testEigen", "tests/testEigen.cu", []),
    ("testElementwiseReduceTorchDump", "tests/testElementwiseReduceTorchDump.cu", []),
    ("demo_gemm", "examples/demo_gemm.cu", []),
    ("demo", "examples/demo.cu", []),
    ("helloworld", "examples/helloworld.cu", []),
    ("helloworld_nn", "examples/helloworld_nn.cu", []),
    ("demo_classification", "examples/demo_classification.cu", []),
    ("demo_mnist", "examples/demo_mnist.cu", []),
    ("knn", "examples/knn.cu", []),
    ("pagerank", "examples/pagerank.cu", []),
    # the reference's own operator benchmark: ConvLayer (cuDNN) and TransformerLayer (its own kernels + cuBLAS batched
    # GEMM on global_handle) ride on this backend's Matrix<CUDAfloat> unchanged -- out-of-scope code paths, built to
    # show that they still work; needs the legacy cuBLAS handle in the launcher
    ("benchmarkCoreOps", "tests/benchmarkCoreOps.cu", ["-DCUDNN_AVAILABLE", "-DJZ_LEGACY_CUBLAS_HANDLE", "-lcudnn", "-lcublas"]),
]
# this repository's own C++ tests (same compute() convention), staged next to the reference's tests
OWN_TESTS = [("test_fusion", "test_fusion.cu", []), ("bench_attention", "bench_attention.cu", ["-lcublas"]),
             ("bench_overhead", "bench_overhead.cu", []), ("bench_mnist_step", "bench_mnist_step.cu", [])]
# own main(): linked without launcher.o
OWN_MAIN = [("test_mg_dot", "test_mg_dot.cu", [])]
# test infrastructure as a shared library (own entry points, no main): the flat C wrapper over the C++ shell that
# tests/test_shell_gpu.py drives through ctypes
OWN_LIBS = [("libjz_shell_capi.so", "shell_capi.cu", [])]
OURS_IN_CPP = {"cumatrix.cuh", "memory.hpp", "jz_lazy.hpp", "jz_mg.hpp"}
REF_CPP = ["core.hpp", "matrix.hpp", "operators.hpp", "helper.hpp", "juzhen.hpp", "cpulinalg.hpp"]


def openblas():
    c = sorted(glob.glob(os.path.join(PYSITE, "opencv_python_headless.libs", "libopenblasp-r0-*.so")))
    if not c:
        raise SystemExit("no Linux OpenBLAS found for the in-binary Matrix<float> CPU path")
    return c[0]


def link(src, dst):
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    if os.path.islink(dst) or os.path.exists(dst):
        os.remove(dst)
    os.symlink(src, dst)


def stage():
    for f in REF_CPP:
        link(os.path.join(REF, "cpp", f), os.path.join(STAGE, "cpp", f))
    for f in OURS_IN_CPP | {"cumatrix.cu", "launcher.cu"}:
        link(os.path.join(HERE, f), os.path.join(STAGE, "cpp", f))
    for sub in ("ml", "examples", "tests"):
        for p in glob.glob(os.path.join(REF, sub, "*")):
            if os.path.isfile(p):
                link(p, os.path.join(STAGE, sub, os.path.basename(p)))
    for _, src, _x in OWN_TESTS + OWN_MAIN + OWN_LIBS:
        link(os.path.join(HERE, "tests", src), os.path.join(STAGE, "tests", src))
    link(os.path.join(REF, "external", "xpu_info", "xpu_info.hpp"), os.path.join(STAGE, "external", "xpu_info", "xpu_info.hpp"))
    eig = os.path.join(REF, "external", "Eigen3")
    if os.path.isdir(eig):  # header tree used only by tests/testEigen.cu: a directory link is fine (no `..` includes)
        link(eig, os.path.join(STAGE, "external", "Eigen3"))
    # writable PROJECT_DIR with the fixtures the programs read (data, not source)
    os.makedirs(os.path.join(PROJECT, "res"), exist_ok=True)
    os.makedirs(os.path.join(PROJECT, "tests"), exist_ok=True)
    shutil.copyfile(os.path.join(REF, "tests", "basic.testdata"), os.path.join(PROJECT, "tests", "basic.testdata"))
    mn = os.path.join(PROJECT, "datasets", "MNIST")
    os.makedirs(mn, exist_ok=True)
    z = os.path.join(REF, "datasets", "MNIST", "dataset.zip")
    if os.path.exists(z) and not os.path.exists(os.path.join(mn, "dataset.zip")):
        shutil.copyfile(z, os.path.join(mn, "dataset.zip"))


def extract_datasets():
    """run on the box that executes the binaries: the image has no `unzip` (ml/util.cuh:311-319 shells out to it)"""
    mn = os.path.join(PROJECT, "datasets", "MNIST")
    z = os.path.join(mn, "dataset.zip")
    if os.path.exists(z) and not os.path.exists(os.path.join(mn, "train_x.matrix")):
        with zipfile.ZipFile(z) as f:
            f.extractall(mn)


def flags():
    ob = openblas()
    return [
        "-std=c++20", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--extended-lambda",
        "-ccbin", "/usr/bin/g++", "-DCUDA", "-DLOGGING_OFF", f'-DPROJECT_DIR="/root/repo/build/dropin/project"',
        "-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "external", "OpenBLAS", "include"),
        "-I", os.path.join(STAGE, "external"), "-Xcudafe", "--diag_suppress=unsigned_compare_with_zero", "-w",
    ], ob


def newer(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(target) > os.path.getmtime(d) for d in deps if os.path.exists(d))


def run(cmd, what):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{what} failed:\n{' '.join(cmd)}\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}")
    return r


def build(only=None, force=False):
    if not os.path.isdir(REF):
        print(f"build_dropin: {REF} absent -- keeping prebuilt binaries under {BIN} if any")
        return []
    stage()
    for d in (OBJ, BIN, LIBDIR):
        os.makedirs(d, exist_ok=True)
    fl, ob = flags()
    ours = [os.path.join(HERE, f) for f in ("cumatrix.cuh", "cumatrix.cu", "memory.hpp", "launcher.cu", "jz_lazy.hpp", "jz_mg.hpp")] + \
           [os.path.join(ROOT, "include", "jz_b200.h")]
    objs = []
    for unit in ("cumatrix", "launcher"):
        o = os.path.join(OBJ, unit + ".o")
        if force or not newer(o, ours):
            run([NVCC, *fl, "-c", os.path.join(STAGE, "cpp", unit + ".cu"), "-o", o], f"compile {unit}.cu")
        objs.append(o)
    # launcher variant that creates Matrix<CUDAfloat>::global_handle for code that calls cuBLAS itself
    launcher_cublas = os.path.join(OBJ, "launcher_cublas.o")
    if force or not newer(launcher_cublas, ours):
        run([NVCC, *fl, "-DJZ_LEGACY_CUBLAS_HANDLE", "-c", os.path.join(STAGE, "cpp", "launcher.cu"), "-o", launcher_cublas],
            "compile launcher.cu (legacy cuBLAS handle)")
    libdir = os.path.join(ROOT, "juzhen_b200")
    # RPATH (not RUNPATH) so OpenBLAS' private libgfortran next to it is found transitively
    link_flags = ["-Xlinker", "--disable-new-dtags", "-L", libdir, "-ljz_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../../juzhen_b200",
                  ob, "-Xlinker", "-rpath", "-Xlinker", os.path.dirname(ob), "-lpthread"]

    def one(prog):
        name, src, extra = prog
        if name.endswith(".so"):   # shared library: its own translation unit + a position-independent cumatrix.cu
            lib = os.path.join(LIBDIR, name)
            if not force and newer(lib, ours + [os.path.realpath(os.path.join(STAGE, src))]):
                return name, "up to date"
            run([NVCC, *fl, *extra, "-Xcompiler", "-fPIC", "-shared", os.path.join(STAGE, src), os.path.join(STAGE, "cpp", "cumatrix.cu"),
                 *link_flags, "-o", lib], f"build {name}")
            return name, "built"
        exe = os.path.join(BIN, name)
        if not force and newer(exe, ours + objs + [os.path.realpath(os.path.join(STAGE, src))]):
            return name, "up to date"
        use = [objs[0], launcher_cublas] if "-DJZ_LEGACY_CUBLAS_HANDLE" in extra else objs
        if name in {n for n, _, _ in OWN_MAIN}:
            use = [objs[0]]
        run([NVCC, *fl, *extra, os.path.join(STAGE, src), *use, *link_flags, "-o", exe], f"build {name}")
        return name, "built"

    todo = [p for p in PROGRAMS + [(n, "tests/" + s, x) for n, s, x in OWN_TESTS + OWN_MAIN + OWN_LIBS] if not only or p[0] in only]
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        res = list(ex.map(one, todo))
    for name, st in res:
        print(f"build_dropin: {name}: {st}")
    return [n for n, _ in res]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*")
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--extract-datasets", action="store_true")
    a = ap.parse_args()
    if a.extract_datasets:
        extract_datasets()
    else:
        build(a.only, a.force)

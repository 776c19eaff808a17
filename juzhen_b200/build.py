"""Build libjz_b200.so (the C-ABI product library) in-tree for sm_100a.

    python -m juzhen_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the
GPU box with the working-tree snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libjz_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v",
]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "jz_b200.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), headers_mtime()):
        return obj, ""
    r = subprocess.run([NVCC, *FLAGS, "-c", srcp, "-o", obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, force), sources()))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "w") as f:
            f.write(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        # extern "C" entry points are exported explicitly via JZ visibility in the link step
        r = subprocess.run([NVCC, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs,
                            "-Xlinker", "--no-undefined"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

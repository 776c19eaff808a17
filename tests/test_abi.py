"""CPU tests: the C-ABI library loads and exports exactly what include/jz_b200.h declares,
and the product path fails loudly (never falls back to a CPU implementation) without a GPU."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "jz_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jz_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import juzhen_b200
    L = juzhen_b200.lib()
    declared = header_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/jz_b200.h but not exported"
    assert sorted(juzhen_b200._lib.exported_symbols()) == declared
    out = subprocess.run(["nm", "-D", "--defined-only", juzhen_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"
    assert L.jz_abi_version() == 1


def test_library_is_sm100a_with_tcgen05_and_tma():
    import juzhen_b200
    sass = subprocess.run(["cuobjdump", "-sass", juzhen_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass, "tcgen05.mma missing from SASS"
    assert "UTMALDG" in sass, "TMA loads missing from SASS"
    assert "LDTM" in sass, "tcgen05.ld missing from SASS"
    for lib in ("cublas", "cudnn", "curand"):
        deps = subprocess.run(["ldd", juzhen_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
        assert lib not in deps, f"product library links {lib}"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import juzhen_b200
    L = juzhen_b200.lib()
    assert L.jz_init(0) == juzhen_b200._lib.JZ_ERR_CUDA
    assert b"no CPU fallback" in L.jz_last_error()
    with pytest.raises(juzhen_b200.JzError):
        juzhen_b200.CM([[1.0, 2.0], [3.0, 4.0]])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "juzhen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "jz_oracle" not in text, f

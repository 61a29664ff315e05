"""The C-ABI library loads without a GPU and exports every symbol include/nbvgpu.h declares."""
import ctypes
import os
import re

import pytest

from nbv_exploration_planner_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "nbvgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nbvgpu_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    N.build()
    lib = ctypes.CDLL(N.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in nbvgpu.h but not exported"
    assert sorted(N.ABI_SYMBOLS) == syms


def test_version_string():
    lib = N.lib()
    assert b"sm_100a" in lib.nbvgpu_version()


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nbv_exploration_planner_b200 import NbvGpu, NbvGpuError
    with pytest.raises(NbvGpuError) as e:
        NbvGpu(0)
    assert e.value.code == N.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/: neither the package nor
    anything under scripts/ does."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "scripts")):
        for f in files:
            txt = open(os.path.join(dirpath, f)).read()
            assert "oracle" not in txt, f
    pkg = os.path.join(ROOT, "nbv_exploration_planner_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "nbv_oracle" not in txt and "nbvo_" not in txt, f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/nbvgpu.h must compile as C99 (no C++ in the signatures) and a C program must link
    against the library; without a device the context constructor reports NBVGPU_ERR_NO_DEVICE."""
    import subprocess
    import torch
    N.build()
    src = tmp_path / "c_abi.c"
    src.write_text('#include <stdio.h>\n#include "nbvgpu.h"\n'
                   "int main(void) {\n"
                   "  nbvgpu_ctx* c = 0;\n"
                   "  int rc = nbvgpu_create(0, &c);\n"
                   '  printf("%s rc=%d\\n", nbvgpu_version(), rc);\n'
                   "  if (rc == NBVGPU_OK) nbvgpu_destroy(c);\n"
                   "  return rc == NBVGPU_OK ? 0 : (rc == NBVGPU_ERR_NO_DEVICE ? 42 : 1);\n"
                   "}\n")
    exe = str(tmp_path / "c_abi")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", exe, "-L", N.CSRC, "-lnbvgpu", f"-Wl,-rpath,{N.CSRC}"])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert "sm_100a" in r.stdout
    assert r.returncode == (0 if torch.cuda.is_available() else 42), r.stdout + r.stderr

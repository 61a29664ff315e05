"""include/nbvgpu_shim.hpp (the reference-signature C++ shims) compiles and links against the C ABI."""
import os
import subprocess

import pytest

from nbv_exploration_planner_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    N.build()
    exe = str(tmp_path / "shim_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "shim_compile_test.cc"), "-o", exe,
                           "-L", N.CSRC, "-lnbvgpu", f"-Wl,-rpath,{N.CSRC}"])
    return exe


def test_shim_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 42 and "no-device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_shim_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("gpu gain="), r.stdout + r.stderr

"""Builds the C++ drop-in test (the reference's README usage against include/b200sort/radixSort.hpp)
and runs it: on CPU it must compile, link and fail loudly (no fallback); on the GPU box it must pass."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "simd-radix-sort_b200"


def _build(tmp_path):
    exe = tmp_path / "dropin_test"
    cmd = ["g++", "-std=c++20", "-O2", "-Wall", "-Wextra", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "cpp" / "dropin_test.cpp"),
           "-o", str(exe), f"-L{LIBDIR}", "-lb200sort", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="no-GPU behaviour")
def test_dropin_header_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_dropin_header_sorts_like_the_reference_readme(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DROPIN OK" in r.stdout, r.stdout + r.stderr

"""The header-only C++ facade (include/hdd_b200.hpp) compiles against the C-ABI and mirrors the reference classes;
on a GPU box the example driver reproduces the ESV2007 goldens through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from dune_hdd_b200 import capi
    capi.lib()
    exe = str(tmp_path / "swipdg_esv2007")
    libdir = os.path.join(ROOT, "dune_hdd_b200")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "swipdg_esv2007.cc"), "-L" + libdir, "-lhdd_b200",
                           "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64",
                           "-o", exe])
    return exe


def test_facade_compiles_and_fails_loudly_without_a_device(tmp_path):
    exe = _build(tmp_path)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([exe, "1"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_reproduces_the_goldens(gpu, tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "MISMATCH" not in r.stdout and "you_are_using_this_wrong: ok" in r.stdout

"""The C++ drop-in adapters (include/orb_slam2/*.h: ORBextractor, ORBmatcher, CeresOptimizer with the reference's
names and call shapes) compile against libcmos_b200.so; without a GPU they fail loudly, with one they run."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "ceres_mono_orb_slam2_b200")


def _build(tmp):
    exe = os.path.join(tmp, "adapter_smoke")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "cpp", "adapter_smoke.cpp"),
                    "-o", exe, "-L" + LIBDIR, "-lcmos_b200", "-Wl,-rpath," + LIBDIR], check=True, capture_output=True)
    return exe


def _have_gpu():
    import torch
    return torch.cuda.is_available()


def test_adapters_compile_and_refuse_to_run_without_gpu(tmp_path):
    exe = _build(str(tmp_path))
    if _have_gpu():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 10 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_adapters_run_on_gpu(tmp_path):
    exe = _build(str(tmp_path))
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ADAPTERS_OK" in r.stdout, r.stdout + r.stderr

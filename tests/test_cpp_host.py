"""The C++ host mirror (include/ndt2d.hpp) compiles against the C ABI with the system g++ (CPU), and the example
SLAM-style driver runs end to end on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "odometry_demo")


def _compile():
    from gtsam_ndt_b200 import build
    assert os.path.exists(build.LIB_CUDA)
    libdir = os.path.dirname(build.LIB_CUDA)
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "odometry_demo.cpp"), "-L", libdir, "-lndt2d", f"-Wl,-rpath,{libdir}", "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_cpp_example_compiles_and_fails_loudly_without_gpu():
    import torch
    _compile()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_example_runs_on_gpu():
    _compile()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "final odometry error" in r.stdout

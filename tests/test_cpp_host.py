"""The C++ host mirror (include/ndt2d.hpp) compiles against the C ABI with the system g++ (CPU), and the example
SLAM-style driver runs end to end on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "odometry_demo")
SLAM = os.path.join(ROOT, "examples", "slam_frontend")


def _compile(exe=EXE):
    from gtsam_ndt_b200 import build
    assert os.path.exists(build.LIB_CUDA)
    libdir = os.path.dirname(build.LIB_CUDA)
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           exe + ".cpp", "-L", libdir, "-lndt2d", f"-Wl,-rpath,{libdir}", "-o", exe]
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


def test_slam_frontend_compiles_and_fails_loudly_without_gpu():
    import torch
    _compile(SLAM)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([SLAM, "/dev/null"], capture_output=True, text=True)
    assert r.returncode == 2 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_slam_frontend_writes_a_pose_graph(tmp_path):
    """configs[4] without GTSAM: GPU odometry + loop-closure factors, exported as a g2o pose graph."""
    _compile(SLAM)
    out = tmp_path / "graph.g2o"
    r = subprocess.run([SLAM, str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = out.read_text().splitlines()
    vertices = [ln.split() for ln in lines if ln.startswith("VERTEX_SE2")]
    edges = [ln.split() for ln in lines if ln.startswith("EDGE_SE2")]
    assert len(vertices) == 97 and len(edges) == 97          # 96 odometry edges + the loop closure 0-96
    assert edges[-1][1:3] == ["0", "96"]
    for e in edges:                                          # information matrices are positive definite
        i = [float(x) for x in e[6:12]]
        c22 = i[0] * i[3] - i[1] * i[1]
        det = i[0] * (i[3] * i[5] - i[4] * i[4]) - i[1] * (i[1] * i[5] - i[4] * i[2]) + i[2] * (i[1] * i[4] - i[3] * i[2])
        assert i[0] > 0 and c22 > 0 and det > 0


def test_gtsam_glue_compiles_against_type_stub_and_rotates_the_hessian(tmp_path):
    """include/ndt2d_gtsam.hpp is compile-gated on GTSAM's headers, which this container does not have. A type stub of the
    four GTSAM names it uses (tests/gtsam_stub: NOT a solver) makes the header parse and lets the check read back the factor:
    the information matrix must be J^T H J, J = blockdiag(R(theta), 1) (it used to be H in the wrong frame)."""
    exe = str(tmp_path / "gtsam_glue_check")
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "gtsam_stub"), os.path.join(ROOT, "tests", "gtsam_glue_check.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "gtsam glue ok" in r.stdout, (r.returncode, r.stdout, r.stderr)

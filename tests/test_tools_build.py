"""The measurement tools that carry a claim in DESIGN.md must keep building (CPU container: nvcc cross-compiles for sm_100a):
tools/mix_probe.py generates and compiles the instruction-mix micro-benchmark with the point loop's own instruction counts,
tools/pipe_probe.cu is the per-instruction-class probe, tools/looplen.py reads the loop length out of the compiled kernel."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")


def test_mix_probe_generates_and_compiles_with_the_loops_instruction_counts():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mix_probe.py"), "build"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(os.path.join(ROOT, "build", "mix_probe"))
    # the interleaved body (variant 0) must hold the packed / FP64 / conversion counts of k_align's loop
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("_Z5k_mixILi0E")][0]
    for token in ("'FFMA2': 30", "'FMUL2': 16", "'FADD2': 5", "'FADD': 8", "'IMAD': 13", "'LDS': 2"):      # (the window between the clock reads also holds a few set-up instructions)
        assert token in line, line


def test_pipe_probe_compiles(tmp_path):
    exe = tmp_path / "pipe_probe"
    r = subprocess.run(["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", str(exe),
                        os.path.join(ROOT, "tools", "pipe_probe.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_looplen_reports_the_point_loop():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "looplen.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    loops = [ln for ln in r.stdout.splitlines() if ln.startswith("loop ") and "'FFMA2': 30" in ln]
    assert loops, r.stdout[-3000:]
    n = int(loops[0].split(":")[1].split()[0])
    assert 120 <= n <= 136, loops[0]      # 128 instructions per 64 points (129 before the shift-add exp scale)

"""Multi-GPU path on real GPUs (needs >= 2 devices on the box; skipped otherwise): the sharded sweep with the
peer-memory best-hypothesis exchange, one process per GPU under torchrun, checked against the unsharded sweep of one GPU
(tools/mgpu_exchange_check.py). The host logic of the same path runs on the CPU in tests/test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_peer_exchange_matches_single_gpu_sweep():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "mgpu_exchange_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=560, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "exchange check" in r.stdout and ", OK," in r.stdout

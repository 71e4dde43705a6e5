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


def test_upload_relay_leaves_the_results_unchanged():
    """ndt2d_set_upload_relay: a share of a host-buffer call's input chunks travels host -> GPU 1 -> GPU 0 (NVLink peer copy);
    the result records must be the same bytes as with every chunk on GPU 0's own link, for f32 and u16 LaserScan input and
    for packed points, and bad arguments must be refused."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    ranges, poses = synth.scans(2048, traj_len=2048, first=0, step=1, **sc)
    init = poses + synth.uniform3(2048) * np.array([0.03, 0.03, np.radians(0.3)])
    ranges, init = np.tile(ranges, (12, 1)), np.tile(init, (12, 1))           # 24 576 scans: six chunks of the pipeline
    map_xy = synth.make_map(512, traj_len=2048, **sc)
    m = g.NdtMatcher2D([0.5], device=0)
    m.set_target(map_xy)
    u16 = np.round(ranges / 0.004).clip(1, 65535).astype(np.uint16)
    pts = synth.polar_to_points(ranges[:6000], sc["angle_min"], sc["angle_inc"])
    xy, off = synth.pack(pts)
    direct = (m.align_batch_ranges(ranges, sc["angle_min"], sc["angle_inc"], init, range_scale=1.0).tobytes(),
              m.align_batch_ranges(u16, sc["angle_min"], sc["angle_inc"], init, range_scale=0.004).tobytes(),
              m.align_batch(xy, off, init[:6000]).tobytes())
    for frac in (0.2, 0.5):
        m.set_upload_relay(1, frac)
        relayed = (m.align_batch_ranges(ranges, sc["angle_min"], sc["angle_inc"], init, range_scale=1.0).tobytes(),
                   m.align_batch_ranges(u16, sc["angle_min"], sc["angle_inc"], init, range_scale=0.004).tobytes(),
                   m.align_batch(xy, off, init[:6000]).tobytes())
        assert relayed == direct
    m.set_upload_relay(-1)
    assert m.align_batch_ranges(ranges, sc["angle_min"], sc["angle_inc"], init, range_scale=1.0).tobytes() == direct[0]
    for dev, frac in ((0, 0.5), (99, 0.5), (1, 0.0), (1, 1.0)):
        with pytest.raises(g.NdtError):
            m.set_upload_relay(dev, frac)
    m.close()

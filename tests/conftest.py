import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def small_world():
    """A seeded scene small enough for the oracle: map cloud, a few scans, true poses."""
    import numpy as np
    from gtsam_ndt_b200 import synth
    traj = 2000
    map_xy = synth.make_map(100, traj_len=traj, sigma=0.01)
    ranges, poses = synth.scans(24, traj_len=traj, first=7, step=83, sigma=0.01, **synth.SCAN_1080)
    scans = synth.polar_to_points(ranges, synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"])
    pert = synth.uniform3(len(scans)) * np.array([0.08, 0.08, np.radians(1.0)])
    return dict(map_xy=map_xy, scans=scans, poses=poses, init=poses + pert, ranges=ranges)

"""CPU counterpart of test_gpu_vs_f64.py: the SPEC ORACLE (bit-identical to the CUDA path, see test_gpu_parity.py)
against the independent double-precision twin in oracle/f64ref.py, at north_star's tolerances. This pins SPEC.md v4's
claim that a f32 per-point algebra on cell-local coordinates stays within 1e-6 of f64 arithmetic (v3 did not: 1.3e-4
on the score, 7e-3 on the Hessian with the same inputs).
PARITY UNPINNED: the twin stands in for the reference's double-precision CPU NDT, which is not in the mount."""
import math

import numpy as np
import pytest

import oracle
from oracle import f64ref


@pytest.fixture(scope="module")
def world():
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    map_xy = synth.make_map(512, traj_len=512, **sc)
    n = 48
    ranges, poses = synth.scans(n, traj_len=65536, first=11, step=1361, **sc)
    scans = synth.polar_to_points(ranges, sc["angle_min"], sc["angle_inc"])
    init = poses + synth.uniform3(n) * np.array([0.03, 0.03, math.radians(0.3)])
    return dict(map_xy=map_xy, scans=scans, init=init)


@pytest.mark.parametrize("overlap", [0, 1])
def test_oracle_evaluation_within_1e6_of_f64_twin(world, overlap):
    o = oracle.Oracle([0.25], overlap=overlap)
    o.set_grid(-100.0, -100.0, 200.0, 200.0)
    o.set_target(world["map_xy"])
    tw = f64ref.NdtF64([o.geometry(0)], overlap=overlap)
    tw.set_target(world["map_xy"])
    sr, hr = [], []
    for s, p0 in list(zip(world["scans"], world["init"]))[:: (3 if overlap else 1)]:
        t = tw.align(s, p0)                      # evaluate where it matters: at a converged pose
        out, cnt = o.evaluate(s, t["pose"])
        assert cnt == t["count"]                 # same cells for every point
        H = np.array([[out[4], out[5], out[6]], [out[5], out[7], out[8]], [out[6], out[8], out[9]]])
        sr.append(abs(out[0] - t["score"]) / t["score"])
        hr.append(np.abs(H - t["hessian"]).max() / np.abs(t["hessian"]).max())
    assert max(sr) <= 1e-6
    assert np.median(hr) <= 2e-7 and max(hr) <= 5e-6
    assert np.mean(np.array(hr) <= 1e-6) >= 0.9


def test_oracle_align_against_f64_twin(world):
    o = oracle.Oracle([0.25])
    o.set_grid(-100.0, -100.0, 200.0, 200.0)
    o.set_target(world["map_xy"])
    tw = f64ref.NdtF64([o.geometry(0)])
    tw.set_target(world["map_xy"])
    dpos, drot = [], []
    for s, p0 in zip(world["scans"], world["init"]):
        r, t = o.align(s, p0), tw.align(s, p0)
        d = r["pose"] - t["pose"]
        dpos.append(math.hypot(d[0], d[1]))
        drot.append(abs((d[2] + math.pi) % (2 * math.pi) - math.pi))
    dpos, drot = np.array(dpos), np.array(drot)
    ok = (dpos <= 1e-5) & (drot <= 1e-6)
    # the median is far inside the tolerance; a few percent of the scans take a different LM path (see test_gpu_vs_f64.py)
    assert np.median(dpos) <= 1e-7 and np.median(drot) <= 1e-8 and ok.mean() >= 0.85

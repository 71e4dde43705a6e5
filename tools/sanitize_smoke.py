#!/usr/bin/env python
"""A small run of every align kernel form for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Block-per-scan, warp-per-scan, warp-per-scan with helper warps, LaserScan input, fused pairs, sweep + top-k."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    traj = 600
    map_xy = synth.make_map(40, traj_len=traj)
    ranges, poses = synth.scans(24, traj_len=traj, first=5, step=23, **sc)
    scans = synth.polar_to_points(ranges, sc["angle_min"], sc["angle_inc"])
    init = poses + synth.uniform3(24) * np.array([0.05, 0.05, np.radians(0.5)])
    xy, off = synth.pack(scans)
    out = {}
    for name, env in (("block", {"NDT2D_BLOCK_ALIGN_MAX": "100000"}), ("warp", {"NDT2D_BLOCK_ALIGN_MAX": "0", "NDT2D_ALIGN_HELP": "0"}),
                      ("help", {"NDT2D_BLOCK_ALIGN_MAX": "0", "NDT2D_ALIGN_HELP": "1"})):
        for k in ("NDT2D_BLOCK_ALIGN_MAX", "NDT2D_ALIGN_HELP"):
            os.environ.pop(k, None)
        os.environ.update(env)
        m = g.NdtMatcher2D([1.0, 0.5], device=0)
        m.set_target(map_xy)
        out[name] = m.align_batch(xy, off, init).tobytes()
        if name == "help":
            out["ranges"] = m.align_batch_ranges(ranges, sc["angle_min"], sc["angle_inc"], init, range_scale=1.0).tobytes()
            hyp = (poses[0] + np.random.default_rng(0).normal(size=(2048, 3)) * [0.3, 0.3, 0.03]).astype(np.float32)
            m.sweep(scans[0], hyp, k=4)
            m.relocalize(scans[0], hyp, k=4)
            pairs = np.array([[i, i + 1] for i in range(8)], np.int32)
            rel = np.zeros((8, 3))
            m.align_pairs(xy, off, pairs, rel)
        m.close()
    assert out["block"] == out["warp"] == out["help"], "kernel forms disagree"
    print("sanitize_smoke ok")


if __name__ == "__main__":
    main()

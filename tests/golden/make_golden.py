#!/usr/bin/env python
"""Regenerates tests/golden/ndt2d_golden.npz from the CPU SPEC ORACLE (run: python tests/golden/make_golden.py).

These are NOT reference vectors: /root/reference holds no source (README.md:1 only), so there is nothing upstream
to generate fixtures from. The file freezes the spec oracle's outputs on a small seeded scene so that (a) an
accidental change to SPEC.md's arithmetic in oracle/ is caught on the CPU, and (b) the CUDA path is checked
against committed numbers on the GPU box as well as against the live oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def scene():
    from gtsam_ndt_b200 import synth
    traj = 300
    map_xy = synth.make_map(30, traj_len=traj)
    ranges, poses = synth.scans(6, traj_len=traj, first=4, step=47, **synth.SCAN_1080)
    scans = synth.polar_to_points(ranges, synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"])
    scans = [s[::3] for s in scans]                       # 360 points each keeps the fixture small
    init = poses + synth.uniform3(6) * np.array([0.06, 0.06, np.radians(0.8)])
    return map_xy, scans, poses, init


def compute(overlap):
    import oracle
    from gtsam_ndt_b200 import synth
    map_xy, scans, poses, init = scene()
    o = oracle.Oracle([1.0, 0.5], overlap=overlap)
    o.set_target(map_xy)
    xy, off = synth.pack(scans)
    res = o.align_batch(xy, off, init, nthreads=1)
    ev = np.array([o.evaluate(scans[i], init[i], level=1)[0] for i in range(len(scans))])
    cnt = np.array([o.evaluate(scans[i], init[i], level=1)[1] for i in range(len(scans))])
    idx = o.cell_index(scans[0], init[0], level=1)
    cells = o.cells(1)
    nz = np.argwhere(cells[..., 7] != 0)
    hyp = (poses[2] + np.stack(np.meshgrid(np.arange(-2, 3) * 0.25, np.arange(-2, 3) * 0.25, np.radians(np.arange(-2, 3) * 2.0),
                                           indexing="ij"), -1).reshape(-1, 3)).astype(np.float32)
    scores, bi, _ = o.sweep(scans[2], hyp, level=0, nthreads=1)
    return dict(res_pose=res["pose"], res_score=res["score"], res_hessian=res["hessian"], res_iter=res["iterations"],
                res_status=res["status"], res_count=res["count"], eval10=ev, eval_count=cnt, cell_index=idx,
                cells_nz_index=nz.astype(np.int32), cells_nz=cells[nz[:, 0], nz[:, 1]], geom=np.array(list(o.geometry(1).values()), np.float64),
                sweep_scores=scores, sweep_best=np.int64(bi))


def main():
    out = {}
    for ov in (0, 1):
        for k, v in compute(ov).items():
            out[f"ov{ov}_{k}"] = v
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ndt2d_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

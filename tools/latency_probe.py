#!/usr/bin/env python
"""Per-evaluation latency of the align kernels as a function of the scan length: 148 scans (one per SM, nothing competes),
convergence switched off (eps = 0) so that every scan runs max_iterations evaluations, scans cut to n points.
Separates the cost of a 64-point step from the serial part of an evaluation (reduction, solver, pose set-up).

  python tools/latency_probe.py            (GPU box; one JSON line per n)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    B, iters = 148, 20
    ranges, poses = synth.scans(B, traj_len=10000, first=0, step=1, **sc)
    cb, sb = synth.beam_table(sc["nbeams"], sc["angle_min"], sc["angle_inc"])
    full = np.stack([ranges * cb[None, :], ranges * sb[None, :]], axis=-1).astype(np.float32)
    map_xy = synth.make_map(2048, traj_len=2048, **sc)
    init = poses + synth.uniform3(B, first=31337) * np.array([0.03, 0.03, 0.005])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    d_init = torch.from_numpy(np.ascontiguousarray(init)).to(dev)
    for n in (64, 256, 512, 768, 1024, 1080):
        xy = np.ascontiguousarray(full[:, :: max(1, 1080 // n)][:, :n]).reshape(-1, 2)
        d_xy = torch.from_numpy(xy).to(dev)
        off = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * n).to(dev)
        row = {"points": n, "steps": (n + 63) // 64}
        for mode in ("warp", "help", "block"):
            os.environ["NDT2D_BLOCK_ALIGN_MAX"] = str(1 << 30) if mode == "block" else "0"
            os.environ["NDT2D_ALIGN_HELP"] = "1" if mode == "help" else "0"
            m = g.NdtMatcher2D([0.25], device=0, stream=stream.cuda_stream, max_iterations=iters, eps_trans=0.0, eps_rot=0.0)
            m.set_grid(-100.0, -100.0, 200.0, 200.0)
            m.set_target(map_xy)
            d_res = torch.zeros(B * 144, dtype=torch.uint8, device=dev)
            for _ in range(5):
                m.align_batch_device(d_xy, off, B, n, d_init, d_res)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                m.align_batch_device(d_xy, off, B, n, d_init, d_res)
            e1.record(stream)
            torch.cuda.synchronize()
            r = np.frombuffer(d_res.cpu().numpy().tobytes(), dtype=g.RESULT_DTYPE)
            ms = e0.elapsed_time(e1) / 20
            row[mode + "_us_per_eval"] = round(ms * 1e3 / r["iterations"].max(), 3)
            row["max_iter"] = int(r["iterations"].max())
            row["min_iter"] = int(r["iterations"].min())
            m.close()
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Mid-size batches: where does the block-per-scan align kernel (k_align_block, eight warps per scan) beat the warp-per-scan
kernel (k_align)? A batch of B scans keeps B warps busy in k_align; below ~3500 scans (148 SMs x 24 warps) the GPU is not
full, and the step ends with the slowest scan (30-90 evaluations on a pyramid, ~9 us each on one warp).

  python tools/midsize_probe.py [--sizes 300 600 ...] [--steps 20]        (GPU box)

Prints one JSON line per (configuration, batch size): ms per batch with each kernel forced through NDT2D_BLOCK_ALIGN_MAX / NDT2D_ALIGN_HELP,
and whether the result records are identical (they must be).
"""
import argparse
import hashlib
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[148, 296, 600, 1250, 2500, 5000, 10000, 20000])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--modes", nargs="+", default=["warp", "help", "block", "auto"])
    a = ap.parse_args()
    import numpy as np
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    nmax = max(a.sizes)
    ranges, poses = synth.scans(nmax, traj_len=10000, first=0, step=1, **sc)
    cb, sb = synth.beam_table(sc["nbeams"], sc["angle_min"], sc["angle_inc"])
    xy = np.stack([ranges * cb[None, :], ranges * sb[None, :]], axis=-1).astype(np.float32).reshape(-1, 2)
    map_xy = synth.make_map(2048, traj_len=2048, **sc)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    d_xy = torch.from_numpy(xy).to(dev)
    d_xy2 = d_xy.clone()
    configs = [("pyramid 2.0/1.0/0.5, prior 0.2 m / 3 deg", [2.0, 1.0, 0.5], (0.2, 3.0)),
               ("single level 0.25 m, prior 0.03 m / 0.3 deg", [0.25], (0.03, 0.3))]
    for label, res, pert in configs:
        init = poses + synth.uniform3(nmax, first=31337) * np.array([pert[0], pert[0], math.radians(pert[1])])
        d_init = torch.from_numpy(np.ascontiguousarray(init)).to(dev)
        for B in a.sizes:
            off = torch.from_numpy(np.arange(B + 1, dtype=np.int64) * 1080).to(dev)
            row = {"config": label, "scans": B}
            for mode in a.modes:
                # the knobs are read when a handle is created. warp: one warp per scan; help: the same with helper warps;
                # block: one block per scan; auto: the library's own choice
                for k in ("NDT2D_BLOCK_ALIGN_MAX", "NDT2D_ALIGN_HELP"):
                    os.environ.pop(k, None)
                if mode != "auto":
                    os.environ["NDT2D_BLOCK_ALIGN_MAX"] = str(1 << 30) if mode == "block" else "0"
                    os.environ["NDT2D_ALIGN_HELP"] = "1" if mode == "help" else "0"
                m = g.NdtMatcher2D(res, device=0, stream=stream.cuda_stream)
                m.set_grid(-100.0, -100.0, 200.0, 200.0)
                m.set_target(map_xy)
                d_res = torch.zeros(B * 144, dtype=torch.uint8, device=dev)
                for i in range(5):
                    m.align_batch_device(d_xy if i & 1 else d_xy2, off, B, 1080, d_init, d_res)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for i in range(a.steps):
                    m.align_batch_device(d_xy if i & 1 else d_xy2, off, B, 1080, d_init, d_res)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.steps
                raw = d_res.cpu().numpy().tobytes()
                r = np.frombuffer(raw, dtype=g.RESULT_DTYPE)
                row[mode + "_ms"] = round(ms, 4)
                row[mode + "_Mmatches_s"] = round(B / ms / 1e3, 3)
                row[mode + "_sha"] = hashlib.sha1(raw).hexdigest()[:10]
                row["mean_iter"] = round(float(r["iterations"].mean()), 2)
                row["max_iter"] = int(r["iterations"].max())
                m.close()
            shas = {row[k] for k in row if k.endswith("_sha")}
            row["identical"] = len(shas) == 1
            print(json.dumps(row), flush=True)
    for k in ("NDT2D_BLOCK_ALIGN_MAX", "NDT2D_ALIGN_HELP"):
        os.environ.pop(k, None)


if __name__ == "__main__":
    main()

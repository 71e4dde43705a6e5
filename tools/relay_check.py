#!/usr/bin/env python
"""Upload relay (ndt2d_set_upload_relay) on a box with at least two GPUs: one handle on GPU 0, its host-buffer batch call
with the input going (a) over GPU 0's own PCIe link only, (b) partly over GPU `--relay`'s link and NVLink. Prints matches/s
and the effective host-to-device rate for each setting and checks that the result records are the same bytes.

  python tools/relay_check.py [--scans 65536] [--relay 1] [--fractions 0.25 0.4 0.5]
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=65536)
    ap.add_argument("--relay", type=int, default=1)
    ap.add_argument("--fractions", type=float, nargs="+", default=[0.25, 0.4, 0.5])
    ap.add_argument("--u16", type=int, default=0)
    a = ap.parse_args()
    import numpy as np
    import torch
    import gtsam_ndt_b200 as g
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    base = min(a.scans, 8192)
    ranges, poses = synth.scans(base, traj_len=base, first=0, step=1, **sc)
    init = poses + synth.uniform3(base) * np.array([0.03, 0.03, math.radians(0.3)])
    reps = a.scans // base
    ranges, init = np.tile(ranges, (reps, 1)), np.tile(init, (reps, 1))
    if a.u16:
        ranges = np.round(ranges * 1000.0).astype(np.uint16)
    map_xy = synth.make_map(2048, traj_len=2048, **sc)
    B = len(ranges)
    m = g.NdtMatcher2D([0.25], device=0)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    m.set_target(map_xy)
    h_in = torch.from_numpy(ranges).pin_memory()
    h_init = torch.from_numpy(np.ascontiguousarray(init)).pin_memory()
    h_res = torch.zeros(B * 144, dtype=torch.uint8).pin_memory()
    rv = h_res.numpy().view(g.RESULT_DTYPE)
    scale = 0.001 if a.u16 else 1.0
    fn = lambda: m.align_batch_ranges(h_in.numpy(), sc["angle_min"], sc["angle_inc"], h_init.numpy(), range_scale=scale, out=rv)
    ref = None
    for frac in [0.0] + list(a.fractions):
        m.set_upload_relay(a.relay if frac > 0 else -1, frac if frac > 0 else 0.5)
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        ms = (time.perf_counter() - t0) * 1e3 / 10
        raw = rv.tobytes()
        if ref is None:
            ref = raw
        print(json.dumps({"relay_fraction": frac, "relay_device": a.relay if frac > 0 else None, "scans": B, "ms": round(ms, 3),
                          "Mmatches_s": round(B / ms / 1e3, 2), "effective_h2d_gbs": round(ranges.nbytes / ms / 1e6, 1),
                          "identical_to_direct": raw == ref}), flush=True)
    m.close()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Kernel tuning experiments on configs[1] (scan-to-map, 1080-beam scans, 0.25 m cells).

  python tools/exp.py build name=-DFLAG,-DFLAG2 ...      (here, no GPU: nvcc -> build/variants/)
  python tools/exp.py gen [--scans N]                     (GPU box: synthetic workload -> /tmp/ndt2d_exp.npz)
  python tools/exp.py run name [--scans N] [--steps K]    (GPU box: one JSON line; `base` = the in-tree library)

Every run prints a hash of the result records, so variants can be checked to be bit-identical to each other.
"""
import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NPZ = "/tmp/ndt2d_exp.npz"


def variant_path(name):
    return os.path.join(ROOT, "build", "variants", f"libndt2d_{name}.so")


def cmd_build(specs):
    from gtsam_ndt_b200 import build
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = build.build_variant(name, [f for f in flags.split(",") if f])
        log = open(out + ".log").read()
        regs = [l for l in log.splitlines() if "Used" in l]
        print(name, flags, "->", os.path.relpath(out, ROOT))


def cmd_gen(scans, overlap=0):
    import numpy as np
    from gtsam_ndt_b200 import synth
    sc = synth.SCAN_1080
    t0 = time.time()
    ranges, poses = synth.scans(scans, traj_len=scans, first=0, step=1, **sc)
    pert = synth.uniform3(scans) * np.array([0.03, 0.03, math.radians(0.3)])
    init = poses + pert
    map_xy = synth.make_map(2048, traj_len=2048, **sc)
    cb, sb = synth.beam_table(sc["nbeams"], sc["angle_min"], sc["angle_inc"])
    xy = np.stack([ranges * cb[None, :], ranges * sb[None, :]], axis=-1).astype(np.float32).reshape(-1, 2)
    off = np.arange(scans + 1, dtype=np.int64) * 1080
    np.savez(NPZ, xy=xy, off=off, init=init, map_xy=map_xy)
    print(f"gen: {scans} scans in {time.time() - t0:.1f} s -> {NPZ}", file=sys.stderr)


def cmd_run(name, scans, steps, warmup, res, overlap, shuffle):
    if name != "base":
        os.environ["NDT2D_LIB"] = variant_path(name)
    import numpy as np
    import torch
    import gtsam_ndt_b200 as g
    if not os.path.exists(NPZ):
        cmd_gen(scans)
    z = np.load(NPZ)
    B = min(scans, len(z["off"]) - 1)
    xy, off, init = z["xy"][: B * 1080], z["off"][: B + 1], z["init"][:B]
    if shuffle:
        perm = np.random.default_rng(1).permutation(B)
        xy = np.ascontiguousarray(xy.reshape(B, 1080, 2)[perm].reshape(-1, 2))
        init = np.ascontiguousarray(init[perm])
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D(res, device=0, stream=stream.cuda_stream, overlap=overlap)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    m.set_target(z["map_xy"])
    d_xy = torch.from_numpy(xy).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_init = torch.from_numpy(np.ascontiguousarray(init)).to(dev)
    d_res = torch.zeros(B * 144, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        m.align_batch_device(d_xy, d_off, B, 1080, d_init, d_res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        m.align_batch_device(d_xy, d_off, B, 1080, d_init, d_res)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    raw = d_res.cpu().numpy().tobytes()
    r = np.frombuffer(raw, dtype=g.RESULT_DTYPE)
    if shuffle:
        inv = np.argsort(perm)
        raw = r[inv].tobytes()
    print(json.dumps({"name": name, "scans": B, "ms": round(ms, 4), "Mmatches_s": round(B / ms / 1e3, 3),
                      "Mevals_s": round(float(r["iterations"].sum()) / ms / 1e3, 2), "mean_iter": round(float(r["iterations"].mean()), 3),
                      "converged": int((r["status"] == 0).sum()), "sha": hashlib.sha1(raw).hexdigest()[:12],
                      "shuffle": shuffle, "overlap": overlap}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["build", "gen", "run"])
    ap.add_argument("names", nargs="*")
    ap.add_argument("--scans", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--res", type=float, nargs="+", default=[0.25])
    ap.add_argument("--overlap", type=int, default=0)
    ap.add_argument("--shuffle", type=int, default=0, help="1: random scan order in the batch (no locality between neighbouring jobs)")
    a = ap.parse_args()
    if a.cmd == "build":
        cmd_build(a.names)
    elif a.cmd == "gen":
        cmd_gen(a.scans)
    else:
        for n in a.names:
            if len(a.names) > 1:
                subprocess.run([sys.executable, __file__, "run", n, "--scans", str(a.scans), "--steps", str(a.steps), "--warmup", str(a.warmup),
                                "--res"] + [str(r) for r in a.res] + ["--overlap", str(a.overlap), "--shuffle", str(a.shuffle)])
            else:
                cmd_run(n, a.scans, a.steps, a.warmup, a.res, a.overlap, a.shuffle)

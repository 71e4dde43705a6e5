#!/usr/bin/env python
"""Multi-GPU check of the peer-memory best-hypothesis exchange (run under torchrun, one rank per GPU):
the sharded sweep published through ndt2d_sweep_publish / ndt2d_exchange_wait must return, on every rank and for
every query, exactly what one GPU returns for the unsharded sweep (SPEC.md section 6: ties to the smaller index), and
distributed.relocalize_sharded must equal the single-GPU ndt2d_relocalize bit for bit.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_exchange_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gtsam_ndt_b200 as g                      # noqa: E402
from gtsam_ndt_b200 import distributed as D, synth  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sc = synth.SCAN_1080
    map_xy = synth.make_map(256, traj_len=256, **sc)
    ranges, poses = synth.scans(1, traj_len=1000, first=137, **sc)
    xy = synth.polar_to_points(ranges[0], sc["angle_min"], sc["angle_inc"])
    gx = (np.arange(41) - 20) * 0.2
    lat = np.stack(np.meshgrid(gx, gx, np.radians(np.arange(60) * 6.0 - 180.0), indexing="ij"), -1).reshape(-1, 3)
    hyp = (poses[0] + lat).astype(np.float32)
    hyp = np.concatenate([hyp, hyp[:997]])          # exact ties across shards: the smaller global index must win
    stream = torch.cuda.current_stream()
    m = g.NdtMatcher2D([0.25], device=local, stream=stream.cuda_stream)
    m.set_grid(-100.0, -100.0, 200.0, 200.0)
    m.set_target(map_xy)
    ex = D.PeerExchange(m, nslots=16)           # batches of 8 in flight: nslots >= 2 x batch (see ndt2d.h)
    d_xy = torch.from_numpy(xy).to(dev)
    nq = 24
    ok = True
    for rep in range(2):                            # second pass reuses the slots (epochs keep growing)
        # query q scores the hypotheses rolled by q, sharded contiguously
        shards = []
        for q in range(nq):
            h = np.roll(hyp, 37 * q, axis=0)
            lo, hi = D.shard_range(len(h), rank, world)
            shards.append((h, lo, hi, torch.from_numpy(h[lo:hi].copy()).to(dev)))
        d_scores = torch.zeros(max(s[2] - s[1] for s in shards), dtype=torch.float64, device=dev)
        for base in range(0, nq, 8):                # a rank is never more than one batch ahead of the slowest rank
            for q in range(base, base + 8):
                h, lo, hi, d_h = shards[q]
                ex.publish(d_xy, len(xy), d_h, hi - lo, d_scores, lo, rep * nq + q)
            for q in range(base, base + 8):
                bi, bs = ex.wait(rep * nq + q, timeout_ms=20000)
                h = shards[q][0]
                _, ri, rs = m.sweep(xy, h, k=1, want_scores=False)     # unsharded, on this GPU
                if bi != int(ri[0]) or bs != float(rs[0]):
                    ok = False
                    print(f"rank {rank} query {q}: exchange ({bi}, {bs}) != single GPU ({int(ri[0])}, {float(rs[0])})", flush=True)
    # timing: queries back to back, one wait per batch of 8
    torch.cuda.synchronize(); dist.barrier()
    h, lo, hi, d_h = shards[0]
    t0 = time.perf_counter()
    for it in range(10):
        for q in range(8):
            ex.publish(d_xy, len(xy), d_h, hi - lo, d_scores, lo, 1000 + it * 8 + q)
        ex.wait(1000 + it * 8 + 7, timeout_ms=20000)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 80
    # multi-GPU relocalisation end to end: sharded sweep -> global top-k -> refinement dealt out over the ranks; every rank
    # must hold, bit for bit, what ndt2d_relocalize returns on one GPU
    for k in (1, 3, 8):
        gi, gr = D.relocalize_sharded(m, xy, hyp, k=k)
        si, sr = m.relocalize(xy, hyp, k=k)
        if not (np.array_equal(gi, si) and gr.tobytes() == sr.tobytes()):
            ok = False
            print(f"rank {rank}: relocalize_sharded(k={k}) differs from the single-GPU ndt2d_relocalize", flush=True)
    # ... and the same through the peer-memory candidate exchange (one exchange per query, no collective), slots reused
    pr = D.PeerRelocalizer(m, nslots=4, kmax=8)
    lo, hi = D.shard_range(len(hyp), rank, world)
    d_shard = torch.from_numpy(hyp[lo:hi].copy()).to(dev)
    for q, k in enumerate((1, 3, 8, 8, 5, 8)):
        pr.publish(d_xy, len(xy), d_shard, hi - lo, lo, k, q)
        gi, gr = pr.wait(q, k, timeout_ms=20000)
        si, sr = m.relocalize(xy, hyp, k=k)
        if not (np.array_equal(gi, si) and gr.tobytes() == sr.tobytes()):
            ok = False
            print(f"rank {rank}: peer relocalisation (k={k}, query {q}) differs from the single-GPU ndt2d_relocalize", flush=True)
        dist.barrier()      # nslots = 4: nobody runs more than a query ahead here
    pr.close()
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ex.close()
    if rank == 0:
        print(f"exchange check: world {world}, {2 * nq} queries x {len(hyp)} hypotheses, "
              f"{'OK' if flag.item() == 1.0 else 'MISMATCH'}, {dt * 1e3:.3f} ms per query incl. publication", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()

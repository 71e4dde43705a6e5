"""Host-side multi-GPU logic on the CPU: world size 2 over gloo (no GPU needed). The per-shard scorer here is the
CPU spec oracle, allowed in tests only; on the GPU box the same functions wrap NdtMatcher2D.sweep."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gtsam_ndt_b200 import distributed as D, synth
    import oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        traj = 400
        map_xy = synth.make_map(40, traj_len=traj)
        r, p = synth.scans(1, traj_len=traj, first=33, **synth.SCAN_1080)
        xy = synth.polar_to_points(r[0], synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"])
        g = np.stack(np.meshgrid(np.arange(-4, 5) * 0.3, np.arange(-4, 5) * 0.3, np.radians(np.arange(-3, 4) * 3.0), indexing="ij"), -1).reshape(-1, 3)
        hyp = (p[0] + g).astype(np.float32)
        hyp = np.concatenate([hyp, hyp[100:140]])          # duplicated hypotheses straddle the shard boundary
        o = oracle.Oracle([1.0])
        o.set_target(map_xy)

        def score_shard(lo, hi, k):
            s, _, _ = o.sweep(xy, hyp[lo:hi], nthreads=1)
            order = np.lexsort((np.arange(len(s)), -s))[:k]
            idx = np.full(k, -1, np.int64); val = np.zeros(k)
            idx[: len(order)] = order; val[: len(order)] = s[order]
            return idx, val

        gi, gs = D.sweep_sharded(score_shard, len(hyp), k=5)
        full, _, _ = o.sweep(xy, hyp, nthreads=1)
        exp = np.lexsort((np.arange(len(full)), -full))[:5]
        ok = np.array_equal(gi, exp) and np.array_equal(gs, full[exp])
        # multi-GPU relocalisation (sharded sweep -> global top-k -> refinement dealt out over the ranks -> gathered records)
        # with the oracle standing in for this rank's GPU matcher: every rank must hold what one matcher computes alone
        class OracleMatcher:
            def sweep(self, xy_, hyp_, k=1, level=0, want_scores=True):
                return (None,) + score_shard_of(hyp_, k)

            def align_batch(self, xy_, off_, init_):
                return o.align_batch(xy_, off_, init_, nthreads=1)

        def score_shard_of(h, k):
            s, _, _ = o.sweep(xy, h, nthreads=1)
            order = np.lexsort((np.arange(len(s)), -s))[:k]
            idx = np.full(k, -1, np.int64); val = np.zeros(k)
            idx[: len(order)] = order; val[: len(order)] = s[order]
            return idx, val

        for kk in (1, 3, 5):
            ri, rr = D.relocalize_sharded(OracleMatcher(), xy, hyp, k=kk)
            ei, _ = score_shard_of(hyp, kk)
            er = o.align_batch(np.tile(xy, (kk, 1)), np.arange(kk + 1, dtype=np.int64) * len(xy), hyp[ei].astype(np.float64), nthreads=1)
            ok = ok and np.array_equal(ri, ei) and rr.tobytes() == er.tobytes()
        lo, hi = D.align_sharded_counts(101)
        # the plumbing of the peer-memory exchange: one 64-byte blob per rank (a CUDA IPC handle on the GPU box), by rank
        blobs = D.all_gather_blobs(bytes([rank + 1]) * 64)
        ok = ok and blobs == [bytes([r + 1]) * 64 for r in range(world)]
        q.put((rank, bool(ok), (lo, hi), gi.tolist()))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions():
    from gtsam_ndt_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_combine_topk_single_process():
    from gtsam_ndt_b200.distributed import combine_topk
    i, s = combine_topk([5, 2, -1], [1.0, 1.0, 0.0], 3)
    assert i.tolist() == [2, 5, -1] and s.tolist() == [1.0, 1.0, 0.0]


@pytest.mark.timeout(300)
def test_sweep_sharded_world2_gloo_matches_unsharded():
    world, port = 2, 29400 + os.getpid() % 500
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    out.sort()
    assert all(o[1] for o in out), out
    assert out[0][3] == out[1][3]                      # every rank holds the same global answer
    assert out[0][2] == (0, 51) and out[1][2] == (51, 101)


def test_upload_relay_pairs_and_refinement():
    """Host logic of the upload relay plan (gtsam_ndt_b200.distributed): rates as measured on the 8-GPU box."""
    from gtsam_ndt_b200 import distributed as D
    pairs = D.upload_relay_pairs([23.4, 23.5, 23.4, 23.5, 34.9, 35.3, 35.2, 35.3])
    assert sorted(pairs) == [0, 1, 2, 3] and sorted(f for f, _ in pairs.values()) == [4, 5, 6, 7]
    assert all(0.19 < x < 0.21 for _, x in pairs.values())
    assert D.upload_relay_pairs([55.3, 55.4]) == {} and D.upload_relay_pairs([50.0]) == {}
    assert D.upload_relay_pairs([20.0, 30.0, 40.0]) == {0: (2, (40.0 - 20.0) / 60.0)}       # odd world: the middle rank stays alone
    x2 = D.refine_relay_fraction(0.2, 11.6, 10.0)
    assert 0.26 < x2 < 0.28
    # at the refined share the modelled step times meet
    assert abs(11.6 / 0.8 * (1 - x2) - 10.0 / 1.2 * (1 + x2)) < 1e-9
    assert D.refine_relay_fraction(0.5, 30.0, 1.0) == 0.6 and D.refine_relay_fraction(0.1, 1.0, 30.0) == 0.05

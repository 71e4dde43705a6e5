"""Multi-GPU host logic for the batched paths: one process per GPU, torch.distributed
for the plumbing. Scans and hypotheses are independent, so they are sharded by contiguous ranges with no
data-path collective; the sweep's only exchange is the best-hypothesis combine (16 B per rank and per k).

Nothing here computes NDT: scoring is done by the caller's matcher (libndt2d.so on that rank's GPU).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, near-equal [lo, hi) slice of n items for `rank` of `world` (first n % world ranks get one more)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device_for_backend():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def combine_topk(local_idx, local_score, k):
    """Merge per-rank top-k lists (global indices, scores) into the global top-k ordered by (-score, index).

    Every rank passes arrays of length k (pad with index -1); every rank gets the same answer.
    Ties go to the smaller global index, so the result equals the unsharded sweep's (SPEC.md section 6)."""
    local_idx = np.asarray(local_idx, np.int64).reshape(-1)
    local_score = np.asarray(local_score, np.float64).reshape(-1)
    assert len(local_idx) == k and len(local_score) == k
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        all_idx, all_score = local_idx, local_score
    else:
        dev = _device_for_backend()
        world = dist.get_world_size()
        ti = torch.from_numpy(local_idx).to(dev)
        ts = torch.from_numpy(local_score).to(dev)
        gi = [torch.empty_like(ti) for _ in range(world)]
        gs = [torch.empty_like(ts) for _ in range(world)]
        dist.all_gather(gi, ti)
        dist.all_gather(gs, ts)
        all_idx = torch.cat(gi).cpu().numpy()
        all_score = torch.cat(gs).cpu().numpy()
    keep = all_idx >= 0
    all_idx, all_score = all_idx[keep], all_score[keep]
    order = np.lexsort((all_idx, -all_score))[:k]
    out_i = np.full(k, -1, np.int64)
    out_s = np.zeros(k, np.float64)
    out_i[: len(order)] = all_idx[order]
    out_s[: len(order)] = all_score[order]
    return out_i, out_s


def sweep_sharded(score_shard, nhyp, k=1):
    """Relocalisation sweep over `nhyp` hypotheses sharded across the ranks of the default process group.

    score_shard(lo, hi, k) -> (idx[k] relative to lo or -1, score[k]) scores hypotheses [lo, hi) on this rank
    (e.g. lambda lo, hi, k: matcher.sweep(xy, hyp[lo:hi], k, want_scores=False)[1:]).
    Returns the global (idx[k], score[k])."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(nhyp, rank, world)
    idx, score = score_shard(lo, hi, k)
    idx = np.asarray(idx, np.int64).copy()
    idx[idx >= 0] += lo
    return combine_topk(idx, score, k)


def all_gather_blobs(blob):
    """All-gather one fixed-size bytes object per rank (e.g. a 64-byte CUDA IPC handle); returns the list by rank.
    Works on both back ends: the bytes travel as a uint8 tensor on the back end's device."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [bytes(blob)]
    dev = _device_for_backend()
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in out]


def all_gather_array(a):
    """All-gather one equally shaped numpy array per rank (any dtype, sent as bytes); returns the list by rank."""
    a = np.ascontiguousarray(a)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [a.copy()]
    blobs = all_gather_blobs(a.tobytes()) if a.nbytes else [b""] * dist.get_world_size()
    return [np.frombuffer(b, dtype=a.dtype).reshape(a.shape).copy() for b in blobs]


def relocalize_sharded(matcher, xy, hyp, k=4, level=0):
    """Multi-GPU relocalisation end to end: every rank sweeps its contiguous shard of `hyp` (the same array on every
    rank) and keeps its top-k, the per-rank lists are combined into the global top-k (combine_topk: 16 B x k per rank),
    the k candidates are dealt out over the ranks again, every rank refines its share with the full align
    (ndt2d_align_batch) and the result records are all-gathered. Every rank returns (idx[k], res[k]) ordered like the
    top-k - bit for bit what ndt2d_relocalize returns on one GPU (SPEC 6: ties to the smaller global index; an align does
    not depend on the batch it runs in)."""
    from .matcher import RESULT_DTYPE, NO_OVERLAP
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    hyp = np.ascontiguousarray(hyp, np.float32).reshape(-1, 3)
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(hyp), rank, world)
    _, li, ls = matcher.sweep(xy, hyp[lo:hi], k=k, level=level, want_scores=False)
    li = np.asarray(li, np.int64).copy()
    li[li >= 0] += lo
    gi, _ = combine_topk(li, ls, k)
    kk = int((gi >= 0).sum())
    a, b = shard_range(kk, rank, world)
    kmax = -(-k // world)                                   # records per rank in the gather (padded)
    mine = np.zeros(kmax, RESULT_DTYPE)
    if b > a:
        scans, off = np.tile(xy, (b - a, 1)), np.arange(b - a + 1, dtype=np.int64) * len(xy)
        mine[: b - a] = matcher.align_batch(scans, off, hyp[gi[a:b]].astype(np.float64))
    res = np.zeros(k, RESULT_DTYPE)
    res["status"][kk:] = NO_OVERLAP
    for r, part in enumerate(all_gather_array(mine)):
        ra, rb = shard_range(kk, r, world)
        res[ra:rb] = part[: rb - ra]
    return gi, res


class PeerExchange:
    """Best-hypothesis combine of a sharded sweep through peer memory (include/ndt2d.h, ndt2d_exchange_*).

    Every rank's arg-max kernel stores its best (global index, score) directly into every rank's table over NVLink,
    so the combine is part of the sweep launch and no collective runs per query; torch.distributed is used once, to
    hand the CUDA IPC handles round. `nslots` bounds the number of queries in flight."""

    def __init__(self, matcher, nslots=64):
        self.m = matcher
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.nslots = nslots
        handle = matcher.exchange_create(self.world, self.rank, nslots)
        matcher.exchange_open(all_gather_blobs(handle))
        if self.world > 1:
            dist.barrier()          # every table is open everywhere before anyone publishes

    def publish(self, d_xy, n, d_hyp, nhyp, d_scores, index_offset, query, level=0):
        self.m.sweep_publish(d_xy, n, d_hyp, nhyp, d_scores, index_offset, query, level)

    def wait(self, query, timeout_ms=10000):
        return self.m.exchange_wait(query, timeout_ms)

    def close(self):
        # every kernel this rank queued has finished storing into the peers' tables before anybody frees a table
        self.m.synchronize()
        if self.world > 1:
            dist.barrier()
        self.m.exchange_close()


class PeerRelocalizer:
    """Multi-GPU relocalisation end to end through peer memory (include/ndt2d.h, ndt2d_reloc_* / ndt2d_relocalize_*): every
    rank sweeps its shard, refines its own k best and stores the k candidates into every rank's table; the global top-k
    with its refined records is a local merge - one exchange per query, no collective. torch.distributed is used once,
    to hand the CUDA IPC handles round."""

    def __init__(self, matcher, nslots=16, kmax=8):
        self.m = matcher
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        matcher.reloc_open(all_gather_blobs(matcher.reloc_create(self.world, self.rank, nslots, kmax)))
        if self.world > 1:
            dist.barrier()

    def publish(self, d_xy, n, d_hyp_shard, nhyp_shard, index_offset, k, query, level=0):
        self.m.relocalize_publish(d_xy, n, d_hyp_shard, nhyp_shard, index_offset, k, query, level)

    def wait(self, query, k, timeout_ms=10000):
        return self.m.relocalize_wait(query, k, timeout_ms)

    def close(self):
        self.m.synchronize()
        if self.world > 1:
            dist.barrier()
        self.m.reloc_close()


def align_sharded_counts(nscans):
    """[lo, hi) of this rank's scans for a batched align; results stay on the rank that computed them."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    return shard_range(nscans, rank, world)


# ---- upload relay between the ranks of a job (ndt2d_set_upload_relay) ---------------------------------------------------------

def upload_relay_pairs(gbs, min_ratio=1.15):
    """Pair ranks whose host-to-device copy rates differ. gbs[r] = GB/s rank r gets for its step input when every rank copies
    at the same time (measure it: one plain pinned copy per rank, all at once). The slowest rank is paired with the fastest,
    the second slowest with the second fastest, ...; a pair whose rates differ by more than min_ratio moves the share
    x = (fast - slow) / (fast + slow) of the slow rank's input chunks onto the fast rank's PCIe link - with that share both links
    finish together - and from that GPU over NVLink. Returns {slow_rank: (fast_rank, x)}; pure host logic, the same on every rank."""
    world = len(gbs)
    order = sorted(range(world), key=lambda r: gbs[r])
    pairs = {}
    for i in range(world // 2):
        slow, fast = order[i], order[-1 - i]
        if gbs[fast] > min_ratio * gbs[slow]:
            pairs[slow] = (fast, (gbs[fast] - gbs[slow]) / (gbs[fast] + gbs[slow]))
    return pairs


def refine_relay_fraction(x, t_slow, t_fast):
    """One calibration step for a pair: t_slow / t_fast = the two ranks' step times with share x relayed. The slow rank's time
    scales with the share 1 - x it still copies itself, its partner's with 1 + x; returns the x where the two meet."""
    return min(max(x + (t_slow - t_fast) / (t_slow / (1.0 - x) + t_fast / (1.0 + x)), 0.05), 0.6)

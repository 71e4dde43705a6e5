"""GPU parity: the CUDA path (through the C ABI) against the CPU spec oracle on the same seeded inputs.

Bars (BASELINE.json north_star): cell assignment and point-to-cell indexing bit-exact; final pose within
1e-5 m / 1e-6 rad; score and Hessian within 1e-6 relative. Because SPEC.md fixes every f32 operation AND the
summation order (and, since v3, the solver and the pose's sin/cos), the per-pair factors, the ten sums and whole
result records are compared bit for bit.
PARITY UNPINNED: the oracle restates SPEC.md, not upstream GTSAM-NDT (no source in /root/reference)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POSE_TOL_M, POSE_TOL_RAD, REL_TOL = 1e-5, 1e-6, 1e-6


@pytest.fixture(scope="module")
def mods():
    import gtsam_ndt_b200 as g
    import oracle
    return g, oracle


def make_pair(mods, res, grid=None, **params):
    g, oracle = mods
    m = g.NdtMatcher2D(res, **params)
    o = oracle.Oracle(res, **params)
    if grid:
        m.set_grid(*grid)
        o.set_grid(*grid)
    return m, o


def assert_results_match(rg, ro, exact_iters=True):
    assert np.array_equal(rg["status"], ro["status"])
    if exact_iters:
        assert np.array_equal(rg["iterations"], ro["iterations"])
    assert np.array_equal(rg["count"], ro["count"])
    dp = rg["pose"] - ro["pose"]
    dp[..., 2] = (dp[..., 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(dp[..., :2]).max(initial=0) <= POSE_TOL_M
    assert np.abs(dp[..., 2]).max(initial=0) <= POSE_TOL_RAD
    assert np.allclose(rg["score"], ro["score"], rtol=REL_TOL, atol=1e-12)
    hs = np.abs(ro["hessian"]).reshape(len(np.atleast_1d(ro["score"])), -1).max(1)
    assert np.all(np.abs(rg["hessian"] - ro["hessian"]).reshape(len(hs), -1).max(1) <= REL_TOL * hs + 1e-12)


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("grid", [None, (-100.0, -100.0, 200.0, 200.0)])
def test_grid_build_bit_exact(mods, small_world, overlap, grid):
    m, o = make_pair(mods, [2.0, 0.5], grid, overlap=overlap)
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    for lv in (0, 1):
        assert m.geometry(lv) == o.geometry(lv)
        ng, sg = m.sums(lv); no, so = o.sums(lv)
        assert np.array_equal(ng, no) and np.array_equal(sg, so)
        cg, co = m.cells(lv), o.cells(lv)
        assert cg.tobytes() == co.tobytes()
        assert np.count_nonzero(co[..., 7]) > 100


def test_grid_build_order_independent_and_incremental(mods, small_world):
    xy = small_world["map_xy"]
    m, o = make_pair(mods, [0.25], (-100.0, -100.0, 200.0, 200.0))
    o.set_target(xy)
    perm = np.random.default_rng(9).permutation(len(xy))
    m.set_target(xy[perm])
    assert m.cells().tobytes() == o.cells().tobytes()
    m.set_target(xy[perm[:1000]]); m.add_target(xy[perm[1000:50000]]); m.add_target(xy[perm[50000:]])
    assert m.cells().tobytes() == o.cells().tobytes()
    assert np.array_equal(m.sums()[1], o.sums()[1])


@pytest.mark.parametrize("overlap", [0, 1])
def test_scan_by_scan_mapping_equals_one_shot(mods, small_world, overlap):
    """Incremental mapping as a SLAM front end does it: one scan at a time into a large fixed lattice (the sparse path of
    ndt2d_add_target: only the touched cells are finalised), on a pyramid; every level must equal the oracle's one-shot
    build from the union of the points, and a further dense rebuild must not be disturbed by the dirty words."""
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, [1.0, 0.25], (-100.0, -100.0, 200.0, 200.0), overlap=overlap)
    world = [synth.transform(s, p) for s, p in zip(small_world["scans"][:12], small_world["poses"][:12])]
    m.set_target(world[0])
    for w in world[1:]:
        m.add_target(w)
    o.set_target(np.concatenate(world))
    for lv in (0, 1):
        assert m.cells(lv).tobytes() == o.cells(lv).tobytes()
        assert np.array_equal(m.sums(lv)[1], o.sums(lv)[1])
    m.set_target(np.concatenate(world[:3])); m.add_target(world[3])
    o.set_target(np.concatenate(world[:4]))
    for lv in (0, 1):
        assert m.cells(lv).tobytes() == o.cells(lv).tobytes()


def test_grid_build_edge_inputs(mods):
    m, o = make_pair(mods, [1.0], min_points=3)
    pts = np.array([[np.nan, 1.0], [np.inf, 0.0], [0.2, 0.2], [0.3, 0.4], [0.5, 0.1], [1e30, 1e30]], np.float32)
    for t in (m, o):
        t.set_grid(-2.0, -2.0, 4.0, 4.0); t.set_target(pts)
    assert m.cells().tobytes() == o.cells().tobytes() and m.sums()[0].sum() == 3
    for t in (m, o):
        t.set_grid(0, 0, 0, 0); t.set_target(np.zeros((0, 2), np.float32))    # empty target, auto-fit
    assert m.geometry() == o.geometry() and m.cells().tobytes() == o.cells().tobytes()
    # every point in one cell, many duplicates: contended atomics stay exact
    dup = np.tile(np.array([[0.3, 0.3], [0.31, 0.33], [0.36, 0.3]], np.float32), (40000, 1))
    for t in (m, o):
        t.set_grid(-2.0, -2.0, 4.0, 4.0); t.set_target(dup)
    assert np.array_equal(m.sums()[1], o.sums()[1]) and m.cells().tobytes() == o.cells().tobytes()


def test_cell_index_bit_exact(mods, small_world):
    m, o = make_pair(mods, [0.25], (-100.0, -100.0, 200.0, 200.0))
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    rng = np.random.default_rng(0)
    k = rng.integers(-5, 805, size=(50000, 2))
    pts = (np.float32(-100.0) + k.astype(np.float32) * np.float32(0.25)).astype(np.float32)
    nudge = rng.integers(-2, 3, size=pts.shape)
    pts = np.where(nudge > 0, np.nextafter(pts, np.float32(np.inf)), np.where(nudge < 0, np.nextafter(pts, np.float32(-np.inf)), pts))
    pts = np.concatenate([pts, np.array([[np.nan, 0], [0, np.inf], [100.0, 0.0], [-100.0, -100.0]], np.float32)]).astype(np.float32)
    assert np.array_equal(m.cell_index(pts), o.cell_index(pts))
    for i in range(4):
        xy, pose = small_world["scans"][i], small_world["init"][i]
        a, b = m.cell_index(xy, pose), o.cell_index(xy, pose)
        assert np.array_equal(a, b) and (a >= 0).sum() > 900


@pytest.mark.parametrize("overlap", [0, 1])
def test_point_terms_bit_exact_and_sums(mods, small_world, overlap):
    m, o = make_pair(mods, [0.5], None, overlap=overlap)
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    for i in (0, 7, 13):
        xy, pose = small_world["scans"][i], small_world["init"][i]
        tg, to = m.point_terms(xy, pose), o.point_terms(xy, pose)
        assert tg.tobytes() == to.tobytes()
        eg, cg = m.evaluate(xy, pose); eo, co = o.evaluate(xy, pose)
        assert cg == co and np.count_nonzero(to[..., 0]) == co
        assert np.array_equal(eg, eo)            # fixed summation order: the sums are bit-identical


def test_evaluate_many_poses(mods, small_world):
    m, o = make_pair(mods, [1.0, 0.25], (-100.0, -100.0, 200.0, 200.0))
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    xy = small_world["scans"][2]
    poses = small_world["init"][2] + np.random.default_rng(4).normal(size=(300, 3)) * [0.2, 0.2, 0.02]
    for lv in (0, 1):
        eg, cg = m.evaluate(xy, poses, level=lv)
        ref = [o.evaluate(xy, p, level=lv) for p in poses]
        eo = np.array([r[0] for r in ref]); co = np.array([r[1] for r in ref])
        assert np.array_equal(cg, co)
        assert np.array_equal(eg, eo)


@pytest.mark.parametrize("cfg", [dict(res=[0.5], overlap=0), dict(res=[0.5], overlap=1), dict(res=[2.0, 1.0, 0.5], overlap=0),
                                 dict(res=[1.0, 0.25], overlap=1)])
def test_align_batch_matches_oracle(mods, small_world, cfg):
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, cfg["res"], None, overlap=cfg["overlap"])
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    xy, off = synth.pack(small_world["scans"])
    rg = m.align_batch(xy, off, small_world["init"])
    ro = o.align_batch(xy, off, small_world["init"])
    assert_results_match(rg, ro)
    assert all(rg[i].tobytes() == ro[i].tobytes() for i in range(len(rg)))   # SPEC v3 fixes every operation: identical bytes
    assert np.all(rg["status"] <= 1) and (rg["status"] == 0).mean() > 0.8
    err = rg["pose"] - small_world["poses"]
    err[:, 2] = (err[:, 2] + np.pi) % (2 * np.pi) - np.pi
    good = rg["status"] == 0
    assert np.median(np.abs(err[:, :2])) < 0.01 and np.percentile(np.abs(err[good, :2]), 80) < 0.02
    one = m.align(small_world["scans"][5], small_world["init"][5])
    assert one.tobytes() == rg[5].tobytes()


def test_align_config1_scan_to_scan_360(mods):
    """BASELINE.json configs[0]: single synthetic 360-beam scan-to-scan align, 0.5 m cells."""
    from gtsam_ndt_b200 import synth
    r, p = synth.scans(2, traj_len=4000, first=11, step=1, **synth.SCAN_360)
    a, b = (synth.polar_to_points(r[i], synth.SCAN_360["angle_min"], synth.SCAN_360["angle_inc"]) for i in (0, 1))
    m, o = make_pair(mods, [0.5])
    m.set_target(a); o.set_target(a)
    rg, ro = m.align(b, [0, 0, 0]), o.align(b, [0, 0, 0])
    assert_results_match(np.array([rg]), np.array([ro]))
    assert rg["status"] == 0 and np.all(np.linalg.eigvalsh(rg["hessian"]) > 0)
    # relative motion between consecutive trajectory poses, in the frame of the first
    c, s = math.cos(p[0, 2]), math.sin(p[0, 2])
    d = p[1] - p[0]
    rel = np.array([c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2]])
    assert np.linalg.norm(rg["pose"][:2] - rel[:2]) < np.linalg.norm(rel[:2]) and abs(rg["pose"][2] - rel[2]) < 2e-3
    # SURVEY 8(d) config 1 offset (0.10 m, -0.05 m, 2 deg), coarse-to-fine: the truth is recovered
    T = np.array([0.10, -0.05, math.radians(2.0)])
    ci, si = math.cos(-T[2]), math.sin(-T[2])
    inv = np.array([-(ci * T[0] - si * T[1]), -(si * T[0] + ci * T[1]), -T[2]])
    moved = synth.transform(a, inv)
    m3, o3 = make_pair(mods, [2.0, 1.0, 0.5])
    m3.set_target(a); o3.set_target(a)
    rg, ro = m3.align(moved, [0, 0, 0]), o3.align(moved, [0, 0, 0])
    assert_results_match(np.array([rg]), np.array([ro]))
    assert np.allclose(rg["pose"][:2], T[:2], atol=0.02) and abs(rg["pose"][2] - T[2]) < 5e-3


def test_align_edge_cases(mods, small_world):
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, [1.0, 0.5])
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    scans = [np.zeros((0, 2), np.float32),                          # empty
             small_world["scans"][0][:1],                           # one point
             small_world["scans"][1] + np.float32(4000.0),          # no overlap
             small_world["scans"][2][:33],                          # ragged
             small_world["scans"][3]]
    init = np.array([[1, 2, 0.3], small_world["init"][0], [0, 0, 0], small_world["init"][2], small_world["init"][3]])
    xy, off = synth.pack(scans)
    rg, ro = m.align_batch(xy, off, init), o.align_batch(xy, off, init)
    assert_results_match(rg, ro)
    assert rg["status"][0] == 3 and rg["status"][2] == 3 and list(rg["pose"][0]) == [1.0, 2.0, 0.3]
    assert rg["iterations"][0] == 2        # one evaluation per level
    # a scan longer than the shared-memory staging slot takes the global-memory path
    big = np.concatenate([small_world["scans"][i] for i in range(5)])
    assert len(big) > 3700
    rg1, ro1 = m.align(big, small_world["init"][2]), o.align(big, small_world["init"][2])
    assert_results_match(np.array([rg1]), np.array([ro1]))
    assert m.align_batch(np.zeros((0, 2), np.float32), [0], np.zeros((0, 3))).shape == (0,)


def test_unusable_points_are_ignored(mods, small_world):
    """SPEC 4: NaN / Inf / huge scan points are replaced by a far-away point; nothing becomes NaN."""
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, [1.0, 0.5], None, overlap=1)
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    bad = np.array([[np.nan, 1.0], [1.0, np.inf], [-np.inf, np.nan], [1e30, 0.0], [3e38, -3e38]], np.float32)
    xy = np.concatenate([small_world["scans"][0][:500], bad, small_world["scans"][0][500:]])
    pose = small_world["init"][0]
    eg, cg = m.evaluate(xy, pose, level=1); eo, co = o.evaluate(xy, pose, level=1)
    assert cg == co and np.array_equal(eg, eo) and np.all(np.isfinite(eg))
    assert np.array_equal(m.cell_index(bad, pose), o.cell_index(bad, pose))
    rg, ro = m.align(xy, pose), o.align(xy, pose)
    assert_results_match(np.array([rg]), np.array([ro]))
    assert np.all(np.isfinite(rg["hessian"])) and rg["status"] == 0
    allbad = np.tile(bad, (30, 1))
    rg = m.align(allbad, pose)
    assert rg["status"] == 3 and rg["count"] == 0


@pytest.mark.parametrize("overlap", [0, 1])
def test_lattice_edges_in_the_gather_path(mods, overlap):
    """The evaluation kernels test `inside` on integers (floor-to-int, unsigned compare) and send outside points to a
    sentinel record; SPEC 2 states it on floats. Points exactly on, one ulp below and one ulp above every kind of
    lattice boundary (low edge, high edge, interior cell edges, (-1, 0), huge, -0.0) must give the oracle's sums."""
    grid = (-8.0, -6.0, 16.0, 12.0)
    m, o = make_pair(mods, [0.5], grid, overlap=overlap)
    rng = np.random.default_rng(11)
    tgt = np.stack([rng.uniform(-8, 8, 60000), rng.uniform(-6, 6, 60000)], 1).astype(np.float32)
    m.set_target(tgt); o.set_target(tgt)
    def around(v):
        v = np.float32(v)
        return [np.nextafter(v, np.float32(-np.inf)), v, np.nextafter(v, np.float32(np.inf))]
    xs = sum((around(v) for v in (-8.0, -8.25, -7.5, 0.0, 7.5, 7.75, 8.0, 8.25)), []) + [np.float32(-0.0), np.float32(1e9), np.float32(-1e9)]
    ys = sum((around(v) for v in (-6.0, -6.25, -5.5, 0.0, 5.5, 5.75, 6.0, 6.25)), []) + [np.float32(-0.0), np.float32(1e9), np.float32(-1e9)]
    pts = np.array([[x, y] for x in xs for y in ys], np.float32)
    pose = np.zeros(3)                                   # identity: transformed points are the points themselves
    assert np.array_equal(m.cell_index(pts, pose), o.cell_index(pts, pose))
    eg, cg = m.evaluate(pts, pose); eo, co = o.evaluate(pts, pose)
    assert cg == co and co > 0 and np.array_equal(eg, eo)
    hyp = np.array([[0, 0, 0], [1e-3, -1e-3, 1e-4], [0.25, 0.25, 0.0]], np.float32)
    sg, _, _ = m.sweep(pts, hyp, k=1); so = o.sweep(pts, hyp)[0]
    assert np.array_equal(sg, so)
    rg, ro = m.align(pts, pose), o.align(pts, pose)      # the LM loop on the same degenerate scan agrees as well
    assert_results_match(np.array([rg]), np.array([ro]))


def test_align_permutation_invariant_and_idempotent(mods, small_world):
    from gtsam_ndt_b200 import synth
    m, _ = make_pair(mods, [1.0, 0.5])
    m.set_target(small_world["map_xy"])
    xy, off = synth.pack(small_world["scans"])
    r1 = m.align_batch(xy, off, small_world["init"])
    perm = np.random.default_rng(3).permutation(len(small_world["scans"]))
    xy2, off2 = synth.pack([small_world["scans"][i] for i in perm])
    r2 = m.align_batch(xy2, off2, small_world["init"][perm])
    assert r2.tobytes() == r1[perm].tobytes()                      # one warp per scan: order cannot matter
    r3 = m.align_batch(xy, off, r1["pose"])
    d = r3["pose"] - r1["pose"]
    assert np.abs(d[:, :2]).max() < 2e-3 and np.abs(d[:, 2]).max() < 2e-4
    assert np.all(r3["score"] >= r1["score"] * (1 - 5e-4))   # the coarse level may move the pose to a neighbouring optimum of the fine level


@pytest.mark.parametrize("u16", [False, True])
def test_align_ranges_input(mods, small_world, u16):
    """SPEC 8: LaserScan input converted on the device equals the oracle run on converted points."""
    from gtsam_ndt_b200 import synth
    import oracle
    m, o = make_pair(mods, [1.0, 0.5])
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    amin, ainc = synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"]
    r = small_world["ranges"].copy()
    r[3, 100:400] = 0.0; r[5, ::7] = np.nan; r[6, :] = 0.0
    if u16:
        r = np.nan_to_num(r * 500.0, nan=0.0).clip(0, 65535).round().astype(np.uint16)
        kw = dict(range_scale=0.002, range_min=0.5, range_max=60.0)
    else:
        kw = dict(range_scale=1.0, range_min=0.5, range_max=60.0)
    rg = m.align_batch_ranges(r, amin, ainc, small_world["init"], **kw)
    pts = [oracle.polar_to_points(r[i], amin, ainc, **kw) for i in range(len(r))]
    xy, off = synth.pack(pts)
    ro = o.align_batch(xy, off, small_world["init"])
    assert_results_match(rg, ro)
    assert rg["status"][6] == 3 and rg["count"][3] < rg["count"][2]


@pytest.mark.parametrize("overlap", [0, 1])
def test_sweep_matches_oracle(mods, small_world, overlap):
    m, o = make_pair(mods, [1.0], None, overlap=overlap)
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    xy, truth = small_world["scans"][4], small_world["poses"][4]
    g = np.stack(np.meshgrid(np.arange(-6, 7) * 0.2, np.arange(-6, 7) * 0.2, np.radians(np.arange(-5, 6) * 3.0), indexing="ij"), -1).reshape(-1, 3)
    hyp = (truth + g).astype(np.float32)
    hyp = np.concatenate([hyp, hyp[:200]])                        # duplicates: ties go to the smaller index
    sg, bi, bs = m.sweep(xy, hyp, k=8)
    so, oi, os_ = o.sweep(xy, hyp)
    assert np.array_equal(sg, so)
    assert bi[0] == oi and bs[0] == sg[oi]
    order = np.lexsort((np.arange(len(sg)), -sg))[:8]
    assert np.array_equal(bi, order) and np.array_equal(bs, sg[order])
    assert np.allclose(hyp[bi[0]], truth, atol=1e-3)
    _, bi2, _ = m.sweep(xy, hyp[:5], k=8, want_scores=False)     # k > nhyp pads with -1
    assert list(bi2[5:]) == [-1, -1, -1] and sorted(bi2[:5]) == [0, 1, 2, 3, 4]


@pytest.mark.parametrize("cfg", [dict(res=[0.5]), dict(res=[2.0, 1.0, 0.5]), dict(res=[1.0, 0.5], overlap=1)])
def test_block_and_warp_align_kernels_agree(mods, small_world, cfg, monkeypatch):
    """Calls with few scans run one BLOCK per scan (low latency), large batches one WARP per scan (throughput). Both
    must give the same bytes, and the oracle's: the same ragged batch through a handle forced to each kernel."""
    from gtsam_ndt_b200 import synth
    g, oracle = mods
    scans = [s[:: 1 + i % 3][: len(s) - 7 * i] for i, s in enumerate(small_world["scans"])] + [np.zeros((0, 2), np.float32), small_world["scans"][0][:5]]
    init = np.vstack([small_world["init"], small_world["init"][:2]])
    xy, off = synth.pack(scans)
    out = {}
    for mode, env in (("warp", "0"), ("block", "100000")):
        monkeypatch.setenv("NDT2D_BLOCK_ALIGN_MAX", env)
        m = g.NdtMatcher2D(cfg["res"], overlap=cfg.get("overlap", 0))
        m.set_target(small_world["map_xy"])
        out[mode] = m.align_batch(xy, off, init)
        one = m.align(scans[3], init[3])
        assert one.tobytes() == out[mode][3].tobytes()
    o = oracle.Oracle(cfg["res"], overlap=cfg.get("overlap", 0))
    o.set_target(small_world["map_xy"])
    ro = o.align_batch(xy, off, init)
    assert out["warp"].tobytes() == out["block"].tobytes() == ro.tobytes()


@pytest.mark.parametrize("cfg", [dict(res=[0.5]), dict(res=[2.0, 1.0, 0.5]), dict(res=[0.25], reps=40), dict(res=[1.0, 0.5], reps=150)])
def test_helper_warps_leave_every_bit_unchanged(mods, small_world, cfg, monkeypatch):
    """k_align with helper warps (NDT2D_ALIGN_HELP=1: warps that find the queue empty compute the factors of the last steps
    of a block-mate's evaluation) against the same kernel without them and against the oracle: ragged scans incl. an empty
    and a five-point one, one level and pyramids, batches smaller than the grid (every owner has three helpers from the
    start), larger than one wave of blocks (helpers appear as block-mates finish) and LaserScan input."""
    from gtsam_ndt_b200 import synth
    g, oracle = mods
    base = [s[:: 1 + i % 3][: len(s) - 7 * i] for i, s in enumerate(small_world["scans"])] + [np.zeros((0, 2), np.float32), small_world["scans"][0][:5]]
    binit = np.vstack([small_world["init"], small_world["init"][:2]])
    reps = cfg.get("reps", 1)
    scans = base * reps
    init = np.tile(binit, (reps, 1)) + (np.arange(len(scans)) % 7)[:, None] * np.array([0.004, -0.003, 0.0005])
    xy, off = synth.pack(scans)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NDT2D_BLOCK_ALIGN_MAX", "0")
        monkeypatch.setenv("NDT2D_ALIGN_HELP", mode)
        m = g.NdtMatcher2D(cfg["res"])
        m.set_target(small_world["map_xy"])
        out[mode] = m.align_batch(xy, off, init)
        m.close()
    assert out["0"].tobytes() == out["1"].tobytes()
    if reps == 1:
        o = oracle.Oracle(cfg["res"])
        o.set_target(small_world["map_xy"])
        assert o.align_batch(xy, off, init).tobytes() == out["1"].tobytes()
    assert out["1"]["iterations"].max() > out["1"]["iterations"].min()        # unequal scans: some warps do turn helper


def test_helper_warps_ranges_input(mods, small_world, monkeypatch):
    """The LaserScan form of the same check (the helpers read the owner's converted points), and the block-per-scan kernel's
    LaserScan form (warp 0 converts the beams) against both."""
    from gtsam_ndt_b200 import synth
    g, _ = mods
    sc = synth.SCAN_1080
    ranges, poses = synth.scans(37, traj_len=600, first=3, step=13, **sc)
    init = poses + synth.uniform3(37) * np.array([0.05, 0.05, np.radians(0.5)])
    out = {}
    u16 = np.round(ranges / 0.004).clip(0, 65535).astype(np.uint16)
    u16[3, 100:140] = 0                                   # no return: those beams are dropped (SPEC 8)
    for mode, env in (("warp", ("0", "0")), ("help", ("0", "1")), ("block", ("100000", "0"))):
        monkeypatch.setenv("NDT2D_BLOCK_ALIGN_MAX", env[0])
        monkeypatch.setenv("NDT2D_ALIGN_HELP", env[1])
        m = g.NdtMatcher2D([1.0, 0.5])
        m.set_target(small_world["map_xy"])
        out[mode] = (m.align_batch_ranges(ranges, sc["angle_min"], sc["angle_inc"], init, range_scale=1.0).tobytes(),
                     m.align_batch_ranges(u16, sc["angle_min"], sc["angle_inc"], init, range_scale=0.004, range_min=0.5, range_max=25.0).tobytes(),
                     m.align_batch_ranges(ranges[:1], sc["angle_min"], sc["angle_inc"], init[:1], range_scale=1.0).tobytes())
        m.close()
    assert out["warp"] == out["help"] == out["block"]      # the three kernel forms, f32 and u16 ranges, one scan (small-call path)
    first = np.frombuffer(out["block"][0], dtype=g.RESULT_DTYPE)
    assert (first["status"] != g.NO_OVERLAP).any() and out["block"][2] == first[:1].tobytes()


def _between(a, b):
    c, s_ = math.cos(a[2]), math.sin(a[2])
    dx, dy = b[0] - a[0], b[1] - a[1]
    return np.array([c * dx + s_ * dy, -s_ * dx + c * dy, b[2] - a[2]])


@pytest.mark.parametrize("cfg", [dict(res=[0.5]), dict(res=[2.0, 1.0, 0.5]), dict(res=[1.0, 0.5], overlap=1),
                                 dict(res=[0.5], grid=(-120.0, -120.0, 240.0, 240.0)), dict(res=[1.0], chunk_bytes=3_000_000, fused=0),
                                 dict(res=[0.5], fused=0), dict(res=[2.0, 0.25], fused=0), dict(res=[0.3, 0.1]), dict(res=[0.5], min_points=2),
                                 dict(res=[1.0, 0.5], min_points=5, eig_ratio=0.05)])
def test_align_pairs_equals_set_target_plus_align(mods, cfg, monkeypatch):
    """Batched scan-to-scan: every pair of ndt2d_align_pairs must be byte-identical to ndt2d_set_target(target scan) +
    ndt2d_align(source scan) on the GPU, and to the same two calls of the CPU spec oracle. Covers both implementations
    (fused: the target's grid built by the aligning warp in shared memory, K = 1; general: per-target hash tables in global
    memory), pyramids, overlapping grids, an explicit lattice, fine cells (many radix passes), other min_points, repeated
    and self targets, empty / tiny / unusable scans, and chunking of the table budget."""
    from gtsam_ndt_b200 import synth
    cfg = dict(cfg)
    grid = cfg.pop("grid", None)
    chunk = cfg.pop("chunk_bytes", None)
    res = cfg.pop("res")
    if chunk:
        monkeypatch.setenv("NDT2D_PAIRS_BYTES", str(chunk))
    if cfg.pop("fused", 1) == 0:        # K = 1 takes the fused shared-memory path by default; this forces the general (global hash table) path
        monkeypatch.setenv("NDT2D_PAIRS_FUSED", "0")
    m, o = make_pair(mods, res, grid, **cfg)
    sc = synth.SCAN_1080
    ranges, poses = synth.scans(14, traj_len=4000, first=50, step=9, sigma=0.01, **sc)     # neighbours 0.85 m apart
    scans = synth.polar_to_points(ranges, sc["angle_min"], sc["angle_inc"])
    scans = [s[:: (1 + i % 3)] for i, s in enumerate(scans)]                                # ragged: 1080 / 540 / 360 points
    scans += [np.zeros((0, 2), np.float32), scans[0][:4], np.full((30, 2), np.nan, np.float32)]
    E, T, N = 14, 15, 16                                                                    # empty, tiny, unusable
    pairs = [(i, i + 1) for i in range(13)] + [(i + 1, i) for i in range(0, 13, 3)] + [(5, 5), (0, 4), (0, 2), (3, 9),
             (E, 0), (0, E), (T, 1), (1, T), (N, 2), (2, N), (E, E)]
    rng = np.random.default_rng(3)
    pp = np.vstack([poses, np.zeros((3, 3))])
    init = np.array([_between(pp[t], pp[s]) + rng.normal(size=3) * [0.03, 0.03, 0.003] for t, s in pairs])
    xy, off = synth.pack(scans)
    rg = m.align_pairs(xy, off, pairs, init)
    if min(res) >= 0.5:
        assert (rg["status"][:13] == 0).sum() >= 11                                         # consecutive scans do align
    for p, (t, s_) in enumerate(pairs):
        m.set_target(scans[t]); o.set_target(scans[t])
        one = m.align(scans[s_], init[p]); ref = o.align(scans[s_], init[p])
        assert rg[p].tobytes() == one.tobytes(), (p, t, s_, rg[p], one)
        assert rg[p].tobytes() == ref.tobytes(), (p, t, s_, rg[p], ref)


@pytest.mark.parametrize("decimate", [3, 8, 24])
def test_align_pairs_small_targets(mods, decimate):
    """Calls whose longest target is short: the fused build's radix sort then runs with narrower digits (its counter matrix must
    fit the smaller record + index areas: 6 bits at 360 points, 5 bits below ~170) and more passes. Still set_target + align."""
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, [1.0, 0.5], None)
    sc = synth.SCAN_1080
    ranges, poses = synth.scans(8, traj_len=4000, first=80, step=7, sigma=0.01, **sc)
    scans = [s[::decimate] for s in synth.polar_to_points(ranges, sc["angle_min"], sc["angle_inc"])]
    pairs = [(i, i + 1) for i in range(7)] + [(4, 4), (6, 1)]
    rng = np.random.default_rng(11)
    init = np.array([_between(poses[t], poses[s]) + rng.normal(size=3) * [0.03, 0.03, 0.003] for t, s in pairs])
    xy, off = synth.pack(scans)
    rg = m.align_pairs(xy, off, pairs, init)
    for p, (t, s_) in enumerate(pairs):
        m.set_target(scans[t]); o.set_target(scans[t])
        one = m.align(scans[s_], init[p]); ref = o.align(scans[s_], init[p])
        assert rg[p].tobytes() == one.tobytes() == ref.tobytes(), (p, t, s_, rg[p], one, ref)


@pytest.mark.parametrize("overlap", [0, 1])
def test_align_pairs_large_scans(mods, overlap):
    """Targets with more occupied cells than the build warp's claimed-slot list holds (full-table finalisation) and
    sources too long for the shared-memory slot (points read from global memory): still set_target + align, byte for byte."""
    from gtsam_ndt_b200 import synth
    m, o = make_pair(mods, [1.0], None, overlap=overlap)
    rng = np.random.default_rng(21)
    big = np.stack([rng.uniform(-60, 60, 4200), rng.uniform(-60, 60, 4200)], 1).astype(np.float32)      # ~3700 distinct 1 m cells
    wall = np.stack([np.linspace(-40, 40, 4200), 5.0 + 0.02 * rng.normal(size=4200)], 1).astype(np.float32)
    scans = [big, wall, (wall + np.float32([0.1, 0.05])).astype(np.float32), big[::2] + np.float32(0.2)]
    pairs = [(0, 3), (1, 2), (2, 1), (0, 0), (3, 0)]
    init = np.array([[0.1, 0.1, 0.0], [0.0, 0.0, 0.0], [0.05, 0.0, 0.001], [0.0, 0.0, 0.0], [-0.2, -0.2, 0.0]])
    xy, off = synth.pack(scans)
    rg = m.align_pairs(xy, off, pairs, init)
    for p, (t, s_) in enumerate(pairs):
        m.set_target(scans[t]); o.set_target(scans[t])
        assert rg[p].tobytes() == m.align(scans[s_], init[p]).tobytes() == o.align(scans[s_], init[p]).tobytes(), p


def test_sweep_publish_world1_equals_sweep(mods, small_world):
    """The peer-memory exchange with a single rank: the arg-max kernel publishes into the rank's own table and the host
    poll returns what ndt2d_sweep returns (index shifted by index_offset); slots are reused as the epochs grow."""
    import torch
    m, o = make_pair(mods, [0.5], None)
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    xy = small_world["scans"][1]
    rng = np.random.default_rng(5)
    handle = m.exchange_create(1, 0, nslots=4)
    assert len(handle) == 64
    m.exchange_open([handle])
    dev = torch.device("cuda", 0)
    d_xy = torch.from_numpy(np.ascontiguousarray(xy, np.float32)).to(dev)
    for q in range(10):
        hyp = (small_world["init"][1] + rng.normal(size=(700 + q, 3)) * [0.3, 0.3, 0.03]).astype(np.float32)
        hyp[5] = hyp[3]                                  # an exact tie: the smaller index wins
        d_hyp = torch.from_numpy(hyp).to(dev)
        m.sweep_publish(d_xy, len(xy), d_hyp, len(hyp), None, 1000 * q, q)
        bi, bs = m.exchange_wait(q, timeout_ms=5000)
        _, ri, rs = m.sweep(xy, hyp, k=1, want_scores=False)
        so, oi, os_ = o.sweep(xy, hyp)
        assert bi == int(ri[0]) + 1000 * q == oi + 1000 * q and bs == float(rs[0]) == os_
        assert m.exchange_wait(q) == (bi, bs)            # answered from the verified snapshot
    with pytest.raises(Exception):
        m.exchange_wait(12345, timeout_ms=50)            # never published: NDT2D_ETIMEOUT, reported, no hang
    m.exchange_close()
    with pytest.raises(Exception):
        m.sweep_publish(d_xy, len(xy), d_hyp, len(hyp), None, 0, 0)


def test_relocalize_refines_topk(mods, small_world):
    m, o = make_pair(mods, [2.0, 0.5])
    m.set_target(small_world["map_xy"]); o.set_target(small_world["map_xy"])
    xy, truth = small_world["scans"][8], small_world["poses"][8]
    g = np.stack(np.meshgrid(np.arange(-5, 6) * 0.5, np.arange(-5, 6) * 0.5, np.radians(np.arange(-4, 5) * 4.0), indexing="ij"), -1).reshape(-1, 3)
    hyp = (truth + g + [0.07, -0.04, 0.004]).astype(np.float32)
    bi, res = m.relocalize(xy, hyp, k=4, level=0)
    _, oi, _ = o.sweep(xy, hyp, level=0)
    assert bi[0] == oi
    for j in range(4):
        ro = o.align(xy, hyp[bi[j]].astype(np.float64))
        assert_results_match(np.array([res[j]]), np.array([ro]))
        assert res[j].tobytes() == ro.tobytes()          # the k refinements are ordinary aligns, bit for bit
    best = res[np.argmax(res["score"])]
    dth = (best["pose"][2] - truth[2] + np.pi) % (2 * np.pi) - np.pi
    assert np.allclose(best["pose"][:2], truth[:2], atol=0.03) and abs(dth) < 3e-3
    # more candidates asked for than there are hypotheses: the surplus entries are empty
    bi2, res2 = m.relocalize(xy, hyp[:3], k=5, level=0)
    assert list(bi2[3:]) == [-1, -1] and np.all(res2["status"][3:] == 3) and np.all(res2["iterations"][3:] == 0)
    assert all(res2[j].tobytes() == o.align(xy, hyp[bi2[j]].astype(np.float64)).tobytes() for j in range(3))
    # the all-device form (ndt2d_relocalize_device): same indices and records, one stream, no host round trip in between
    import torch
    g, _ = mods
    dev = torch.device("cuda", 0)
    d_xy, d_hyp = torch.from_numpy(xy).to(dev), torch.from_numpy(hyp).to(dev)
    d_idx = torch.zeros(4, dtype=torch.int64, device=dev)
    d_res = torch.zeros(4 * 144, dtype=torch.uint8, device=dev)
    m.relocalize_device(d_xy, len(xy), d_hyp, len(hyp), 4, d_idx, d_res, level=0)
    m.synchronize()
    assert np.array_equal(d_idx.cpu().numpy(), bi)
    assert d_res.cpu().numpy().tobytes() == res.tobytes()
    # the sharded form through the peer-memory candidate exchange, world = 1 here (2..8 ranks: tools/mgpu_exchange_check.py):
    # publish (sweep + refinement of the shard's k best + stores into the table) and wait (merge) give the same answer;
    # slots are reused by later queries
    from gtsam_ndt_b200 import distributed as D
    pr = D.PeerRelocalizer(m, nslots=2, kmax=4)
    for q in range(5):
        pr.publish(d_xy, len(xy), d_hyp, len(hyp), 0, 4, q)
        gi, gr = pr.wait(q, 4)
        assert np.array_equal(gi, bi) and gr.tobytes() == res.tobytes()
    pr.publish(d_xy, len(xy), d_hyp, 2, 7, 3, 5)       # a 2-hypothesis shard at offset 7, k = 3
    gi, gr = pr.wait(5, 3)
    assert list(gi) == [int(b) + 7 for b in m.relocalize(xy, hyp[:2], k=3)[0][:2]] + [-1] and gr["status"][2] == 3
    with pytest.raises(g.NdtError, match="kmax"):
        pr.publish(d_xy, len(xy), d_hyp, len(hyp), 0, 5, 6)
    pr.close()


def test_set_cells_roundtrip(mods, small_world):
    g, _ = mods
    m, _ = make_pair(mods, [0.5], (-100.0, -100.0, 200.0, 200.0))
    m.set_target(small_world["map_xy"])
    cells = m.cells()
    r1 = m.align(small_world["scans"][1], small_world["init"][1])
    m2, _ = make_pair(mods, [0.5], (-100.0, -100.0, 200.0, 200.0))
    m2.set_cells(cells)
    assert m2.cells().tobytes() == cells.tobytes()
    assert m2.align(small_world["scans"][1], small_world["init"][1]).tobytes() == r1.tobytes()
    with pytest.raises(g.NdtError, match="no sums"):
        m2.add_target(small_world["map_xy"][:10])
    # a table of another size is refused (it used to be copied into the larger table, rows sheared)
    with pytest.raises(g.NdtError, match="records"):
        m2.set_cells(cells[:-1])
    with pytest.raises(ValueError):     # right record count, wrong shape: rows would be sheared
        m2.set_cells(cells.reshape(cells.shape[0] * 2, cells.shape[1] // 2, 8))


@pytest.mark.parametrize("cfg", [dict(res=[2.0, 0.5], overlap=1, grid=(-100.0, -100.0, 200.0, 200.0)),
                                 dict(res=[2.0, 0.5], overlap=0, grid=None),            # auto-fitted: every level has its own origin and size
                                 dict(res=[0.3], overlap=0, grid=None),                 # ceil((nhx*st)/st) != nhx for many sizes at 0.3 m
                                 dict(res=[1.0, 0.3], overlap=1, grid=(-71.3, -64.9, 150.4, 133.3))])
def test_save_and_load_map(mods, small_world, tmp_path, cfg):
    """ndt2d_save_map / ndt2d_load_map (C ABI): every level's lattice is restored exactly as it was built, the records and
    the integer sums with it, so aligns are byte-identical and add_target continues the loaded map bit for bit."""
    g, oracle = mods
    xy = small_world["map_xy"]
    m = g.NdtMatcher2D(cfg["res"], overlap=cfg["overlap"])
    if cfg["grid"]:
        m.set_grid(*cfg["grid"])
    m.set_target(xy[:60000])
    path = str(tmp_path / "map.ndt2d")
    m.save_map(path)
    m2 = g.NdtMatcher2D([1.0])
    m2.load_map(path)
    assert m2.nlevels == len(cfg["res"]) and m2.params.overlap == cfg["overlap"]
    for lv in range(m.nlevels):
        assert m2.geometry(lv) == m.geometry(lv)
        assert m2.cells(lv).tobytes() == m.cells(lv).tobytes()
        assert all(np.array_equal(a, b) for a, b in zip(m2.sums(lv), m.sums(lv)))
    a = m.align(small_world["scans"][3], small_world["init"][3])
    b = m2.align(small_world["scans"][3], small_world["init"][3])
    assert a.tobytes() == b.tobytes()
    # the loaded map is extended exactly like the original one (points outside an auto-fitted lattice are ignored by both)
    m.add_target(xy[60000:]); m2.add_target(xy[60000:])
    for lv in range(m.nlevels):
        assert m2.cells(lv).tobytes() == m.cells(lv).tobytes()
    # without sums: aligns still identical, extending is refused
    m.save_map(path, with_sums=False)
    m3 = g.NdtMatcher2D([0.7], overlap=1 - cfg["overlap"])
    m3.load_map(path)
    assert m3.align(small_world["scans"][5], small_world["init"][5]).tobytes() == m.align(small_world["scans"][5], small_world["init"][5]).tobytes()
    with pytest.raises(g.NdtError, match="no sums"):
        m3.add_target(xy[:10])
    with pytest.raises(g.NdtError, match="no sums"):
        m3.save_map(path, with_sums=True)
    open(path, "wb").write(b"not a map file at all")
    with pytest.raises(g.NdtError, match="NDT2DMAP"):
        m3.load_map(path)


def test_wrapper_size_checks(mods, small_world):
    """the Python mirror validates sizes before raw pointers reach the C side."""
    g, _ = mods
    from gtsam_ndt_b200 import synth
    m = g.NdtMatcher2D([0.5])
    m.set_target(small_world["map_xy"])
    xy, off = synth.pack(small_world["scans"][:3])
    init = small_world["init"][:3]
    with pytest.raises(ValueError):
        m.align_batch(xy[:100], off, init)                       # offsets overrun xy
    with pytest.raises(ValueError):
        m.align_batch(xy, off, init, out=np.zeros(2, g.RESULT_DTYPE))   # short result buffer
    with pytest.raises(ValueError):
        m.align_batch(xy, off, init, out=np.zeros(3 * 144, np.uint8))   # wrong dtype
    with pytest.raises(ValueError):
        m.align_pairs(xy[:100], off, [[0, 1]], init[:1])          # offsets overrun xy
    with pytest.raises(ValueError):
        m.align_batch_ranges(small_world["ranges"][:3], -2.0, 0.004, init[:2])   # one initial pose per scan


def test_errors_are_reported(mods):
    g, _ = mods
    m = g.NdtMatcher2D([0.5])
    with pytest.raises(g.NdtError, match="no target"):
        m.align(np.zeros((4, 2), np.float32), [0, 0, 0])
    with pytest.raises(g.NdtError):
        m.set_resolutions([0.0])
    with pytest.raises(g.NdtError):
        m.set_params(min_points=1)
    n0 = m.kernel_launches
    m.set_target(np.random.default_rng(0).normal(size=(100, 2)).astype(np.float32))
    assert m.kernel_launches >= n0 + 2          # accumulate + finalise (the bounding box of a small host target is taken on the host)

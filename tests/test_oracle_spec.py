"""Pins the SPEC ORACLE without a reference: closed-form cells, an independent
numpy f64 model, finite differences, structural identities. PARITY UNPINNED: none of this compares with
upstream GTSAM-NDT arithmetic, because /root/reference holds no source (README.md:1 only)."""
import math

import numpy as np
import pytest

import oracle
from oracle import Oracle
from tests import npmodel


def test_expneg_accuracy_and_monotone():
    h = np.linspace(0, 29.999, 20001).astype(np.float32)
    e = np.array([oracle.expneg(float(v)) for v in h], np.float64)
    ref = np.exp(-h.astype(np.float64))
    rel = np.abs(e - ref) / ref
    assert rel.max() < 2.5e-7          # about 2 ulp of f32
    assert np.all(np.diff(e) <= 0)     # monotone on the sampled grid
    assert oracle.expneg(0.0) == 1.0


def test_expneg_integer_boundaries():
    # half-integer multiples of ln2 exercise the round-to-even split of n
    for k in range(0, 43):
        for d in (-1e-3, 0.0, 1e-3):
            h = np.float32((k + 0.5) * math.log(2) + d)
            assert abs(oracle.expneg(float(h)) / math.exp(-float(h)) - 1) < 2.5e-7


def test_single_cell_closed_form():
    # dyadic coordinates: the res * 2^-22 quantisation of SPEC 3 is exact, so the hand values are exact
    o = Oracle([1.0], min_points=3)
    o.set_grid(0.0, 0.0, 4.0, 4.0)
    pts = np.array([[1.125, 2.125], [1.375, 2.125], [1.25, 2.5]], np.float32)
    o.set_target(pts)
    g = o.geometry()
    assert (g["nhx"], g["nhy"], g["njx"], g["njy"]) == (4, 4, 4, 4)
    cells = o.cells()
    rec = cells[2, 1]
    # mean (1.25, 2.25) = cell centre (1.5, 2.5) + (-0.25, -0.25); cov = diag(1/64, 3/64), xy = 0; B = diag(64, 64/3);
    # record = mean relative to the centre | B00 B01 | B01 B11 | n valid
    assert rec[0] == np.float32(-0.25) and rec[1] == np.float32(-0.25)
    assert rec[2] == np.float32(64.0) and rec[3] == 0.0 and rec[4] == 0.0 and rec[5] == np.float32(64.0 / 3.0)
    assert rec[6] == 3.0 and rec[7] == 1.0
    assert np.count_nonzero(cells[..., 7]) == 1
    n, s = o.sums()
    assert n[2, 1] == 3 and n.sum() == 3
    q = 1 << 22
    # offsets from the cell centre (1.5, 2.5) in res * 2^-22 units (res = 1 m)
    dx = np.array([-0.375, -0.125, -0.25]) * q
    dy = np.array([-0.375, -0.375, 0.0]) * q
    assert list(s[2, 1]) == [dx.sum(), dy.sum(), (dx * dx).sum(), (dx * dy).sum(), (dy * dy).sum()]


def test_min_points_and_degenerate_cells():
    o = Oracle([1.0], min_points=3, eig_ratio=0.001)
    o.set_grid(0.0, 0.0, 4.0, 4.0)
    pts = np.array([[0.5, 0.5], [0.6, 0.6],                      # two points: below min_points
                    [2.5, 2.5], [2.5, 2.5], [2.5, 2.5],          # identical: l1 = 0 -> invalid
                    [1.25, 0.25], [1.5, 0.25], [1.75, 0.25]],    # collinear: regularised
                   np.float32)
    o.set_target(pts)
    c = o.cells()
    assert c[0, 0, 7] == 0 and c[2, 2, 7] == 0 and c[0, 1, 7] == 1
    B = np.array([[c[0, 1, 2], c[0, 1, 3]], [c[0, 1, 4], c[0, 1, 5]]], np.float64)
    cov = np.linalg.inv(B)
    w = np.linalg.eigvalsh(cov)
    assert w[1] == pytest.approx(0.0625, rel=1e-6)        # var of {-.25, 0, .25} with n-1
    assert w[0] / w[1] == pytest.approx(0.001, rel=1e-5)  # raised to eig_ratio * l1


def test_eigen_regularisation_matches_numpy():
    rng = np.random.default_rng(3)
    o = Oracle([2.0], min_points=3, eig_ratio=0.01)
    o.set_grid(-4.0, -4.0, 8.0, 8.0)
    # thin, rotated blobs in several cells
    pts = []
    for cx, cy, ang in [(-3, -3, 0.3), (1, 1, 1.2), (3, -1, 2.5), (-1, 3, -0.7)]:
        d = rng.normal(size=(40, 2)) * np.array([0.3, 0.002])
        R = npmodel.rot(ang)
        pts.append(d @ R.T + np.array([cx, cy]))
    pts = np.concatenate(pts).astype(np.float32)
    o.set_target(pts)
    cells = o.cells()
    g = o.geometry()
    idx = o.cell_index(pts)
    for ci in np.unique(idx):
        p = pts[idx == ci].astype(np.float64)
        p = np.round(p * 2**20) / 2**20 if False else p
        cov = np.cov(p.T)
        w, V = np.linalg.eigh(cov)
        w[0] = max(w[0], 0.01 * w[1])
        cov_r = (V * w) @ V.T
        rec = cells[ci // g["nhx"], ci % g["nhx"]]
        B = np.array([[rec[2], rec[3]], [rec[4], rec[5]]], np.float64)
        assert rec[3] == rec[4]
        assert np.allclose(B, np.linalg.inv(cov_r), rtol=2e-4)
        assert np.allclose(npmodel.cell_centre(g, ci % g["nhx"], ci // g["nhx"]) + rec[:2], p.mean(0), atol=2e-7)
        assert 1.0 / np.linalg.det(B) == pytest.approx(np.linalg.det(cov_r), rel=2e-4)


def test_cell_index_matches_numpy_f64_on_edges():
    o = Oracle([0.25])
    o.set_grid(-100.0, -100.0, 200.0, 200.0)
    o.set_target(np.zeros((3, 2), np.float32))
    g = o.geometry()
    rng = np.random.default_rng(0)
    k = rng.integers(-5, 805, size=(20000, 2))
    base = (np.float32(-100.0) + k.astype(np.float32) * np.float32(0.25)).astype(np.float32)
    ulps = rng.integers(-3, 4, size=base.shape)
    pts = base.copy()
    for _ in range(3):
        pts = np.where(ulps > 0, np.nextafter(pts, np.float32(np.inf)), pts); ulps = ulps - (ulps > 0)
        pts = np.where(ulps < 0, np.nextafter(pts, np.float32(-np.inf)), pts); ulps = ulps + (ulps < 0)
    pts = np.concatenate([pts, np.array([[np.nan, 0], [0, np.inf], [-np.inf, 1], [99.999, 99.999], [100.0, 0.0],
                                         [-100.0, -100.0]], np.float32)])
    idx = o.cell_index(pts)
    # SPEC 2 (v4): f = ((double)X - (double)origin) * (1.0 / (double)st), h = floor(f): exact for these dyadic lattices
    p64 = pts.astype(np.float64)
    with np.errstate(invalid="ignore"):
        fx = (p64[:, 0] - float(g["ox"])) * (1.0 / float(g["st"]))
        fy = (p64[:, 1] - float(g["oy"])) * (1.0 / float(g["st"]))
        inside = (fx >= 0) & (fx < g["nhx"]) & (fy >= 0) & (fy < g["nhy"])
        exp = np.where(inside, np.floor(np.where(inside, fy, 0)).astype(np.int64) * g["nhx"] + np.floor(np.where(inside, fx, 0)).astype(np.int64), -1)
    assert np.array_equal(idx, exp)
    assert idx[-1] == 0 and idx[-2] == -1 and idx[-6] == -1


def test_transformed_cell_index_matches_numpy_f64(small_world):
    o = Oracle([0.5])
    o.set_target(small_world["map_xy"])
    g = o.geometry()
    xy = small_world["scans"][0]
    pose = small_world["init"][0]
    # SPEC 4 (v4): the transform runs in f64 in cell units; plain f64 numpy agrees except within ~1e-12 cells of an edge
    sn, cs = oracle.sincos(pose[2])
    inv = 1.0 / float(g["st"])
    x, y = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
    fx = cs * inv * x - sn * inv * y + (pose[0] - float(g["ox"])) * inv
    fy = sn * inv * x + cs * inv * y + (pose[1] - float(g["oy"])) * inv
    inside = (fx >= 0) & (fx < g["nhx"]) & (fy >= 0) & (fy < g["nhy"])
    exp = np.where(inside, np.floor(fy).astype(np.int64) * g["nhx"] + np.floor(fx).astype(np.int64), -1)
    safe = np.minimum(np.abs(fx - np.round(fx)), np.abs(fy - np.round(fy))) > 1e-9
    got = o.cell_index(xy, pose)
    assert safe.mean() > 0.99 and np.array_equal(got[safe], exp[safe])


@pytest.mark.parametrize("overlap", [0, 1])
def test_evaluate_matches_numpy_f64_model(small_world, overlap):
    o = Oracle([0.5], overlap=overlap)
    o.set_target(small_world["map_xy"])
    cells, geom = o.cells(), o.geometry()
    for i in (0, 5, 11):
        xy, pose = small_world["scans"][i], small_world["init"][i]
        pts, mu, B, edge = npmodel.gather(cells, geom, xy, pose, overlap)
        S, g, H = npmodel.score_terms(pts, pose, mu, B)
        out, cnt = o.evaluate(xy, pose)
        assert cnt > 200 * (4 if overlap else 1)
        # SPEC v4: the point-to-cell geometry is f64 and the f32 algebra works on cell-local coordinates, so the
        # evaluation agrees with exact f64 arithmetic on the same table to about 1e-7 of the natural scale
        # (v3, with f32 map-frame coordinates, needed 2e-4 on the score and 2e-3 on g and H here)
        assert out[0] == pytest.approx(S, rel=1e-6)
        Hn = np.array([[out[4], out[5], out[6]], [out[5], out[7], out[8]], [out[6], out[8], out[9]]])
        assert np.allclose(out[1:4], g, atol=1e-5 * np.abs(g).max())
        assert np.allclose(Hn, H, atol=2e-5 * np.abs(H).max())


def test_derivatives_match_finite_differences(small_world):
    # the analytic g/H of the numpy model against central differences of its own S (fixed assignment),
    # then the oracle against the model in test_evaluate_matches_numpy_f64_model: closes the chain
    o = Oracle([1.0])
    o.set_target(small_world["map_xy"])
    cells, geom = o.cells(), o.geometry()
    xy, pose = small_world["scans"][3], small_world["init"][3]
    pts, mu, B, _ = npmodel.gather(cells, geom, xy, pose)
    S, g, H = npmodel.score_terms(pts, pose, mu, B)
    f = lambda p: -npmodel.score_terms(pts, p, mu, B)[0]
    gf = lambda p: npmodel.score_terms(pts, p, mu, B)[1]
    eps = 1e-6
    for k in range(3):
        d = np.zeros(3); d[k] = eps
        assert (f(pose + d) - f(pose - d)) / (2 * eps) == pytest.approx(g[k], rel=1e-5, abs=1e-6)
        assert np.allclose((gf(pose + d) - gf(pose - d)) / (2 * eps), H[k], rtol=1e-5, atol=1e-4)


def f32_fma(a, b, c):
    """exact fused multiply-add on float32 values, rounded once (fractions keep it exact)"""
    from fractions import Fraction
    v = Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c))
    d = np.float64(float(v))                   # correctly rounded to f64 ...
    r = np.float32(d)                          # ... and again to f32: fix the rare double rounding by checking neighbours
    best = min((np.nextafter(r, np.float32(-np.inf)), r, np.nextafter(r, np.float32(np.inf))),
               key=lambda x: (abs(Fraction(float(x)) - v), int(np.float32(x).view(np.uint32)) & 1))
    return np.float32(best)


def test_point_terms_sum_to_evaluate(small_world):
    """SPEC 4 summation restated in Python from the per-pair factors: 64 f32 partials with fma, then the f64 butterfly."""
    o = Oracle([0.5], overlap=1)
    o.set_target(small_world["map_xy"])
    xy, pose = small_world["scans"][1][:400], small_world["init"][1]
    T = o.point_terms(xy, pose)
    out, cnt = o.evaluate(xy, pose)
    assert T.shape == (len(xy), 4, 10)
    part = np.zeros((64, 10), np.float32)
    for i in range(len(xy)):
        for k in range(4):
            if T[i, k, 0] != 0:
                acc = part[i & 63]
                acc[0] = acc[0] + T[i, k, 0]
                for t in range(1, 10):
                    acc[t] = f32_fma(T[i, k, 0], T[i, k, t], acc[t])
    D = part[:32].astype(np.float64) + part[32:].astype(np.float64)
    for o_ in (16, 8, 4, 2, 1):
        D = D + D[np.arange(32) ^ o_]
    assert np.array_equal(D[0], out)
    exact = (T[..., :1].astype(np.float64) * np.concatenate([np.ones_like(T[..., :1]), T[..., 1:]], -1).astype(np.float64)).sum((0, 1))
    assert np.allclose(exact, out, rtol=2e-6, atol=2e-6 * np.abs(out).max())
    assert np.count_nonzero(T[..., 0]) == cnt


def test_unusable_points_are_ignored(small_world):
    """SPEC 4: NaN / Inf / huge points are replaced by a far-away point and contribute nothing."""
    o = Oracle([0.5])
    o.set_target(small_world["map_xy"])
    xy, pose = small_world["scans"][0], small_world["init"][0]
    bad = np.array([[np.nan, 1.0], [1.0, np.inf], [-np.inf, np.nan], [1e30, 0.0], [3e38, -3e38]], np.float32)
    mixed = np.concatenate([xy[:500], bad, xy[500:]])
    a, ca = o.evaluate(xy, pose)
    # the bad points shift which partial later points fall into: compare with those points moved to the same slots
    far = np.tile(np.array([[1e18, 1e18]], np.float32), (5, 1))
    b, cb = o.evaluate(mixed, pose)
    c, cc = o.evaluate(np.concatenate([xy[:500], far, xy[500:]]), pose)
    assert ca == cb == cc and np.array_equal(b, c) and np.all(np.isfinite(b))
    assert np.allclose(a, b, rtol=1e-5, atol=1e-5 * np.abs(a).max())
    assert np.all(o.cell_index(bad, pose) == -1)


def test_overlap_even_entries_equal_single_grid():
    rng = np.random.default_rng(5)
    pts = rng.uniform(-9.5, 9.5, size=(4000, 2)).astype(np.float32)
    a = Oracle([1.0], overlap=0); a.set_grid(-10.0, -10.0, 20.0, 20.0); a.set_target(pts)
    b = Oracle([1.0], overlap=1); b.set_grid(-10.0, -10.0, 20.0, 20.0); b.set_target(pts)
    ga, gb = a.geometry(), b.geometry()
    assert (ga["nhx"], gb["nhx"], gb["njx"]) == (20, 40, 41)
    na, sa = a.sums(); nb, sb = b.sums()
    # entries with odd jx, jy are the unshifted grid (jx - ov even)
    assert np.array_equal(nb[1::2, 1::2], na) and np.array_equal(sb[1::2, 1::2], sa)
    assert np.array_equal(b.cells()[1::2, 1::2], a.cells())
    assert nb.sum() == 4 * len(pts)


def test_incremental_add_equals_union_build(small_world):
    m = small_world["map_xy"]
    a = Oracle([2.0, 0.5]); a.set_grid(-100.0, -100.0, 200.0, 200.0); a.set_target(m)
    b = Oracle([2.0, 0.5]); b.set_grid(-100.0, -100.0, 200.0, 200.0)
    perm = np.random.default_rng(1).permutation(len(m))
    b.set_target(m[perm[: len(m) // 3]])
    b.add_target(m[perm[len(m) // 3:]])
    for lv in (0, 1):
        assert np.array_equal(a.cells(lv), b.cells(lv))
        assert np.array_equal(a.sums(lv)[1], b.sums(lv)[1])


def test_solve_matches_numpy():
    rng = np.random.default_rng(2)
    for _ in range(50):
        M = rng.normal(size=(3, 3)); H = M @ M.T + 0.1 * np.eye(3); g = rng.normal(size=3)
        H6 = [H[0, 0], H[0, 1], H[0, 2], H[1, 1], H[1, 2], H[2, 2]]
        lam = 10 ** rng.uniform(-6, 1)
        ok, d = oracle.solve(g, H6, lam)
        A = H + lam * np.diag(np.maximum(np.abs(np.diag(H)), 1e-9))
        assert ok and np.allclose(d, np.linalg.solve(A, -g), rtol=1e-9)
    ok, _ = oracle.solve([1, 1, 1], [-1, 0, 0, 1, 0, 1], 0.0)
    assert not ok
    ok, _ = oracle.solve([1, 1, 1], [float("nan"), 0, 0, 1, 0, 1], 0.0)
    assert not ok


def test_sincos_spec_is_accurate_and_exact_at_special_angles():
    """SPEC 4.2: the specified f64 sin/cos sequence is within 1 ulp(1) of libm on the pose range, rounds to the same f32
    as libm on a dense sample, and is exact where it has to be (0, NaN, infinities, quadrant boundaries' signs)."""
    rng = np.random.default_rng(9)
    th = np.concatenate([rng.uniform(-10, 10, 20000), rng.uniform(-1e4, 1e4, 5000), np.arange(-40, 41) * math.pi / 4])
    worst, bad32 = 0.0, 0
    for t in th:
        s, c = oracle.sincos(float(t))
        worst = max(worst, abs(s - math.sin(t)), abs(c - math.cos(t)))
        bad32 += np.float32(s) != np.float32(math.sin(t)) or np.float32(c) != np.float32(math.cos(t))
    assert worst <= 2.3e-16 and bad32 == 0
    assert oracle.sincos(0.0) == (0.0, 1.0)
    s, c = oracle.sincos(math.pi / 2)
    assert s == 1.0 and abs(c) < 1e-16
    s, c = oracle.sincos(-math.pi)
    assert c == -1.0 and abs(s) < 2e-16
    for t in (float("nan"), float("inf"), -float("inf")):
        s, c = oracle.sincos(t)
        assert math.isnan(s) and math.isnan(c)


def test_align_identity_and_recovery():
    from gtsam_ndt_b200 import synth
    r, p = synth.scans(1, traj_len=1000, first=17, **synth.SCAN_360)
    scan = synth.polar_to_points(r[0], synth.SCAN_360["angle_min"], synth.SCAN_360["angle_inc"])
    o = Oracle([0.5])
    o.set_target(scan)
    res = o.align(scan, [0, 0, 0])
    assert res["status"] == 0 and np.allclose(res["pose"], 0, atol=2e-3)
    # scan-to-scan offset (0.10 m, -0.05 m, 2 deg)
    T = np.array([0.10, -0.05, math.radians(2.0)])
    moved = synth.transform(scan, -np.array([*(npmodel.rot(-T[2]) @ T[:2]), T[2]]))  # inverse transform
    o2 = Oracle([2.0, 1.0, 0.5])
    o2.set_target(scan)
    res = o2.align(moved, [0, 0, 0])
    assert res["status"] == 0
    assert np.allclose(res["pose"][:2], T[:2], atol=0.02) and abs(res["pose"][2] - T[2]) < 5e-3
    assert np.all(np.linalg.eigvalsh(res["hessian"]) > 0)


def test_align_scan_to_map_recovers_truth(small_world):
    o = Oracle([1.0, 0.5])
    o.set_target(small_world["map_xy"])
    offsets_xy, offsets = __import__("gtsam_ndt_b200.synth", fromlist=["pack"]).pack(small_world["scans"])
    res = o.align_batch(offsets_xy, offsets, small_world["init"])
    err = res["pose"] - small_world["poses"]
    err[:, 2] = (err[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.all(res["status"] == 0)
    assert np.abs(err[:, :2]).max() < 0.03 and np.abs(err[:, 2]).max() < 3e-3
    one = o.align(small_world["scans"][4], small_world["init"][4])
    assert one.tobytes() == res[4].tobytes()      # batch == single, bit for bit
    assert res["iterations"].max() <= 60


def test_align_empty_and_no_overlap(small_world):
    o = Oracle([0.5])
    o.set_target(small_world["map_xy"])
    r = o.align(np.zeros((0, 2), np.float32), [1.0, 2.0, 0.3])
    assert r["status"] == 3 and r["iterations"] == 1 and list(r["pose"]) == [1.0, 2.0, 0.3]
    far = small_world["scans"][0] + np.float32(5000.0)
    r = o.align(far, [0, 0, 0])
    assert r["status"] == 3 and r["count"] == 0 and r["score"] == 0


def test_sweep_best_and_ties(small_world):
    o = Oracle([1.0])
    o.set_target(small_world["map_xy"])
    xy, truth = small_world["scans"][2], small_world["poses"][2]
    g = np.stack(np.meshgrid(np.arange(-2, 3) * 0.4, np.arange(-2, 3) * 0.4, np.radians(np.arange(-2, 3) * 2.0),
                             indexing="ij"), -1).reshape(-1, 3)
    hyp = (truth + g).astype(np.float32)
    hyp = np.concatenate([hyp, hyp])           # duplicates: the tie must go to the smaller index
    scores, bi, bs = o.sweep(xy, hyp)
    assert bi == int(np.argmax(scores)) < len(hyp) // 2 and bs == scores[bi]
    assert np.allclose(hyp[bi], truth, atol=1e-3)
    s1, _ = o.evaluate(xy, hyp[7].astype(np.float64))
    assert s1[0] == scores[7]
    _, bi1, _ = o.sweep(xy, hyp, nthreads=1)
    assert bi1 == bi


def test_polar_to_points_matches_numpy():
    from gtsam_ndt_b200 import synth
    r, _ = synth.scans(2, traj_len=500, max_range=30.0, **synth.SCAN_1080)
    for i in range(2):
        a = oracle.polar_to_points(r[i], synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"], range_min=0.05, range_max=30.0)
        b = synth.polar_to_points(r[i], synth.SCAN_1080["angle_min"], synth.SCAN_1080["angle_inc"], 0.05, 30.0)
        assert 300 < len(a) < 1080 and np.array_equal(a, b)
    u = np.round(r[0] * 1000).clip(0, 65535).astype(np.uint16)
    a = oracle.polar_to_points(u, -2.0, 0.004, range_scale=0.001, range_min=0.05, range_max=30.0)
    cb, sb = synth.beam_table(1080, -2.0, 0.004)
    rho = u.astype(np.float32) * np.float32(0.001)
    k = (u != 0) & (rho >= np.float32(0.05)) & (rho <= np.float32(30.0))
    assert np.array_equal(a, np.stack([rho[k] * cb[k], rho[k] * sb[k]], 1))

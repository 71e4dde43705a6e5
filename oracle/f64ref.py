"""TEST INFRASTRUCTURE (like everything under oracle/): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import it; the product never does.

An independent double-precision 2D NDT ("f64 twin") used to bound how far the product's arithmetic is from a
plain f64 implementation of the same algorithm (north_star tolerances: pose 1e-5 m / 1e-6 rad, score and Hessian
1e-6 relative, cell assignment exact).

It shares no code with oracle/ or the CUDA path and follows none of SPEC.md's bit-level choices: points are taken to
f64 as given, cells are floor((x - origin) / stride) in f64, statistics are plain f64 sums (no fixed point), exp is
numpy's, sums are numpy's pairwise sums, the 3x3 system is solved by numpy.linalg. What it does share with SPEC.md is
the algorithm: lattice geometry (SPEC 2), n-1 covariance with the eigenvalue floor (SPEC 3), the score and its
derivatives (SPEC 4), the Levenberg-Marquardt schedule and its parameters (SPEC 5).
PARITY UNPINNED: this is a stand-in for "the reference's CPU NDT in double precision", which does not exist in the mount.
"""
import numpy as np


class LevelF64:
    def __init__(self, res, ox, oy, nhx, nhy, overlap=0):
        self.res = float(res)
        self.ov = int(overlap)
        self.st = self.res * 0.5 if self.ov else self.res
        self.ox, self.oy = float(ox), float(oy)
        self.nhx, self.nhy = int(nhx), int(nhy)
        self.njx, self.njy = self.nhx + self.ov, self.nhy + self.ov
        nc = self.njx * self.njy
        self.valid = np.zeros(nc, bool)
        self.mu = np.zeros((nc, 2))
        self.B = np.zeros((nc, 2, 2))

    def lattice(self, X, Y):
        fx = (X - self.ox) / self.st
        fy = (Y - self.oy) / self.st
        inside = (fx >= 0) & (fx < self.nhx) & (fy >= 0) & (fy < self.nhy)
        hx = np.floor(np.where(inside, fx, 0)).astype(np.int64)
        hy = np.floor(np.where(inside, fy, 0)).astype(np.int64)
        return inside, hx, hy

    def build(self, xy, min_points=3, eig_ratio=0.01):
        xy = np.asarray(xy, np.float64)
        ok = np.isfinite(xy).all(1)
        xy = xy[ok]
        inside, hx, hy = self.lattice(xy[:, 0], xy[:, 1])
        xy, hx, hy = xy[inside], hx[inside], hy[inside]
        nc = self.njx * self.njy
        K = 2 if self.ov else 1
        n = np.zeros(nc)
        s = np.zeros((nc, 5))
        for b in range(K):
            for a in range(K):
                jx, jy = hx + a, hy + b
                c = jy * self.njx + jx
                # sums about the cell centre (keeps the f64 sums well conditioned; an implementation detail)
                cx = self.ox + (jx - self.ov) * self.st + 0.5 * self.res
                cy = self.oy + (jy - self.ov) * self.st + 0.5 * self.res
                dx, dy = xy[:, 0] - cx, xy[:, 1] - cy
                n += np.bincount(c, minlength=nc)
                for k, w in enumerate((dx, dy, dx * dx, dx * dy, dy * dy)):
                    s[:, k] += np.bincount(c, weights=w, minlength=nc)
        idx = np.nonzero(n >= min_points)[0]
        N = n[idx]
        mx, my = s[idx, 0] / N, s[idx, 1] / N
        cxx = (s[idx, 2] - s[idx, 0] * mx) / (N - 1)
        cxy = (s[idx, 3] - s[idx, 0] * my) / (N - 1)
        cyy = (s[idx, 4] - s[idx, 1] * my) / (N - 1)
        tr, hd = cxx + cyy, 0.5 * (cxx - cyy)
        rad = np.sqrt(hd * hd + cxy * cxy)
        l1, l2 = 0.5 * tr + rad, 0.5 * tr - rad
        good = l1 > 1e-10
        fl = good & (l2 < eig_ratio * l1)
        l2n = eig_ratio * l1
        vx = np.where(hd >= 0, hd + rad, cxy)
        vy = np.where(hd >= 0, cxy, rad - hd)
        nn = vx * vx + vy * vy
        nn = np.where(nn > 0, nn, 1.0)
        dl = l1 - l2n
        cxx = np.where(fl, l2n + dl * vx * vx / nn, cxx)
        cxy = np.where(fl, dl * vx * vy / nn, cxy)
        cyy = np.where(fl, l2n + dl * vy * vy / nn, cyy)
        det = cxx * cyy - cxy * cxy
        jx, jy = idx % self.njx, idx // self.njx
        cx = self.ox + (jx - self.ov) * self.st + 0.5 * self.res
        cy = self.oy + (jy - self.ov) * self.st + 0.5 * self.res
        idx, sel = idx[good], good
        self.valid[:] = False
        self.valid[idx] = True
        self.mu[idx, 0] = (cx + mx)[sel]
        self.mu[idx, 1] = (cy + my)[sel]
        self.B[idx, 0, 0] = (cyy / det)[sel]
        self.B[idx, 0, 1] = self.B[idx, 1, 0] = (-cxy / det)[sel]
        self.B[idx, 1, 1] = (cxx / det)[sel]

    def evaluate(self, xy, pose):
        """S, g(3), H(3,3) of f = -S, count; f64 throughout."""
        xy = np.asarray(xy, np.float64)
        xy = xy[np.isfinite(xy).all(1)]
        c, s = np.cos(pose[2]), np.sin(pose[2])
        rx, ry = c * xy[:, 0] - s * xy[:, 1], s * xy[:, 0] + c * xy[:, 1]
        X, Y = rx + pose[0], ry + pose[1]
        inside, hx, hy = self.lattice(X, Y)
        K = 2 if self.ov else 1
        S, g, H, cnt = 0.0, np.zeros(3), np.zeros((3, 3)), 0
        for b in range(K):
            for a in range(K):
                cidx = (hy + b) * self.njx + (hx + a)
                ok = inside & self.valid[np.where(inside, cidx, 0)]
                ci = cidx[ok]
                q = np.stack([X[ok], Y[ok]], 1) - self.mu[ci]
                Bc = self.B[ci]
                u = np.einsum("nij,nj->ni", Bc, q)
                h = 0.5 * np.einsum("ni,ni->n", q, u)
                keep = h < 30.0
                q, u, Bc, h = q[keep], u[keep], Bc[keep], h[keep]
                r = np.stack([rx[ok][keep], ry[ok][keep]], 1)
                e = np.exp(-h)
                n = len(e)
                J = np.zeros((n, 2, 3))
                J[:, 0, 0] = 1.0
                J[:, 1, 1] = 1.0
                J[:, 0, 2] = -r[:, 1]
                J[:, 1, 2] = r[:, 0]
                av = np.einsum("ni,nik->nk", u, J)
                Hn = -np.einsum("nk,nl->nkl", av, av) + np.einsum("nik,nij,njl->nkl", J, Bc, J)
                Hn[:, 2, 2] += -(u * r).sum(1)
                S += e.sum()
                g += (e[:, None] * av).sum(0)
                H += (e[:, None, None] * Hn).sum(0)
                cnt += n
        return S, g, H, cnt


DEFAULTS = dict(min_points=3, eig_ratio=0.01, max_iterations=30, eps_trans=1e-4, eps_rot=1e-5, max_step_trans=0.5,
                max_step_rot=0.2, lambda_init=1e-3, lambda_min=1e-9, lambda_max=1e7, lambda_up=10.0, lambda_down=5.0,
                lambda_fail_up=3.0)


class NdtF64:
    """Pyramid of LevelF64 with the SPEC 5 Levenberg-Marquardt loop in plain f64 numpy."""

    def __init__(self, geoms, overlap=0, **params):
        """geoms: list of dicts with res, ox, oy, nhx, nhy (e.g. matcher.geometry(l))."""
        self.P = dict(DEFAULTS)
        self.P.update(params)
        self.levels = [LevelF64(float(g["res"]), float(g["ox"]), float(g["oy"]), int(g["nhx"]), int(g["nhy"]), overlap) for g in geoms]

    def set_target(self, xy):
        for L in self.levels:
            L.build(xy, self.P["min_points"], self.P["eig_ratio"])

    def evaluate(self, xy, pose, level=0):
        return self.levels[level].evaluate(xy, pose)

    @staticmethod
    def _solve(g, H, lam):
        A = H.copy()
        for k in range(3):
            A[k, k] += lam * max(abs(H[k, k]), 1e-9)
        try:
            np.linalg.cholesky(A)
        except np.linalg.LinAlgError:
            return None
        return -np.linalg.solve(A, g)

    def align(self, xy, init):
        P = self.P
        p = np.array(init, np.float64)
        evals_total, status, E = 0, 3, None
        for L in self.levels:
            lam = P["lambda_init"]
            E = L.evaluate(xy, p)
            evals, status = 1, 1
            if len(xy) == 0 or E[3] == 0:
                evals_total += evals
                status = 3
                continue
            while evals < P["max_iterations"]:
                d = self._solve(E[1], E[2], lam)
                stalled = False
                while d is None:
                    lam *= P["lambda_fail_up"]
                    if lam > P["lambda_max"]:
                        stalled = True
                        break
                    d = self._solve(E[1], E[2], lam)
                if stalled:
                    status = 2
                    break
                n2 = d[0] * d[0] + d[1] * d[1]
                if n2 > P["max_step_trans"] ** 2:
                    d = d * (P["max_step_trans"] / np.sqrt(n2))
                    n2 = P["max_step_trans"] ** 2
                if abs(d[2]) > P["max_step_rot"]:
                    sc = P["max_step_rot"] / abs(d[2])
                    d = d * sc
                    n2 *= sc * sc
                small = n2 < P["eps_trans"] ** 2 and abs(d[2]) < P["eps_rot"]
                En = L.evaluate(xy, p + d)
                evals += 1
                if En[0] > E[0]:
                    p, E = p + d, En
                    lam = max(lam / P["lambda_down"], P["lambda_min"])
                    if small:
                        status = 0
                        break
                else:
                    if small:
                        status = 0
                        break
                    lam *= P["lambda_up"]
                    if lam > P["lambda_max"]:
                        status = 2
                        break
            evals_total += evals
        p[2] = p[2] - 2 * np.pi * np.rint(p[2] / (2 * np.pi))
        return dict(pose=p, score=E[0], grad=E[1], hessian=E[2], iterations=evals_total, status=status, count=E[3])


def compare(impl, twin, scans, poses, inits, level=0):
    """Distance between an implementation (oracle.Oracle or NdtMatcher2D: evaluate(xy, pose, level) -> (out10, count) and
    align(xy, init)) and the f64 twin: max relative error of S, g, H at `poses` and final-pose differences from `inits`."""
    out = dict(score_rel=0.0, grad_rel=0.0, hess_rel=0.0, count_mismatch=0, dpos=[], drot=[], iters_equal=0, n=0)
    for xy, p in zip(scans, poses):
        o, cnt = impl.evaluate(xy, p, level)
        S, g, H, c2 = twin.evaluate(xy, p, level)
        Ho = np.array([[o[4], o[5], o[6]], [o[5], o[7], o[8]], [o[6], o[8], o[9]]])
        out["score_rel"] = max(out["score_rel"], abs(o[0] - S) / max(abs(S), 1e-300))
        out["grad_rel"] = max(out["grad_rel"], np.abs(o[1:4] - g).max() / max(np.abs(g).max(), 1e-300))
        out["hess_rel"] = max(out["hess_rel"], np.abs(Ho - H).max() / max(np.abs(H).max(), 1e-300))
        out["count_mismatch"] += int(cnt != c2)
    for xy, p0 in zip(scans, inits):
        r = impl.align(xy, p0)
        t = twin.align(xy, p0)
        d = r["pose"] - t["pose"]
        out["dpos"].append(float(np.hypot(d[0], d[1])))
        out["drot"].append(float(abs((d[2] + np.pi) % (2 * np.pi) - np.pi)))
        out["iters_equal"] += int(r["iterations"] == t["iterations"])
        out["n"] += 1
    return out

"""Independent numpy f64 model of the NDT score (SPEC.md section 4), used to pin the spec oracle.

It shares no code with oracle/ or the CUDA path: plain matrix formulas, exact exp, f64 throughout.
The cell table and the point-to-cell assignment are inputs (taken from the implementation under
test), so what this checks is the per-point algebra, the derivatives and the sums.
"""
import numpy as np


def rot(th):
    c, s = np.cos(th), np.sin(th)
    return np.array([[c, -s], [s, c]])


def score_terms(xy, pose, mu, B):
    """xy (n,2), mu (n,2), B (n,2,2) per-point assigned cell. Returns S, g(3), H(3,3) of f = -S."""
    xy = np.asarray(xy, np.float64)
    R = rot(pose[2])
    r = xy @ R.T
    X = r + np.asarray(pose[:2])
    q = X - mu
    u = np.einsum("nij,nj->ni", B, q)
    m = np.einsum("ni,ni->n", q, u)
    e = np.exp(-0.5 * m)
    n = len(xy)
    J = np.zeros((n, 2, 3))
    J[:, 0, 0] = 1.0
    J[:, 1, 1] = 1.0
    J[:, 0, 2] = -r[:, 1]
    J[:, 1, 2] = r[:, 0]
    a = np.einsum("ni,nik->nk", u, J)
    S = e.sum()
    g = (e[:, None] * a).sum(0)
    JBJ = np.einsum("nik,nij,njl->nkl", J, B, J)
    H = -np.einsum("nk,nl->nkl", a, a) + JBJ
    H[:, 2, 2] += -(u * r).sum(1)
    H = (e[:, None, None] * H).sum(0)
    return S, g, H


def cell_centre(geom, jx, jy, overlap=0):
    """SPEC 2: centre of table entry (jx, jy) in the map frame, f64."""
    ov = 1 if overlap else 0
    st, res = float(geom["st"]), float(geom["res"])
    return np.array([float(geom["ox"]) + (jx - ov) * st + 0.5 * res, float(geom["oy"]) + (jy - ov) * st + 0.5 * res])


def gather(cells, geom, xy, pose, overlap=0):
    """Per-(point, cell) pairs of a pose: returns (xy_rep, mu, B) for valid pairs; mu in the map frame (the records
    carry the mean relative to the cell centre since SPEC v4; the centre is added back here in f64)."""
    xy = np.asarray(xy, np.float32)
    x64 = xy.astype(np.float64)
    # the lattice index is plain f64 math here (SPEC 2 is f64 as well; points closer than 1e-9 cells to an edge could differ)
    X = np.cos(pose[2]) * x64[:, 0] - np.sin(pose[2]) * x64[:, 1] + pose[0]
    Y = np.sin(pose[2]) * x64[:, 0] + np.cos(pose[2]) * x64[:, 1] + pose[1]
    fx = (X - float(geom["ox"])) / float(geom["st"])
    fy = (Y - float(geom["oy"])) / float(geom["st"])
    inside = (fx >= 0) & (fx < geom["nhx"]) & (fy >= 0) & (fy < geom["nhy"])
    hx = np.floor(fx).astype(int); hy = np.floor(fy).astype(int)
    edge = np.minimum(np.abs(fx - np.round(fx)), np.abs(fy - np.round(fy)))
    K = 2 if overlap else 1
    pts, mus, Bs = [], [], []
    for i in np.nonzero(inside)[0]:
        for b in range(K):
            for a in range(K):
                rec = cells[hy[i] + b, hx[i] + a].astype(np.float64)
                if rec[7] == 0:
                    continue
                ctr = cell_centre(geom, hx[i] + a, hy[i] + b, overlap)
                pts.append(x64[i]); mus.append(ctr + rec[0:2]); Bs.append([[rec[2], rec[3]], [rec[4], rec[5]]])
    if not pts:
        return np.zeros((0, 2)), np.zeros((0, 2)), np.zeros((0, 2, 2)), edge
    return np.array(pts), np.array(mus), np.array(Bs), edge

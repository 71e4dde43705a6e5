"""Synthetic 2D laser world: ctypes front end of csrc/synth.c plus numpy helpers.

Host-only helper shared by tests and bench.py. It does not touch the GPU or the oracle.
"""
import ctypes as C
import math
import os

import numpy as np

from . import build as _build

_lib = None

SCAN_1080 = dict(nbeams=1080, angle_min=-0.75 * math.pi, angle_inc=math.radians(0.25))  # 270 deg
SCAN_360 = dict(nbeams=360, angle_min=-math.pi, angle_inc=math.radians(1.0))            # 360 deg
WORLD_SEED = 12345       # noise / world seed
PERTURB_SEED = 67890     # pose perturbation seed
NBOXES = 400


def lib():
    global _lib
    if _lib is None:
        so = _build.LIB_SYNTH
        if not os.path.exists(so) or os.path.getmtime(os.path.join(_build.CSRC, "synth.c")) > os.path.getmtime(so):
            _build.build_synth()
        _lib = C.CDLL(so)
        _lib.synth_world_boxes.restype = C.c_int
    return _lib


def scans(nscans, traj_len=None, first=0, step=1, nbeams=1080, angle_min=-0.75 * math.pi,
          angle_inc=math.radians(0.25), max_range=0.0, sigma=0.01, seed=WORLD_SEED, noise_seed=WORLD_SEED, nboxes=NBOXES):
    """ranges[nscans, nbeams] f32 (0 = no return) and true poses[nscans, 3] f64."""
    traj_len = traj_len or nscans * step
    ranges = np.zeros((nscans, nbeams), np.float32)
    poses = np.zeros((nscans, 3), np.float64)
    rc = lib().synth_scans(C.c_uint64(seed), C.c_uint64(noise_seed), C.c_int(nboxes), C.c_int64(traj_len), C.c_int64(first), C.c_int64(step),
                           C.c_int(nscans), C.c_int(nbeams), C.c_double(angle_min), C.c_double(angle_inc),
                           C.c_double(max_range), C.c_double(sigma),
                           ranges.ctypes.data_as(C.POINTER(C.c_float)), poses.ctypes.data_as(C.POINTER(C.c_double)))
    if rc != 0:
        raise RuntimeError("synth_scans failed")
    return ranges, poses


def uniform3(n, first=0, seed=PERTURB_SEED):
    out = np.zeros((n, 3), np.float64)
    lib().synth_uniform3(C.c_uint64(seed), C.c_int64(first), C.c_int(n), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def boxes(seed=WORLD_SEED, nboxes=NBOXES):
    out = np.zeros((32768, 4), np.float64)
    n = lib().synth_world_boxes(C.c_uint64(seed), C.c_int(nboxes), out.ctypes.data_as(C.POINTER(C.c_double)), 32768)
    return out[:n]


DENSE_NBOXES = 30000


def set_threads(n):
    """OpenMP threads of the generator (torchrun sets OMP_NUM_THREADS=1 for every rank; a bench rank takes its share of the cores)."""
    lib().synth_set_threads(C.c_int(int(n)))


def set_world(dense=False):
    """Box shape of the synthetic world for every later call: the SURVEY 8(d) room (default), or the cluttered "dense"
    world (pass nboxes=DENSE_NBOXES to scans / boxes): ~17 000 boxes of 0.6-2 m, 1.5 m clear of the trajectory."""
    if dense:
        lib().synth_set_world(C.c_double(0.3), C.c_double(0.7), C.c_double(1.5))
    else:
        lib().synth_set_world(C.c_double(1.0), C.c_double(3.0), C.c_double(3.0))


def dense_map(seed=WORLD_SEED, nboxes=DENSE_NBOXES, spacing=0.04, sigma=0.01):
    """Map cloud of the dense world: every box edge and the outer walls sampled every `spacing` m with Gaussian noise, as a
    surveyed map would hold them (a lidar on the loop alone sees a few percent of the boxes). Call set_world(dense=True) first."""
    bx = boxes(seed, nboxes)
    rng = np.random.default_rng(seed)
    segs = []
    for x0, y0, x1, y1 in np.concatenate([bx, [[-95.0, -95.0, 95.0, 95.0]]]):
        segs += [(x0, y0, x1, y0), (x1, y0, x1, y1), (x1, y1, x0, y1), (x0, y1, x0, y0)]
    segs = np.array(segs)
    ln = np.hypot(segs[:, 2] - segs[:, 0], segs[:, 3] - segs[:, 1])
    cnt = np.maximum(2, np.ceil(ln / spacing).astype(np.int64))
    t = np.concatenate([np.linspace(0.0, 1.0, c, endpoint=False) for c in cnt])
    si = np.repeat(np.arange(len(segs)), cnt)
    pts = segs[si, :2] + t[:, None] * (segs[si, 2:] - segs[si, :2])
    return (pts + rng.normal(size=pts.shape) * sigma).astype(np.float32)


def beam_table(nbeams, angle_min, angle_inc):
    """SPEC.md section 8: (cb_i, sb_i) = f32(cos/sin(angle_min + i*angle_inc)), phi in f64."""
    phi = angle_min + np.arange(nbeams, dtype=np.float64) * angle_inc
    return np.cos(phi).astype(np.float32), np.sin(phi).astype(np.float32)


def polar_to_points(ranges, angle_min, angle_inc, range_min=0.0, range_max=np.inf):
    """SPEC.md section 8 in numpy (f32 multiply): list of per-scan (n_i, 2) f32 arrays, or one for 1-D input."""
    r = np.asarray(ranges, np.float32)
    single = r.ndim == 1
    r = np.atleast_2d(r)
    cb, sb = beam_table(r.shape[1], angle_min, angle_inc)
    keep = (r >= np.float32(range_min)) & (r <= np.float32(range_max)) & (r != 0)
    out = []
    for i in range(r.shape[0]):
        k = keep[i]
        out.append(np.stack([r[i, k] * cb[k], r[i, k] * sb[k]], axis=1).astype(np.float32))
    return out[0] if single else out


def pack(scan_list):
    """Ragged list of (n_i, 2) arrays -> (xy[sum n, 2] f32, offsets[B+1] i64)."""
    offsets = np.zeros(len(scan_list) + 1, np.int64)
    offsets[1:] = np.cumsum([len(s) for s in scan_list])
    xy = np.concatenate(scan_list, axis=0).astype(np.float32) if len(scan_list) else np.zeros((0, 2), np.float32)
    return np.ascontiguousarray(xy), offsets


def transform(xy, pose):
    """World-frame points (f64 math, rounded to f32 once): used to assemble map point clouds."""
    c, s = math.cos(pose[2]), math.sin(pose[2])
    x = xy[:, 0].astype(np.float64)
    y = xy[:, 1].astype(np.float64)
    return np.stack([c * x - s * y + pose[0], s * x + c * y + pose[1]], axis=1).astype(np.float32)


def make_map(nscans, traj_len=None, nbeams=1080, angle_min=-0.75 * math.pi, angle_inc=math.radians(0.25),
             max_range=0.0, sigma=0.01, seed=WORLD_SEED):
    """Map point cloud: `nscans` noisy scans evenly spaced on the loop, placed at their true poses."""
    traj_len = traj_len or nscans
    step = max(1, traj_len // nscans)
    r, p = scans(nscans, traj_len=traj_len, first=0, step=step, nbeams=nbeams, angle_min=angle_min,
                 angle_inc=angle_inc, max_range=max_range, sigma=sigma, seed=seed, noise_seed=seed ^ 0x5EED)
    pts = polar_to_points(r, angle_min, angle_inc)
    return np.concatenate([transform(q, p[i]) for i, q in enumerate(pts)], axis=0)

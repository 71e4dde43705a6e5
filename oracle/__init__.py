"""ctypes loader for the SPEC ORACLE (oracle/ndt2d_oracle.c).

Test infrastructure only: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package never imports this module.
PARITY UNPINNED: the reference mount has no source (/root/reference/README.md:1), so this is a
restatement of SPEC.md, not of upstream GTSAM-NDT.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libndt2d_oracle.so")


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("ndt2d_oracle.c", "ndt2d_oracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _SO


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("eig_ratio", "eps_trans", "eps_rot", "max_step_trans", "max_step_rot",
                                          "lambda_init", "lambda_min", "lambda_max", "lambda_up", "lambda_down", "lambda_fail_up")] + \
               [(n, C.c_int32) for n in ("min_points", "max_iterations", "overlap", "reserved")]


class Result(C.Structure):
    _fields_ = [("pose", C.c_double * 3), ("score", C.c_double), ("grad", C.c_double * 3),
                ("hessian", C.c_double * 9), ("iterations", C.c_int32), ("status", C.c_int32),
                ("count", C.c_int32), ("reserved", C.c_int32)]


RESULT_DTYPE = np.dtype([("pose", "f8", 3), ("score", "f8"), ("grad", "f8", 3), ("hessian", "f8", (3, 3)),
                         ("iterations", "i4"), ("status", "i4"), ("count", "i4"), ("reserved", "i4")])
assert RESULT_DTYPE.itemsize == C.sizeof(Result) == 144

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_create.restype = C.c_void_p
        L.oracle_expneg.restype = C.c_float
        L.oracle_expneg.argtypes = [C.c_float]
        for name in ("oracle_destroy", "oracle_default_params"):
            getattr(L, name).restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Mirror of the matcher API (set resolutions / grid / target, evaluate, align, sweep) on the CPU."""

    def __init__(self, resolutions=(1.0,), **params):
        self.L = lib()
        self.h = C.c_void_p(self.L.oracle_create())
        self.params = Params()
        self.L.oracle_default_params(C.byref(self.params))
        self.set_params(**params)
        self.set_resolutions(resolutions)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.oracle_destroy(self.h)
            self.h = None

    def _ck(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"oracle {what} failed rc={rc}")

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        self._ck(self.L.oracle_set_params(self.h, C.byref(self.params)), "set_params")

    def set_resolutions(self, res):
        r = _f32(np.atleast_1d(res))
        self.nlevels = len(r)
        self._ck(self.L.oracle_set_resolutions(self.h, _p(r, C.c_float), len(r)), "set_resolutions")

    def set_grid(self, ox, oy, ex, ey):
        self._ck(self.L.oracle_set_grid(self.h, C.c_float(ox), C.c_float(oy), C.c_float(ex), C.c_float(ey)), "set_grid")

    def set_target(self, xy):
        xy = _f32(xy).reshape(-1, 2)
        self._ck(self.L.oracle_set_target(self.h, _p(xy, C.c_float), len(xy)), "set_target")

    def add_target(self, xy):
        xy = _f32(xy).reshape(-1, 2)
        self._ck(self.L.oracle_add_target(self.h, _p(xy, C.c_float), len(xy)), "add_target")

    def geometry(self, level=0):
        g = np.zeros(5, np.float32)
        d = np.zeros(4, np.int32)
        self._ck(self.L.oracle_level_geometry(self.h, level, _p(g, C.c_float), _p(d, C.c_int32)), "geometry")
        return dict(res=g[0], st=g[1], inv_st=g[2], ox=g[3], oy=g[4], nhx=int(d[0]), nhy=int(d[1]),
                    njx=int(d[2]), njy=int(d[3]))

    def cells(self, level=0):
        g = self.geometry(level)
        out = np.zeros((g["njy"], g["njx"], 8), np.float32)
        self._ck(self.L.oracle_get_cells(self.h, level, _p(out, C.c_float)), "get_cells")
        return out

    def sums(self, level=0):
        g = self.geometry(level)
        n = np.zeros((g["njy"], g["njx"]), np.uint32)
        s = np.zeros((g["njy"], g["njx"], 5), np.int64)
        self._ck(self.L.oracle_get_sums(self.h, level, _p(n, C.c_uint32), _p(s, C.c_int64)), "get_sums")
        return n, s

    def cell_index(self, xy, pose=None, level=0):
        xy = _f32(xy).reshape(-1, 2)
        idx = np.zeros(len(xy), np.int32)
        pp = None
        if pose is not None:
            pose = np.ascontiguousarray(pose, np.float64)
            pp = _p(pose, C.c_double)
        self._ck(self.L.oracle_cell_index(self.h, level, _p(xy, C.c_float), len(xy), pp, _p(idx, C.c_int32)), "cell_index")
        return idx

    def evaluate(self, xy, pose, level=0):
        xy = _f32(xy).reshape(-1, 2)
        pose = np.ascontiguousarray(pose, np.float64)
        out = np.zeros(10, np.float64)
        cnt = C.c_int32(0)
        self._ck(self.L.oracle_evaluate(self.h, level, _p(xy, C.c_float), len(xy), _p(pose, C.c_double),
                                        _p(out, C.c_double), C.byref(cnt)), "evaluate")
        return out, cnt.value

    def point_terms(self, xy, pose, level=0):
        xy = _f32(xy).reshape(-1, 2)
        pose = np.ascontiguousarray(pose, np.float64)
        K = 4 if self.params.overlap else 1
        out = np.zeros((len(xy), K, 10), np.float32)
        self._ck(self.L.oracle_point_terms(self.h, level, _p(xy, C.c_float), len(xy), _p(pose, C.c_double),
                                           _p(out, C.c_float)), "point_terms")
        return out

    def align(self, xy, init):
        xy = _f32(xy).reshape(-1, 2)
        init = np.ascontiguousarray(init, np.float64)
        r = np.zeros(1, RESULT_DTYPE)
        self._ck(self.L.oracle_align(self.h, _p(xy, C.c_float), len(xy), _p(init, C.c_double),
                                     r.ctypes.data_as(C.c_void_p)), "align")
        return r[0]

    def align_batch(self, xy, offsets, init, nthreads=0):
        xy = _f32(xy).reshape(-1, 2)
        offsets = np.ascontiguousarray(offsets, np.int64)
        init = np.ascontiguousarray(init, np.float64).reshape(-1, 3)
        nb = len(offsets) - 1
        assert len(init) == nb and offsets[-1] <= len(xy)
        r = np.zeros(nb, RESULT_DTYPE)
        self._ck(self.L.oracle_align_batch(self.h, _p(xy, C.c_float), _p(offsets, C.c_int64), nb,
                                           _p(init, C.c_double), r.ctypes.data_as(C.c_void_p), nthreads), "align_batch")
        return r

    def sweep(self, xy, hyp, level=0, nthreads=0, want_scores=True):
        xy = _f32(xy).reshape(-1, 2)
        hyp = _f32(hyp).reshape(-1, 3)
        scores = np.zeros(len(hyp), np.float64) if want_scores else None
        bi = C.c_int64(-1)
        bs = C.c_double(0)
        self._ck(self.L.oracle_sweep(self.h, level, _p(xy, C.c_float), len(xy), _p(hyp, C.c_float),
                                     C.c_int64(len(hyp)), _p(scores, C.c_double) if want_scores else None,
                                     C.byref(bi), C.byref(bs), nthreads), "sweep")
        return scores, bi.value, bs.value

    def num_threads(self):
        return self.L.oracle_num_threads()


def expneg(h):
    return lib().oracle_expneg(C.c_float(h))


def sincos(th):
    """SPEC 4.2: (sin, cos) of a pose angle by the specified f64 sequence."""
    sn, cs = C.c_double(0), C.c_double(0)
    lib().oracle_sincos(C.c_double(th), C.byref(sn), C.byref(cs))
    return sn.value, cs.value


def solve(g, H6, lam):
    g = np.ascontiguousarray(g, np.float64)
    H6 = np.ascontiguousarray(H6, np.float64)
    d = np.zeros(3)
    ok = lib().oracle_solve(_p(g, C.c_double), _p(H6, C.c_double), C.c_double(lam), _p(d, C.c_double))
    return bool(ok), d


def polar_to_points(ranges, angle_min, angle_inc, range_scale=0.001, range_min=0.0, range_max=1e30):
    r = np.ascontiguousarray(ranges)
    out = np.zeros((len(r), 2), np.float32)
    if r.dtype == np.uint16:
        k = lib().oracle_polar_to_points(None, _p(r, C.c_uint16), len(r), C.c_double(angle_min), C.c_double(angle_inc),
                                         C.c_float(range_scale), C.c_float(range_min), C.c_float(range_max), _p(out, C.c_float))
    else:
        r = _f32(r)
        k = lib().oracle_polar_to_points(_p(r, C.c_float), None, len(r), C.c_double(angle_min), C.c_double(angle_inc),
                                         C.c_float(range_scale), C.c_float(range_min), C.c_float(range_max), _p(out, C.c_float))
    return out[:k].copy()

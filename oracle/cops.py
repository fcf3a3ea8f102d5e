"""ctypes front-end of oracle_ops.c (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

numpy in, numpy out; layouts are exactly those of the reference pybind modules
(vgtk/vgtk/cuda/grouping_cuda.cpp:71-86,160-174, gathering_cuda.cpp:29-58,
extensions/chamfer_dist/chamfer_cuda.cpp:22-39).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_ops.so")
_SRC = os.path.join(_HERE, "oracle_ops.c")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def opt_n_threads(n):
    return int(lib().oracle_opt_n_threads(int(n)))


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f(new_xyz), _f(xyz)
    b, _, m = new_xyz.shape
    n = xyz.shape[2]
    idx = np.zeros((b, m, nsample), np.int32)
    lib().oracle_ball_query(b, n, m, ctypes.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx))
    return idx


def furthest_point_sampling(xyz, m):
    xyz = _f(xyz)
    b, _, n = xyz.shape
    idx = np.zeros((b, m), np.int32)
    lib().oracle_fps(b, n, int(m), _p(xyz), _p(idx))
    return idx


def fps_plain(xyz, m):
    xyz = _f(xyz)
    b, _, n = xyz.shape
    idx = np.zeros((b, m), np.int32)
    lib().oracle_fps_plain(b, n, int(m), _p(xyz), _p(idx))
    return idx


def gather_points_forward(points, idx):
    points, idx = _f(points), _i(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.zeros((b, c, m), np.float32)
    lib().oracle_gather_fwd(b, c, n, m, _p(points), _p(idx), _p(out))
    return out


def gather_points_backward(grad_out, idx, npoint):
    grad_out, idx = _f(grad_out), _i(idx)
    b, c, m = grad_out.shape
    out = np.zeros((b, c, npoint), np.float32)
    lib().oracle_gather_bwd(b, c, int(npoint), m, _p(grad_out), _p(idx), _p(out))
    return out


def chamfer_forward(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1, d2 = np.zeros((b, n), np.float32), np.zeros((b, m), np.float32)
    i1, i2 = np.zeros((b, n), np.int32), np.zeros((b, m), np.int32)
    lib().oracle_chamfer_fwd(b, n, _p(xyz1), m, _p(xyz2), _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2, idx1, idx2, g1, g2 = _f(xyz1), _f(xyz2), _i(idx1), _i(idx2), _f(g1), _f(g2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    gx1, gx2 = np.zeros_like(xyz1), np.zeros_like(xyz2)
    lib().oracle_chamfer_bwd(b, n, _p(xyz1), m, _p(xyz2), _p(idx1), _p(idx2), _p(g1), _p(g2), _p(gx1), _p(gx2))
    return gx1, gx2


def knn(pos, centers, k):
    """PointNet2.py:85-87 -> idx int32 [B,S,k], dist [B,S,k] (ascending, ties to the smaller index)."""
    pos, centers = _f(pos), _f(centers)
    b, n, _ = pos.shape
    s = centers.shape[1]
    idx, dist = np.zeros((b, s, k), np.int32), np.zeros((b, s, k), np.float32)
    lib().oracle_knn(b, n, s, int(k), _p(pos), _p(centers), _p(idx), _p(dist))
    return idx, dist


def three_nn(p1, p2):
    """PointNet2.py:114-123 -> idx int32 [B,n2,3], weights [B,n2,3] (unused slots: index 0, weight 0)."""
    p1, p2 = _f(p1), _f(p2)
    b, n1, _ = p1.shape
    n2 = p2.shape[1]
    idx, w = np.zeros((b, n2, 3), np.int32), np.zeros((b, n2, 3), np.float32)
    lib().oracle_three_nn(b, n1, n2, _p(p1), _p(p2), _p(idx), _p(w))
    return idx, w

"""NumPy front-end of the C oracle (oracle/ops_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ops_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs import this module.

Function names and argument order mirror the reference Python wrappers
(tf_ops/sampling/tf_sampling.py:13-57, tf_ops/grouping/tf_grouping.py:9-88).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile liboracle_ops.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle_ops.so")
    src = os.path.join(_HERE, "ops_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    elif os.path.isdir("/root/reference") and not os.path.exists(os.path.join(_HERE, "_ref", "libref_grouping.so")):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        _LIB = ctypes.CDLL(so)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def num_threads():
    return int(lib().oracle_num_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_f if a.dtype == np.float32 else _i)


def farthest_point_sample(npoint, inp):
    inp = _f32(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    lib().oracle_farthest_point_sample(b, n, npoint, _p(inp), _p(out))
    return out


def gather_point(inp, idx):
    inp, idx = _f32(inp), _i32(idx)
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, 3), np.float32)
    lib().oracle_gather_point(b, n, m, _p(inp), _p(idx), _p(out))
    return out


def gather_point_grad(inp, idx, out_g):
    inp, idx, out_g = _f32(inp), _i32(idx), _f32(out_g)
    b, n, _ = inp.shape
    m = idx.shape[1]
    inp_g = np.empty((b, n, 3), np.float32)
    lib().oracle_gather_point_grad(b, n, m, _p(out_g), _p(idx), _p(inp_g))
    return inp_g


def query_ball_point(radius, nsample, xyz1, xyz2, fill=-12345):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.full((b, m, nsample), fill, np.int32)
    cnt = np.zeros((b, m), np.int32)
    lib().oracle_query_ball_point(b, n, m, ctypes.c_float(radius), nsample, _p(xyz1), _p(xyz2), _p(idx), _p(cnt))
    return idx, cnt


def query_ball_point2(radii, nsample, xyz1, xyz2, fill=-12345):
    """Empty rows keep `fill` (the reference leaves them uninitialised, tf_grouping_g.cu:56-90)."""
    xyz1, xyz2, radii = _f32(xyz1), _f32(xyz2), _f32(radii)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.full((b, m, nsample), fill, np.int32)
    cnt = np.zeros((b, m), np.int32)
    lib().oracle_query_ball_point2(b, n, m, nsample, _p(xyz1), _p(xyz2), _p(radii), _p(idx), _p(cnt))
    return idx, cnt


def group_point(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    lib().oracle_group_point(b, n, c, m, ns, _p(points), _p(idx), _p(out))
    return out


def group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f32(points), _i32(idx), _f32(grad_out)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.empty((b, n, c), np.float32)
    lib().oracle_group_point_grad(b, n, c, m, ns, _p(grad_out), _p(idx), _p(g))
    return g


def select_top_k(k, dist):
    dist = _f32(dist)
    b, m, n = dist.shape
    outi = np.empty((b, m, n), np.int32)
    out = np.empty((b, m, n), np.float32)
    lib().oracle_selection_sort(b, n, m, k, _p(dist), _p(outi), _p(out))
    return outi, out


def knn_dist(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, m, n), np.float32)
    lib().oracle_knn_dist(b, n, m, c, _p(xyz1), _p(xyz2), _p(dist))
    return dist


def knn_point(k, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    val = np.empty((b, m, k), np.float32)
    idx = np.empty((b, m, k), np.int32)
    lib().oracle_knn_point(b, n, m, c, k, _p(xyz1), _p(xyz2), _p(val), _p(idx))
    return val, idx


def cumsum(inp):
    inp = _f32(inp)
    b, n = inp.shape
    out = np.empty((b, n), np.float32)
    lib().oracle_cumsum(b, n, _p(inp), _p(out))
    return out


def prob_sample(inp, inpr):
    inp, inpr = _f32(inp), _f32(inpr)
    b, n = inp.shape
    m = inpr.shape[1]
    temp = np.empty((b, n), np.float32)
    out = np.empty((b, m), np.int32)
    lib().oracle_prob_sample(b, n, m, _p(inp), _p(inpr), _p(temp), _p(out))
    return out

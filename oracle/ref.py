"""Loaders for the reference's OWN code, compiled as it lies into oracle/_ref/ (oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Two kinds of reference artefact:

* CPU loops of tf_ops/grouping/test/query_ball_point.cpp:19-84 (ball query without the fallback
  branch, group_point, group_point_grad) -> libref_cpu_grouping.so, called through their C++-mangled
  names with NumPy buffers.  Runs anywhere.
* The reference CUDA kernels tf_ops/{sampling,grouping}/*_g.cu built unmodified for sm_100a ->
  libref_{sampling,grouping}.so, called through the ten mangled launchers
  (tf_sampling_g.cu:194-211, tf_grouping_g.cu:179-199) with torch CUDA tensors.  GPU box only.
  They launch on the legacy default stream and check nothing, so every call here is bracketed by
  a device synchronize and followed by an error check.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p


def available(name):
    return os.path.exists(os.path.join(_REF, name))


# ----------------------------------------------------------------------------- CPU reference loops
_cpu = None


def cpu_lib():
    global _cpu
    if _cpu is None:
        _cpu = ctypes.CDLL(os.path.join(_REF, "libref_cpu_grouping.so"))
    return _cpu


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def cpu_query_ball_point(radius, nsample, xyz1, xyz2, fill=0):
    """test/query_ball_point.cpp:19-48 -- note: no pts_cnt output and no empty-ball fallback."""
    xyz1, xyz2 = _np(xyz1, np.float32), _np(xyz2, np.float32)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.full((b, m, nsample), fill, np.int32)
    fn = getattr(cpu_lib(), "_Z20query_ball_point_cpuiiifiPKfS0_Pi")
    fn(b, n, m, ctypes.c_float(radius), nsample, xyz1.ctypes.data_as(_f), xyz2.ctypes.data_as(_f),
       idx.ctypes.data_as(_i))
    return idx


def cpu_group_point(points, idx):
    points, idx = _np(points, np.float32), _np(idx, np.int32)
    b, n, c = points.shape
    _, m, ns = idx.shape
    out = np.empty((b, m, ns, c), np.float32)
    fn = getattr(cpu_lib(), "_Z15group_point_cpuiiiiiPKfPKiPf")
    fn(b, n, c, m, ns, points.ctypes.data_as(_f), idx.ctypes.data_as(_i), out.ctypes.data_as(_f))
    return out


def cpu_group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _np(points, np.float32), _np(idx, np.int32), _np(grad_out, np.float32)
    b, n, c = points.shape
    _, m, ns = idx.shape
    g = np.zeros((b, n, c), np.float32)  # the reference program relies on a zeroed buffer
    fn = getattr(cpu_lib(), "_Z20group_point_grad_cpuiiiiiPKfPKiPf")
    fn(b, n, c, m, ns, grad_out.ctypes.data_as(_f), idx.ctypes.data_as(_i), g.ctypes.data_as(_f))
    return g


# ----------------------------------------------------------------------------- reference CUDA kernels
_gs = None
_gg = None


def _gpu_libs():
    global _gs, _gg
    if _gs is None:
        _gs = ctypes.CDLL(os.path.join(_REF, "libref_sampling.so"))
        _gg = ctypes.CDLL(os.path.join(_REF, "libref_grouping.so"))
    return _gs, _gg


def _sync():
    import torch

    torch.cuda.synchronize()


def _ptr(t):
    return _vp(t.data_ptr())


def gpu_farthest_point_sample(npoint, inp):
    import torch

    gs, _ = _gpu_libs()
    b, n, _c = inp.shape
    inp = inp.contiguous()
    temp = torch.empty((32, n), dtype=torch.float32, device=inp.device)
    out = torch.zeros((b, npoint), dtype=torch.int32, device=inp.device)
    _sync()
    getattr(gs, "_Z29farthestpointsamplingLauncheriiiPKfPfPi")(b, n, npoint, _ptr(inp), _ptr(temp), _ptr(out))
    _sync()
    return out


def gpu_gather_point(inp, idx):
    import torch

    gs, _ = _gpu_libs()
    b, n, _c = inp.shape
    m = idx.shape[1]
    inp, idx = inp.contiguous(), idx.contiguous()
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    _sync()
    getattr(gs, "_Z19gatherpointLauncheriiiPKfPKiPf")(b, n, m, _ptr(inp), _ptr(idx), _ptr(out))
    _sync()
    return out


def gpu_gather_point_grad(inp, idx, out_g):
    import torch

    gs, _ = _gpu_libs()
    b, n, _c = inp.shape
    m = idx.shape[1]
    idx, out_g = idx.contiguous(), out_g.contiguous()
    inp_g = torch.zeros((b, n, 3), dtype=torch.float32, device=inp.device)
    _sync()
    getattr(gs, "_Z23scatteraddpointLauncheriiiPKfPKiPf")(b, n, m, _ptr(out_g), _ptr(idx), _ptr(inp_g))
    _sync()
    return inp_g


def gpu_prob_sample(inp, inpr):
    import torch

    gs, _ = _gpu_libs()
    b, n = inp.shape
    m = inpr.shape[1]
    inp, inpr = inp.contiguous(), inpr.contiguous()
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
    _sync()
    getattr(gs, "_Z18probsampleLauncheriiiPKfS0_PfPi")(b, n, m, _ptr(inp), _ptr(inpr), _ptr(temp), _ptr(out))
    _sync()
    return out, temp


def gpu_query_ball_point(radius, nsample, xyz1, xyz2):
    import torch

    _, gg = _gpu_libs()
    b, n, _c = xyz1.shape
    m = xyz2.shape[1]
    xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
    idx = torch.full((b, m, nsample), -12345, dtype=torch.int32, device=xyz1.device)
    cnt = torch.zeros((b, m), dtype=torch.int32, device=xyz1.device)
    _sync()
    getattr(gg, "_Z22queryBallPointLauncheriiifiPKfS0_PiS1_")(
        b, n, m, ctypes.c_float(radius), nsample, _ptr(xyz1), _ptr(xyz2), _ptr(idx), _ptr(cnt))
    _sync()
    return idx, cnt


def gpu_query_ball_point2(radii, nsample, xyz1, xyz2):
    import torch

    _, gg = _gpu_libs()
    b, n, _c = xyz1.shape
    m = xyz2.shape[1]
    xyz1, xyz2, radii = xyz1.contiguous(), xyz2.contiguous(), radii.contiguous()
    idx = torch.full((b, m, nsample), -12345, dtype=torch.int32, device=xyz1.device)
    cnt = torch.zeros((b, m), dtype=torch.int32, device=xyz1.device)
    _sync()
    getattr(gg, "_Z23queryBallPoint2LauncheriiiiPKfS0_S0_PiS1_")(
        b, n, m, nsample, _ptr(xyz1), _ptr(xyz2), _ptr(radii), _ptr(idx), _ptr(cnt))
    _sync()
    return idx, cnt


def gpu_select_top_k(k, dist):
    import torch

    _, gg = _gpu_libs()
    b, m, n = dist.shape
    dist = dist.contiguous()
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    _sync()
    getattr(gg, "_Z21selectionSortLauncheriiiiPKfPiPf")(b, n, m, k, _ptr(dist), _ptr(outi), _ptr(out))
    _sync()
    return outi, out


def gpu_group_point(points, idx):
    import torch

    _, gg = _gpu_libs()
    b, n, c = points.shape
    _b, m, ns = idx.shape
    points, idx = points.contiguous(), idx.contiguous()
    out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
    _sync()
    getattr(gg, "_Z18groupPointLauncheriiiiiPKfPKiPf")(b, n, c, m, ns, _ptr(points), _ptr(idx), _ptr(out))
    _sync()
    return out


def gpu_group_point_grad(points, idx, grad_out):
    import torch

    _, gg = _gpu_libs()
    b, n, c = points.shape
    _b, m, ns = idx.shape
    idx, grad_out = idx.contiguous(), grad_out.contiguous()
    g = torch.zeros((b, n, c), dtype=torch.float32, device=points.device)
    _sync()
    getattr(gg, "_Z22groupPointGradLauncheriiiiiPKfPKiPf")(b, n, c, m, ns, _ptr(grad_out), _ptr(idx), _ptr(g))
    _sync()
    return g

"""Grouping operators -- same names, argument order and return arity as the reference's
tf_ops/grouping/tf_grouping.py:9-88, on torch CUDA tensors, backed by csrc/grouping.cu and csrc/scatter.cu.

Argument checks follow tf_grouping.cpp:90-103,132-148,179-184,215-221,246-257 (ValueError <-> InvalidArgument).
"""
import importlib

import torch

# resolves both as 3dfeatnet_b200.tf_ops... (package) and as tf_ops... (drop-in: package dir on sys.path)
_ROOT = __name__.split(".")[0]
_lib = importlib.import_module("3dfeatnet_b200._lib" if _ROOT == "3dfeatnet_b200" else "_lib")


def _f32(t, name):
    if t.dtype != torch.float32:
        raise ValueError("%s must be float32" % name)
    _lib.require_cuda(t)
    return t.contiguous()


def _i32(t, name):
    if t.dtype != torch.int32:
        raise ValueError("%s must be int32" % name)
    _lib.require_cuda(t)
    return t.contiguous()


def _check_xyz_pair(op, xyz1, xyz2):
    if xyz1.dim() != 3 or xyz1.shape[2] != 3:
        raise ValueError("%s expects (batch_size, ndataset, 3) xyz1 shape." % op)
    if xyz2.dim() != 3 or xyz2.shape[2] != 3 or xyz2.shape[0] != xyz1.shape[0]:
        raise ValueError("%s expects (batch_size, npoint, 3) xyz2 shape." % op)


class BallGrid(object):
    """The spatial binning of a cloud for query_ball_point (f3d_ball_grid_build): build it once, hand it to every query on the same
    (radius, xyz1) as `grid=` -- the attention pass of the file flow queries one cloud five times (inference.py:118-131)."""
    __slots__ = ("ws", "ws_bytes", "b", "n", "radius", "xyz1")

    def __init__(self, radius, xyz1, max_centres=0):
        xyz1 = _f32(xyz1.detach(), "xyz1")
        _lib.require_cuda(xyz1)
        L = _lib.lib()
        self.b, self.n, self.radius, self.xyz1 = xyz1.shape[0], xyz1.shape[1], float(radius), xyz1
        nbytes = L.f3d_query_ball_point_workspace_bytes(self.b, self.n)
        if max_centres >= 4096:  # room for the spatial binning of the centres of a many-centre query
            nbytes = ((nbytes + 255) // 256) * 256 + L.f3d_query_ball_point_workspace_bytes(self.b, int(max_centres))
        self.ws_bytes = nbytes
        self.ws = torch.empty((nbytes,), dtype=torch.uint8, device=xyz1.device)
        _lib.check(L.f3d_ball_grid_build(self.b, self.n, self.radius, _lib.ptr(xyz1), _lib.ptr(self.ws), nbytes, _lib.stream()), "ball_grid_build")


def query_ball_point(radius, nsample, xyz1, xyz2, use_grid=True, grid=None):
    '''
    Input:
        radius: float32, ball search radius
        nsample: int32, number of points selected in each ball region
        xyz1: (batch_size, ndataset, 3) float32 array, input points
        xyz2: (batch_size, npoint, 3) float32 array, query points
    Output:
        idx: (batch_size, npoint, nsample) int32 array, indices to input points
        pts_cnt: (batch_size, npoint) int32 array, number of unique points in each local region
    (tf_grouping.py:9-21; NoGradient)
    '''
    if not radius > 0:
        raise ValueError("QueryBallPoint expects positive radius")  # tf_grouping.cpp:90
    if nsample <= 0:
        raise ValueError("QueryBallPoint expects positive nsample")  # :93
    _check_xyz_pair("QueryBallPoint", xyz1, xyz2)
    xyz1, xyz2 = _f32(xyz1.detach(), "xyz1"), _f32(xyz2.detach(), "xyz2")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.empty((b, m, nsample), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    L = _lib.lib()
    if grid is not None:  # a BallGrid of this cloud and radius: only the query runs
        if (grid.b, grid.n, grid.radius) != (b, n, float(radius)) or grid.xyz1.data_ptr() != xyz1.data_ptr():
            raise ValueError("query_ball_point: grid was built for another cloud or radius")
        _lib.check(L.f3d_ball_grid_query(b, n, m, float(radius), nsample, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx), _lib.ptr(cnt),
                                         _lib.ptr(grid.ws), grid.ws_bytes, _lib.stream()), "ball_grid_query")
        return idx, cnt
    if use_grid:  # grid-accelerated kernel (identical results); use_grid=False forces the plain scan
        ws_bytes = L.f3d_query_ball_point_workspace_bytes(b, n)
        if m >= 4096:  # room for the spatial binning of the centres (many centres per cloud: the attention pass of inference.py)
            ws_bytes = ((ws_bytes + 255) // 256) * 256 + L.f3d_query_ball_point_workspace_bytes(b, m)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=xyz1.device)
        _lib.check(L.f3d_query_ball_point_ws(b, n, m, float(radius), nsample, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx),
                                             _lib.ptr(cnt), _lib.ptr(ws), ws_bytes, _lib.stream()), "query_ball_point")
    else:
        _lib.check(L.f3d_query_ball_point(b, n, m, float(radius), nsample, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(idx),
                                          _lib.ptr(cnt), _lib.stream()), "query_ball_point")
    return idx, cnt


def query_ball_point2(radii, nsample, xyz1, xyz2):
    '''
    Input:
        radii: (batch_size, npoint), ball search radius
        nsample: int32, number of points selected in each ball region
        xyz1: (batch_size, ndataset, 3) float32 array, input points
        xyz2: (batch_size, npoint, 3) float32 array, query points
    Output:
        idx: (batch_size, npoint, nsample) int32 array, indices to input points
        pts_cnt: (batch_size, npoint) int32 array, number of unique points in each local region
    Rows of empty balls are undefined in the reference (tf_grouping_g.cu:56-90); here they are zero.
    '''
    if nsample <= 0:
        raise ValueError("QueryBallPoint2 expects positive nsample")  # tf_grouping.cpp:132
    _check_xyz_pair("QueryBallPoint2", xyz1, xyz2)
    if radii.dim() != 2 or radii.shape[0] != xyz2.shape[0] or radii.shape[1] != xyz2.shape[1]:
        raise ValueError("QueryBallPoint2 expects (batch_size, npoint) radii shape.")  # :148
    xyz1, xyz2, radii = _f32(xyz1.detach(), "xyz1"), _f32(xyz2.detach(), "xyz2"), _f32(radii.detach(), "radii")
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.zeros((b, m, nsample), dtype=torch.int32, device=xyz1.device)
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    L = _lib.lib()
    _lib.check(L.f3d_query_ball_point2(b, n, m, nsample, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(radii), _lib.ptr(idx),
                                       _lib.ptr(cnt), _lib.stream()), "query_ball_point2")
    return idx, cnt


def select_top_k(k, dist):
    '''
    Input:
        k: int32, number of k SMALLEST elements selected
        dist: (b,m,n) float32 array, distance matrix, m query points, n dataset points
    Output:
        idx: (b,m,n) int32 array, first k in n are indices to the top k
        dist_out: (b,m,n) float32 array, first k in n are the top k
    (tf_grouping.py:37-46; NoGradient)
    '''
    if k <= 0:
        raise ValueError("SelectionSort expects positive k")  # tf_grouping.cpp:179
    if dist.dim() != 3:
        raise ValueError("SelectionSort expects (b,m,n) dist shape.")  # :184
    dist = _f32(dist.detach(), "dist")
    b, m, n = dist.shape
    outi = torch.empty((b, m, n), dtype=torch.int32, device=dist.device)
    out = torch.empty((b, m, n), dtype=torch.float32, device=dist.device)
    L = _lib.lib()
    _lib.check(L.f3d_selection_sort(b, n, m, k, _lib.ptr(dist), _lib.ptr(outi), _lib.ptr(out), _lib.stream()),
               "select_top_k")
    return outi, out


class _GroupPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, idx):
        b, n, c = points.shape
        _, m, ns = idx.shape
        out = torch.empty((b, m, ns, c), dtype=torch.float32, device=points.device)
        L = _lib.lib()
        _lib.check(L.f3d_group_point(b, n, c, m, ns, _lib.ptr(points), _lib.ptr(idx), _lib.ptr(out), _lib.stream()),
                   "group_point")
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return group_point_grad(ctx.n, idx, grad_out), None


def group_point(points, idx):
    '''
    Input:
        points: (batch_size, ndataset, channel) float32 array, points to sample from
        idx: (batch_size, npoint, nsample) int32 array, indices to points
    Output:
        out: (batch_size, npoint, nsample, channel) float32 array, values sampled from points
    (tf_grouping.py:48-56; gradient = GroupPointGrad, :57-61)
    '''
    if points.dim() != 3:
        raise ValueError("GroupPoint expects (batch_size, num_points, channel) points shape")  # tf_grouping.cpp:215
    if idx.dim() != 3 or idx.shape[0] != points.shape[0]:
        raise ValueError("GroupPoint expects (batch_size, npoints, nsample) idx shape")  # :221
    return _GroupPoint.apply(_f32(points, "points"), _i32(idx, "idx"))


def group_point_grad(n, idx, grad_out):
    """GroupPointGrad (tf_grouping.cpp:240-274): scatter-add (b,m,nsample,c) -> (b,n,c); deterministic, no atomics."""
    idx = _i32(idx, "idx")
    grad_out = _f32(grad_out, "grad_out")
    b, m, ns = idx.shape
    if grad_out.dim() != 4 or grad_out.shape[:3] != (b, m, ns):
        raise ValueError("GroupPointGrad expects (batch_size, npoints, nsample, channel) grad_out shape")  # :257
    c = grad_out.shape[3]
    g = torch.empty((b, n, c), dtype=torch.float32, device=grad_out.device)
    L = _lib.lib()
    ws_bytes = L.f3d_scatter_add_workspace_bytes(b, n, m * ns)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=grad_out.device)
    _lib.check(L.f3d_group_point_grad(b, n, c, m, ns, _lib.ptr(grad_out), _lib.ptr(idx), _lib.ptr(g), _lib.ptr(ws),
                                      ws_bytes, _lib.stream()), "group_point_grad")
    return g


def knn_point(k, xyz1, xyz2):
    '''
    Input:
        k: int32, number of k in k-nn search
        xyz1: (batch_size, ndataset, c) float32 array, input points
        xyz2: (batch_size, npoint, c) float32 array, query points
    Output:
        val: (batch_size, npoint, k) float32 array, L2 distances   (squared, as in the reference, tf_grouping.py:81)
        idx: (batch_size, npoint, k) int32 array, indices to input points
    (tf_grouping.py:63-88)
    '''
    if k <= 0:
        raise ValueError("knn_point expects positive k")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[0] != xyz2.shape[0] or xyz1.shape[2] != xyz2.shape[2]:
        raise ValueError("knn_point expects (b,n,c) xyz1 and (b,m,c) xyz2")
    xyz1, xyz2 = _f32(xyz1.detach(), "xyz1"), _f32(xyz2.detach(), "xyz2")
    b, n, c = xyz1.shape
    m = xyz2.shape[1]
    val = torch.empty((b, m, k), dtype=torch.float32, device=xyz1.device)
    idx = torch.empty((b, m, k), dtype=torch.int32, device=xyz1.device)
    L = _lib.lib()
    ws_bytes = L.f3d_knn_workspace_bytes(b, n, m, c, k)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=xyz1.device)
    _lib.check(L.f3d_knn_point(b, n, m, c, k, _lib.ptr(xyz1), _lib.ptr(xyz2), _lib.ptr(val), _lib.ptr(idx), _lib.ptr(ws),
                               ws_bytes, _lib.stream()), "knn_point")
    return val, idx

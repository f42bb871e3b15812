"""Sampling operators -- same names, argument order and return arity as the reference's
tf_ops/sampling/tf_sampling.py:13-57, on torch CUDA tensors, backed by csrc/sampling.cu and csrc/scatter.cu.

Argument checks follow the TF op kernels (tf_sampling.cpp:76-79,99,105,131-135,156-167) and raise ValueError where
the reference raises InvalidArgument.
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


def farthest_point_sample(npoint, inp):
    """
    input:
        int32
        batch_size * ndataset * 3   float32
    returns:
        batch_size * npoint         int32
    (tf_sampling.py:48-56; NoGradient)
    """
    if npoint <= 0:
        raise ValueError("FarthestPointSample expects positive npoint")  # tf_sampling.cpp:99
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("FarthestPointSample expects (batch_size,num_points,3) inp shape")  # :105
    inp = _f32(inp.detach(), "inp")
    b, n, _ = inp.shape
    out = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    temp = None
    if n > 131072:  # beyond the 8-CTA cluster kernel: global-memory fallback needs a (b,n) scratch
        temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    L = _lib.lib()
    _lib.check(L.f3d_farthest_point_sample(b, n, npoint, _lib.ptr(inp), _lib.ptr(temp), _lib.ptr(out), _lib.stream()),
               "farthest_point_sample")
    return out


class _GatherPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, idx):
        b, n, _ = inp.shape
        m = idx.shape[1]
        out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
        L = _lib.lib()
        _lib.check(L.f3d_gather_point(b, n, m, _lib.ptr(inp), _lib.ptr(idx), _lib.ptr(out), _lib.stream()), "gather_point")
        ctx.save_for_backward(idx)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, out_g):
        (idx,) = ctx.saved_tensors
        return gather_point_grad(ctx.n, idx, out_g), None


def gather_point(inp, idx):
    """
    input:
        batch_size * ndataset * 3   float32
        batch_size * npoints        int32
    returns:
        batch_size * npoints * 3    float32
    (tf_sampling.py:29-37; gradient = GatherPointGrad, :43-47)
    """
    if inp.dim() != 3 or inp.shape[2] != 3:
        raise ValueError("GatherPoint expects (batch_size,num_points,3) inp shape")  # tf_sampling.cpp:131
    if idx.dim() != 2 or idx.shape[0] != inp.shape[0]:
        raise ValueError("GatherPoint expects (batch_size,num_result) idx shape")  # :135
    return _GatherPoint.apply(_f32(inp, "inp"), _i32(idx, "idx"))


def gather_point_grad(n, idx, out_g):
    """GatherPointGrad (tf_sampling.cpp:151-178): scatter-add of out_g (b,m,3) into zeros (b,n,3); deterministic."""
    idx = _i32(idx, "idx")
    out_g = _f32(out_g, "out_g")
    b, m = idx.shape
    if out_g.shape != (b, m, 3):
        raise ValueError("GatherPointGrad expects (batch_size,num_result,3) out_g shape")  # :167
    inp_g = torch.empty((b, n, 3), dtype=torch.float32, device=out_g.device)
    L = _lib.lib()
    ws_bytes = L.f3d_scatter_add_workspace_bytes(b, n, m)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=out_g.device)
    _lib.check(L.f3d_gather_point_grad(b, n, m, _lib.ptr(out_g), _lib.ptr(idx), _lib.ptr(inp_g), _lib.ptr(ws), ws_bytes,
                                       _lib.stream()), "gather_point_grad")
    return inp_g


def prob_sample(inp, inpr):
    """
    input:
        batch_size * ncategory float32
        batch_size * npoints   float32
    returns:
        batch_size * npoints   int32
    (tf_sampling.py:13-21; NoGradient)
    """
    if inp.dim() != 2 or inpr.dim() != 2 or inp.shape[0] != inpr.shape[0]:
        raise ValueError("ProbSample expects (batch_size,num_choices) inp and (batch_size,num_points) inpr")  # :76-79
    inp, inpr = _f32(inp.detach(), "inp"), _f32(inpr.detach(), "inpr")
    b, n = inp.shape
    m = inpr.shape[1]
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    out = torch.empty((b, m), dtype=torch.int32, device=inp.device)
    L = _lib.lib()
    _lib.check(L.f3d_prob_sample(b, n, m, _lib.ptr(inp), _lib.ptr(inpr), _lib.ptr(temp), _lib.ptr(out), _lib.stream()),
               "prob_sample")
    return out

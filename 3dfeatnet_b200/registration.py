"""Descriptor matching and RANSAC rigid registration on the device -- the step the reference leaves to MATLAB
(scripts/computeAndVisualizeMatches.m:43-52, scripts/external/{ransacfitRt,ransac,estimateRigidTransform,quat2rot}.m).

    matches = match_descriptors(desc1, desc2)                       # pdist2(desc2, desc1, 'euclidean', 'smallest', 1)
    Rt, inliers, trialcount = ransacfitRt(xyz1[...], xyz2[matches], 1.0)

MATLAB's random stream (randsample on the reset global stream) cannot be reproduced, so the 3-point samples come from a seeded
numpy Generator (or from the caller, `triples=`); given the same samples the result follows the reference's sequential algorithm.
No CPU fallback: CUDA tensors only."""
import importlib

import numpy as np
import torch

_ROOT = __name__.split(".")[0]
_lib = importlib.import_module("3dfeatnet_b200._lib" if _ROOT == "3dfeatnet_b200" else "_lib")

MAX_TRIALS = 10000  # ransac.m:107


def match_descriptors(desc1, desc2, return_dist=False):
    """desc1 (n1,d), desc2 (n2,d) CUDA float32 -> int32 (n1,): the row of desc2 nearest to each row of desc1 (lowest index on
    ties), i.e. `matches12` of computeAndVisualizeMatches.m:43-44 (0-based)."""
    _lib.require_cuda(desc1, desc2)
    if desc1.dim() != 2 or desc2.dim() != 2 or desc1.shape[1] != desc2.shape[1] or desc2.shape[0] == 0:
        raise ValueError("match_descriptors expects (n1,d) and (n2,d) with n2 > 0")
    if desc1.dtype != torch.float32 or desc2.dtype != torch.float32:
        raise ValueError("match_descriptors expects float32 tensors")
    desc1, desc2 = desc1.contiguous(), desc2.contiguous()
    n1, d = desc1.shape
    match = torch.empty((n1,), dtype=torch.int32, device=desc1.device)
    dist2 = torch.empty((n1,), dtype=torch.float32, device=desc1.device) if return_dist else None
    _lib.check(_lib.lib().f3d_match_descriptors(n1, desc2.shape[0], d, _lib.ptr(desc1), _lib.ptr(desc2), _lib.ptr(match), _lib.ptr(dist2),
                                                _lib.stream()), "match_descriptors")
    return (match, dist2) if return_dist else match


def draw_triples(npts, ntrials=MAX_TRIALS + 1, seed=0):
    """ntrials samples of 3 distinct correspondence indices (randsample(npts, 3), ransac.m:137-141)."""
    if npts < 3:
        raise ValueError("need at least 3 correspondences")
    rng = np.random.default_rng(seed)
    t = np.empty((ntrials, 3), dtype=np.int64)
    t[:, 0] = rng.integers(0, npts, ntrials)
    t[:, 1] = rng.integers(0, npts - 1, ntrials)
    t[:, 2] = rng.integers(0, npts - 2, ntrials)
    # map the draws from shrinking ranges onto distinct indices (Fisher-Yates on 3 slots)
    t[:, 1] += t[:, 1] >= t[:, 0]
    lo, hi = np.minimum(t[:, 0], t[:, 1]), np.maximum(t[:, 0], t[:, 1])
    t[:, 2] += t[:, 2] >= lo
    t[:, 2] += t[:, 2] >= hi
    return t.astype(np.int32)


def estimateRt(pts1, pts2, mask=None):
    """estimateRt.m: least-squares rigid transform with pts1 ~ R pts2 + t.  (n,3) CUDA float32 -> (3,4) float64 tensor."""
    _lib.require_cuda(pts1, pts2)
    pts1, pts2 = pts1.contiguous().float(), pts2.contiguous().float()
    if pts1.shape != pts2.shape or pts1.dim() != 2 or pts1.shape[1] != 3:
        raise ValueError("estimateRt expects two (n,3) tensors")
    if pts1.shape[0] < 3:
        raise ValueError("At least 3 point matches are needed")  # estimateRigidTransform.m:52-54
    Rt = torch.empty((3, 4), dtype=torch.float64, device=pts1.device)
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    _lib.check(_lib.lib().f3d_rigid_fit(pts1.shape[0], _lib.ptr(pts1), _lib.ptr(pts2), _lib.ptr(m), _lib.ptr(Rt), None, _lib.stream()),
               "rigid_fit")
    return Rt


def ransacfitRt(pts1, pts2, t, triples=None, seed=0, max_trials=MAX_TRIALS):
    """ransacfitRt(x, t) of scripts/external/ransacfitRt.m with x = [pts1'; pts2'].  pts1, pts2: (n,3) CUDA float32.
    Returns (Rt (3,4) float64 tensor or None, inlier indices (int64 tensor), trialcount)."""
    _lib.require_cuda(pts1, pts2)
    pts1, pts2 = pts1.contiguous().float(), pts2.contiguous().float()
    if pts1.shape != pts2.shape or pts1.dim() != 2 or pts1.shape[1] != 3:
        raise ValueError("ransacfitRt expects two (n,3) tensors")
    if not t > 0:
        raise ValueError("the distance threshold must be positive")
    n, dev = pts1.shape[0], pts1.device
    if n < 3:  # ransacfitRt.m:44-48
        return None, torch.zeros((0,), dtype=torch.int64, device=dev), 0
    if triples is None:
        triples = draw_triples(n, max_trials + 1, seed) if n > 3 else np.zeros((0, 3), np.int32)
    tri = torch.as_tensor(np.ascontiguousarray(triples, dtype=np.int32)).to(dev)
    if tri.dim() != 2 or (tri.numel() and tri.shape[1] != 3):
        raise ValueError("triples must be (ntrials,3)")
    if tri.numel() and (int(tri.min()) < 0 or int(tri.max()) >= n):
        raise ValueError("sample index out of range")
    L = _lib.lib()
    ntr = tri.shape[0]
    ws_bytes = L.f3d_ransac_workspace_bytes(n, ntr)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    Rt = torch.empty((3, 4), dtype=torch.float64, device=dev)
    mask = torch.empty((n,), dtype=torch.uint8, device=dev)
    info = torch.empty((4,), dtype=torch.int32, device=dev)
    _lib.check(L.f3d_ransac_fit_rt(n, _lib.ptr(pts1), _lib.ptr(pts2), ntr, _lib.ptr(tri) if ntr else None, float(t), int(max_trials),
                                   _lib.ptr(Rt), _lib.ptr(mask), _lib.ptr(info), _lib.ptr(ws), ws_bytes, _lib.stream()), "ransac_fit_rt")
    num, trialcount, _, status = info.cpu().tolist()
    if status:
        raise _lib.F3DError("ransacfitRt: %d sample triples ran out before the stopping rule was met" % ntr)
    inliers = torch.nonzero(mask, as_tuple=False).flatten()
    if inliers.numel() < 3:  # ransacfitRt.m:70-73
        return None, torch.zeros((0,), dtype=torch.int64, device=dev), trialcount
    return Rt, inliers, trialcount


def register(xyz1, desc1, xyz2, desc2, t=1.0, seed=0):
    """The per-pair body of computeAndVisualizeMatches.m:42-52: match, then RANSAC.  Returns (Rt, inlier indices into the
    matches, trialcount, matches12 (n1,2) int32 [index in cloud 1, index in cloud 2])."""
    m = match_descriptors(desc1, desc2)
    Rt, inl, trials = ransacfitRt(xyz1.contiguous(), xyz2[m.long()].contiguous(), t, seed=seed)
    matches12 = torch.stack([torch.arange(m.numel(), dtype=torch.int32, device=m.device), m], dim=1)
    return Rt, inl, trials, matches12

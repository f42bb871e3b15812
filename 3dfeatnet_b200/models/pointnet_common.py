"""Sample-and-group compositions on torch CUDA tensors.

Public names, argument order and return arity are those of the reference's models/pointnet_common.py
(sample_points :14, query_and_group_points :32, sample_and_group :69, sample_and_group_all :138) so that code written
against that module keeps working; the bodies are built from three private helpers on top of the CUDA operators of
tf_ops/ (no TF, no CPU path).  Every op here is differentiable through the GatherPoint / GroupPoint gradients.
"""
import importlib

import torch

_ROOT = __name__.split(".")[0]
_pfx = "3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else ""
_tg = importlib.import_module(_pfx + "tf_ops.grouping.tf_grouping")
_ts = importlib.import_module(_pfx + "tf_ops.sampling.tf_sampling")
query_ball_point, group_point, knn_point = _tg.query_ball_point, _tg.group_point, _tg.knn_point
farthest_point_sample, gather_point = _ts.farthest_point_sample, _ts.gather_point


# ------------------------------------------------------------------------------------------------ private helpers
def _neighbour_indices(xyz, centres, nsample, radius, knn, neighbours=None):
    """(idx (B,M,S) int32, pts_cnt): ball query, or kNN where the reference sets pts_cnt to nsample ("Hack", :37/:99).
    The reference runs the identical ball query twice per forward (detector :39, descriptor :102, same cloud, same centres).
    This module keeps no state between calls: a caller that already holds the answer for exactly these arguments hands it
    over as `neighbours=(idx, pts_cnt)` (models/feat3dnet.py passes the detector's query to the descriptor)."""
    if neighbours is not None:
        idx, pts_cnt = neighbours
        if tuple(idx.shape) != (xyz.shape[0], centres.shape[1], nsample):
            raise ValueError("neighbours: idx has shape %s, expected %s" % (tuple(idx.shape), (xyz.shape[0], centres.shape[1], nsample)))
        return idx, pts_cnt
    if knn:
        return knn_point(nsample, xyz, centres)[1], nsample
    return query_ball_point(radius, nsample, xyz, centres)


class _LocalFrames(torch.autograd.Function):
    """group_point - centre, / radius, rotation about z by the cluster's angle: one CUDA launch (csrc/frames.cu) instead of a gather
    and a dozen element-wise passes; differentiable with respect to the angles (what the training step needs).  Returns
    (rotated, before, R): `before` = the rows ahead of the rotation, R (B,M,3,3) the rotation matrices (None unless want_before)."""

    @staticmethod
    def forward(ctx, xyz, centres, idx, angles, radius, normalize_radius, clockwise, want_before):
        _lib = importlib.import_module(_pfx + "_lib")
        x, c, i = xyz.detach().contiguous().float(), centres.detach().contiguous().float(), idx.contiguous()
        a = angles.detach().contiguous().float() if angles is not None else None
        _lib.require_cuda(x, c, i)
        b, n, m, s = x.shape[0], x.shape[1], c.shape[1], i.shape[2]
        out = torch.empty((b, m, s, 3), dtype=torch.float32, device=x.device)
        before = torch.empty_like(out) if want_before else None
        rot = torch.empty((b, m, 3, 3), dtype=torch.float32, device=x.device) if (want_before and a is not None) else None
        _lib.check(_lib.lib().f3d_group_local_frames(b, n, m, s, _lib.ptr(x), _lib.ptr(c), _lib.ptr(i), _lib.ptr(a) if a is not None else None,
                                                     int(bool(clockwise)), float(radius), int(bool(normalize_radius)),
                                                     _lib.ptr(before) if before is not None else None, _lib.ptr(rot) if rot is not None else None,
                                                     _lib.ptr(out), _lib.stream()), "group_local_frames")
        ctx.clockwise = bool(clockwise)
        ctx.has_angles = a is not None
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(out)
        ctx.mark_non_differentiable(*[t for t in (before, rot) if t is not None])
        return out, before, rot

    @staticmethod
    def backward(ctx, gout, _gbefore, _grot):
        if not ctx.has_angles or gout is None or not ctx.needs_input_grad[3]:
            return None, None, None, None, None, None, None, None
        _lib = importlib.import_module(_pfx + "_lib")
        (out,) = ctx.saved_tensors
        b, m, s, _ = out.shape
        g = gout.contiguous().float()
        dangle = torch.empty((b, m), dtype=torch.float32, device=out.device)
        _lib.check(_lib.lib().f3d_group_local_frames_angle_grad(b, m, s, _lib.ptr(out), _lib.ptr(g), int(ctx.clockwise), _lib.ptr(dangle), _lib.stream()),
                   "group_local_frames_angle_grad")
        return None, None, None, dangle, None, None, None, None


FUSED_FRAMES = True  # False: the op-by-op statement below also where the fused op applies


def _fused_frames_ok(xyz, centres, idx):
    """The fused op carries no gradient to xyz / centres: it serves the forwards where neither needs one (training and inference of the
    model; the saliency gradients of compute_det_gradients differentiate with respect to xyz and take the op-by-op statement)."""
    needs = torch.is_grad_enabled() and (xyz.requires_grad or centres.requires_grad)
    return FUSED_FRAMES and xyz.is_cuda and xyz.dim() == 3 and xyz.shape[2] == 3 and idx.dtype == torch.int32 and not needs


def _local_frames(xyz, centres, idx, radius, normalize_radius):
    """Gather the neighbourhoods and express them relative to their centre, optionally in units of the radius."""
    local = group_point(xyz, idx) - centres.unsqueeze(2)
    return local / radius if normalize_radius else local


def _spin_about_z(local, angles, clockwise):
    """Rotate every neighbourhood about z by its own angle.
    clockwise=False: x' = x cos - y sin, y' = x sin + y cos  (sample_and_group: grouped_xyz @ [[c,s,0],[-s,c,0],[0,0,1]], :112-119)
    clockwise=True : x' = x cos + y sin, y' = -x sin + y cos (query_and_group_points' own convention, :49-54)"""
    c = torch.cos(angles).unsqueeze(2)
    s = torch.sin(angles).unsqueeze(2)
    if clockwise:
        s = -s
    x, y, z = local.unbind(dim=3)
    return torch.stack((x * c - y * s, x * s + y * c, z), dim=3)


def _with_features(local_xyz, points, idx, use_xyz):
    if points is None:
        return local_xyz
    feats = group_point(points, idx)
    return torch.cat((local_xyz, feats), dim=-1) if use_xyz else feats


# ------------------------------------------------------------------------------------------------ reference surface
def sample_points(xyz, npoint):
    """Cluster centres: `npoint` farthest-point samples of xyz, or every point when npoint <= 0 (reference :24-25)."""
    if npoint <= 0:
        return xyz.clone()
    return gather_point(xyz, farthest_point_sample(npoint, xyz))


def query_and_group_points(xyz, points, new_xyz, nsample, radius, knn=False,
                           use_xyz=True, normalize_radius=True, orientations=None, end_points=None, neighbours=None):
    """Detector-side grouping (reference :32-66): returns (new_points (B,M,S,3[+C]), idx).
    The pts_cnt the reference only logs as a histogram (:41) is stored in the caller's `end_points` dict (when given) as
    end_points['pts_cnt']; `neighbours=(idx, pts_cnt)` skips the query (see _neighbour_indices)."""
    idx, pts_cnt = _neighbour_indices(xyz, new_xyz, nsample, radius, knn, neighbours)
    if end_points is not None:
        end_points['pts_cnt'] = pts_cnt
    if _fused_frames_ok(xyz, new_xyz, idx):
        local = _LocalFrames.apply(xyz, new_xyz, idx, orientations, radius, normalize_radius, True, False)[0]
    else:
        local = _local_frames(xyz, new_xyz, idx, radius, normalize_radius)
        if orientations is not None:
            local = _spin_about_z(local, orientations, clockwise=True)
    return _with_features(local, points, idx, use_xyz), idx


def sample_and_group(npoint, radius, nsample, xyz, points, tnet_spec=None, knn=False, use_xyz=True,
                     keypoints=None, orientations=None, normalize_radius=False, neighbours=None):
    """Descriptor-side sampling + grouping (reference :69-135).  `neighbours=(idx, pts_cnt)`: the answer of the identical
    query the detector already ran on (xyz, keypoints, nsample, radius) -- see _neighbour_indices.
    Returns (new_xyz (B,M,3), new_points (B,M,S,3[+C]), idx (B,M,S), grouped_xyz (B,M,S,3), end_points) where end_points
    carries 'grouped_xyz_before', 'rotation' (when orientations are given), 'grouped_xyz' and 'pts_cnt'."""
    if tnet_spec is not None:  # the reference calls an undefined tnet() here (:122-123); unused by 3DFeat-Net
        raise ValueError("tnet_spec is not supported")
    centres = keypoints if keypoints is not None else sample_points(xyz, npoint)
    idx, pts_cnt = _neighbour_indices(xyz, centres, nsample, radius, knn, neighbours)
    if _fused_frames_ok(xyz, centres, idx):
        local, before, rot = _LocalFrames.apply(xyz, centres, idx, orientations, radius, normalize_radius, False, orientations is not None)
        end_points = {'pts_cnt': pts_cnt, 'grouped_xyz_before': local if before is None else before, 'grouped_xyz': local}
        if rot is not None:
            end_points['rotation'] = rot  # (B,M,3,3), the reference's R
        return centres, _with_features(local, points, idx, use_xyz), idx, local, end_points
    before = _local_frames(xyz, centres, idx, radius, normalize_radius)
    local = before
    end_points = {'pts_cnt': pts_cnt, 'grouped_xyz_before': before}
    if orientations is not None:
        local = _spin_about_z(before, orientations, clockwise=False)
        c, s = torch.cos(orientations), torch.sin(orientations)
        o, z = torch.ones_like(c), torch.zeros_like(c)
        end_points['rotation'] = torch.stack((torch.stack((c, s, z), -1), torch.stack((-s, c, z), -1),
                                              torch.stack((z, z, o), -1)), dim=-2)  # (B,M,3,3), the reference's R
    end_points['grouped_xyz'] = local
    return centres, _with_features(local, points, idx, use_xyz), idx, local, end_points


def sample_and_group_all(xyz, points, use_xyz=True):
    """One group holding the whole cloud, centred on the origin (reference :138-165)."""
    b, n = xyz.shape[0], xyz.shape[1]
    centre = xyz.new_zeros((b, 1, 3))
    idx = torch.arange(n, dtype=torch.int32, device=xyz.device).expand(b, 1, n).contiguous()
    whole = xyz.reshape(b, 1, n, 3)
    if points is None:
        return centre, whole, idx, whole
    feats = torch.cat((xyz, points), dim=2) if use_xyz else points
    return centre, feats.unsqueeze(1), idx, whole

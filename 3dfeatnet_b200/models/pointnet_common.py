"""PointNet++ sample-and-group compositions -- the reference's models/pointnet_common.py:14-165 on torch CUDA
tensors, built on the CUDA operators of tf_ops/ (no TF, no CPU path)."""
import importlib

import torch

_ROOT = __name__.split(".")[0]
_pfx = "3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else ""
_tg = importlib.import_module(_pfx + "tf_ops.grouping.tf_grouping")
_ts = importlib.import_module(_pfx + "tf_ops.sampling.tf_sampling")
query_ball_point, group_point, knn_point = _tg.query_ball_point, _tg.group_point, _tg.knn_point
farthest_point_sample, gather_point = _ts.farthest_point_sample, _ts.gather_point


def sample_points(xyz, npoint):
    '''
    :param xyz:
    :param npoint:
    :return: new_xyz - Cluster centers   (pointnet_common.py:14-29: identity when npoint <= 0)
    '''
    if npoint <= 0:
        new_xyz = xyz.clone()
    else:
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))
    return new_xyz


def query_and_group_points(xyz, points, new_xyz, nsample, radius, knn=False,
                           use_xyz=True, normalize_radius=True, orientations=None):
    """pointnet_common.py:32-66.  Returns (new_points, idx); `query_and_group_points.last_pts_cnt` keeps the
    pts_cnt the reference only logs as a histogram (:41)."""
    if knn:
        _, idx = knn_point(nsample, xyz, new_xyz)
        pts_cnt = nsample
    else:
        idx, pts_cnt = query_ball_point(radius, nsample, xyz, new_xyz)
    query_and_group_points.last_pts_cnt = pts_cnt

    grouped_xyz = group_point(xyz, idx)  # (batch_size, npoint, nsample, 3)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(2)  # translation normalization
    if normalize_radius:
        grouped_xyz = grouped_xyz / radius  # Scale normalization
    if orientations is not None:  # :49-54 (note: opposite sign convention to sample_and_group)
        cosval = torch.cos(orientations).unsqueeze(2)
        sinval = torch.sin(orientations).unsqueeze(2)
        grouped_xyz = torch.stack([cosval * grouped_xyz[:, :, :, 0] + sinval * grouped_xyz[:, :, :, 1],
                                   -sinval * grouped_xyz[:, :, :, 0] + cosval * grouped_xyz[:, :, :, 1],
                                   grouped_xyz[:, :, :, 2]], dim=3)

    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_points, idx


def sample_and_group(npoint, radius, nsample, xyz, points, tnet_spec=None, knn=False, use_xyz=True,
                     keypoints=None, orientations=None, normalize_radius=False):
    '''pointnet_common.py:69-135.
    Output:
        new_xyz: (batch_size, npoint, 3), new_points: (batch_size, npoint, nsample, 3+channel),
        idx: (batch_size, npoint, nsample), grouped_xyz: (batch_size, npoint, nsample, 3), end_points
    '''
    end_points = {}
    if tnet_spec is not None:
        raise ValueError("tnet_spec is unused (and undefined) in 3DFeat-Net: pointnet_common.py:122-123")
    if keypoints is not None:
        new_xyz = keypoints
    else:
        new_xyz = gather_point(xyz, farthest_point_sample(npoint, xyz))

    if knn:
        _, idx = knn_point(nsample, xyz, new_xyz)
        pts_cnt = nsample
    else:
        idx, pts_cnt = query_ball_point(radius, nsample, xyz, new_xyz)
    end_points['pts_cnt'] = pts_cnt

    grouped_xyz = group_point(xyz, idx)
    grouped_xyz = grouped_xyz - new_xyz.unsqueeze(2)
    if normalize_radius:
        grouped_xyz = grouped_xyz / radius
    end_points['grouped_xyz_before'] = grouped_xyz

    if orientations is not None:  # :110-120  R = [[c,s,0],[-s,c,0],[0,0,1]];  grouped_xyz @ R
        cosval = torch.cos(orientations)
        sinval = torch.sin(orientations)
        one = torch.ones_like(cosval)
        zero = torch.zeros_like(cosval)
        R = torch.stack([torch.stack([cosval, sinval, zero], dim=-1),
                         torch.stack([-sinval, cosval, zero], dim=-1),
                         torch.stack([zero, zero, one], dim=-1)], dim=-2)  # (B,M,3,3)
        grouped_xyz = torch.matmul(grouped_xyz, R)
        end_points['rotation'] = R

    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    end_points['grouped_xyz'] = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz, end_points


def sample_and_group_all(xyz, points, use_xyz=True):
    '''pointnet_common.py:138-165: one group holding every point, centroid (0,0,0).'''
    batch_size, nsample = xyz.shape[0], xyz.shape[1]
    new_xyz = torch.zeros((batch_size, 1, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.arange(nsample, dtype=torch.int32, device=xyz.device).reshape(1, 1, nsample).repeat(batch_size, 1, 1)
    grouped_xyz = xyz.reshape(batch_size, 1, nsample, 3)
    if points is not None:
        new_points = torch.cat([xyz, points], dim=2) if use_xyz else points
        new_points = new_points.unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz

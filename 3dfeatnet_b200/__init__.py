"""3dfeatnet_b200 -- B200-native (sm_100a) sample-and-group + set-abstraction hot path of 3DFeat-Net.

The host side mirrors the reference's module tree so that it drops in for that path:

    tf_ops.sampling.tf_sampling   farthest_point_sample, gather_point, prob_sample
    tf_ops.grouping.tf_grouping   query_ball_point, query_ball_point2, select_top_k, group_point, knn_point
    models.pointnet_common        sample_points, query_and_group_points, sample_and_group, sample_and_group_all
    models.layers                 conv2d, pairwise_dist, batch_norm_for_{fc,conv1d,conv2d,conv3d}, fully_connected, dropout
    models.feat3dnet              pointnet_sa_module, feature_detection_module, feature_extraction_module, Feat3dNet
    inference                     nms

Because a Python package name cannot start with a digit in an `import` statement, load it with
`importlib.import_module("3dfeatnet_b200")`, or call `install_dropin()` (below) / put this directory on
sys.path to get the reference's own import paths (`from tf_ops.grouping.tf_grouping import query_ball_point`).
All tensors are torch CUDA tensors (fp32 contiguous; indices int32).  Everything calls hand-written CUDA through
the C ABI of include/feat3dnet_b200.h; there is no CPU fallback.
"""
import os
import sys

__version__ = "0.1.0"

_HERE = os.path.dirname(os.path.abspath(__file__))


def install_dropin():
    """Make `tf_ops.*` and `models.*` importable under the reference's own module paths."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)


def lib():
    from . import _lib

    return _lib.lib()

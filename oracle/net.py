"""CPU restatement (PyTorch-CPU, fp32 or fp64) of the 3DFeat-Net detector/descriptor graph, its
loss and its TF-1 Adam step.

TEST INFRASTRUCTURE ONLY (see oracle/ops_oracle.c).  GRAPH PINNED, PRIMITIVES UNPINNED: the reference's
own models/{layers,pointnet_common,feat3dnet}.py are imported unmodified and executed on an eager float64
stand-in for the TensorFlow primitives they call (tests/golden/tf_shim.py, make_golden_net.py ->
tests/golden/ref_net.npz), and tests/test_oracle_net_cpu.py holds this file to those outputs to 1e-9 --
layer order, scopes, BN / activation flags, pooling, concat, rotation, heads and loss are the reference's.
The arithmetic inside the primitives lives in TensorFlow 1.15 + tf.contrib.slim (requirements.txt:45-47),
which is neither vendored under /root/reference nor installable here, and the reference holds no test or
golden vector for it: there the documented semantics of the TF ops are restated,
cited per function, and cross-checked against PyTorch's own library modules (same test file).

Parameters are a flat dict keyed by the TF variable scope names:
    <scope>/conv2d/weights  (Cin, Cout)      slim.conv2d 1x1 kernel, layers.py:32-37
    <scope>/conv2d/biases   (Cout,)
    <scope>/bn/beta|gamma   (Cout,)          layers.py:246-249
    <scope>/bn/moving_mean|moving_variance   the two EMA shadows of layers.py:252-269
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import ops

BN_EPS = 1e-3  # layers.py:271
BN_DECAY = 0.9  # layers.py:251

# (scope, Cin, Cout, has_bn) in graph order -- feat3dnet.py:277-284 (detector), :297-310 (descriptor)
DET_LAYERS = [
    ("detection/conv0", 3, 64, True),
    ("detection/conv1", 64, 128, True),
    ("detection/conv2", 128, 256, True),
    ("detection/conv_post_0", 256, 128, True),
    ("detection/conv_post_1", 128, 64, True),
    ("detection/attention", 64, 1, False),
    ("detection/orientation", 64, 2, False),
]


def desc_layers(feature_dim=32):
    mid = 128 if feature_dim <= 64 else 256  # feat3dnet.py:300
    return [
        ("description/layer1/conv0", 3, 32, True),
        ("description/layer1/conv1", 32, 64, True),
        ("description/layer1/conv_mid_0", 128, mid, True),
        ("description/layer1/conv_post_0", mid, feature_dim, True),
    ]


def init_params(seed=0, feature_dim=32, randomize_bn=False):
    """Random-init state of the TF graph: variance_scaling_initializer() default (factor 2.0, FAN_IN,
    truncated normal => stddev sqrt(1.3*2/fan_in), layers.py:35), zero biases, gamma=1, beta=0, EMA
    shadows zero (ema.apply zero-debias is off for non-Variables => shadows start at 0).  With
    randomize_bn every BN statistic/affine gets a non-trivial value so BN parity is exercised."""
    rng = np.random.default_rng(seed)
    p = {}
    for scope, cin, cout, bn in DET_LAYERS + desc_layers(feature_dim):
        std = math.sqrt(1.3 * 2.0 / cin)
        w = rng.standard_normal((cin, cout)) * std
        w = np.clip(w, -2 * std, 2 * std)  # truncated at 2 sigma
        p[scope + "/conv2d/weights"] = w.astype(np.float32)
        p[scope + "/conv2d/biases"] = (
            rng.standard_normal(cout).astype(np.float32) * 0.1 if randomize_bn else np.zeros(cout, np.float32))
        if bn:
            if randomize_bn:
                p[scope + "/bn/beta"] = (rng.standard_normal(cout) * 0.1).astype(np.float32)
                p[scope + "/bn/gamma"] = (1.0 + 0.2 * rng.standard_normal(cout)).astype(np.float32)
                p[scope + "/bn/moving_mean"] = (rng.standard_normal(cout) * 0.2).astype(np.float32)
                p[scope + "/bn/moving_variance"] = (0.5 + rng.random(cout)).astype(np.float32)
            else:
                p[scope + "/bn/beta"] = np.zeros(cout, np.float32)
                p[scope + "/bn/gamma"] = np.ones(cout, np.float32)
                p[scope + "/bn/moving_mean"] = np.zeros(cout, np.float32)
                p[scope + "/bn/moving_variance"] = np.zeros(cout, np.float32)
    return p


def to_torch(params, dtype=torch.float32, requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v)).to(dtype).clone()
        if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def conv2d(x, P, scope, bn=True, activation="relu", is_training=False, new_stats=None):
    """models/layers.py:11-46: slim 1x1 conv WITH bias -> BN (layers.py:225-272) -> activation.
    x is (..., Cin).  BN: training => moments over all leading axes (population variance), else the
    EMA shadows; y = (x-mean)*rsqrt(var+1e-3)*gamma + beta (tf.nn.batch_normalization)."""
    y = x @ P[scope + "/conv2d/weights"] + P[scope + "/conv2d/biases"]
    if bn:
        if is_training:
            flat = y.reshape(-1, y.shape[-1])
            mean = flat.mean(0)
            var = flat.var(0, unbiased=False)
            if new_stats is not None:
                mm, mv = P[scope + "/bn/moving_mean"], P[scope + "/bn/moving_variance"]
                new_stats[scope + "/bn/moving_mean"] = (mm - (1 - BN_DECAY) * (mm - mean)).detach()
                new_stats[scope + "/bn/moving_variance"] = (mv - (1 - BN_DECAY) * (mv - var)).detach()
        else:
            mean, var = P[scope + "/bn/moving_mean"], P[scope + "/bn/moving_variance"]
        inv = torch.rsqrt(var + BN_EPS) * P[scope + "/bn/gamma"]
        y = y * inv + (P[scope + "/bn/beta"] - mean * inv)
    if activation == "relu":
        y = torch.relu(y)
    elif activation == "softplus":
        y = F.softplus(y)
    return y


def l2_normalize(x, dim, eps):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), eps))."""
    ss = (x * x).sum(dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=eps))


def group_normalised(xyz_np, new_xyz_np, radius, nsample, dtype):
    """pointnet_common.py:32-47 / :96-107: ball query (oracle), group, translate, divide by radius."""
    idx, cnt = ops.query_ball_point(radius, nsample, xyz_np, new_xyz_np)
    grouped = ops.group_point(xyz_np, idx)  # (B,M,S,3) fp32 gather, exact
    g = torch.as_tensor(grouped).to(dtype) - torch.as_tensor(new_xyz_np).to(dtype)[:, :, None, :]
    g = g / torch.tensor(radius, dtype=torch.float32).to(dtype)
    return g, idx, cnt


def detector(xyz_np, P, num_clusters, radius, nsample, is_training=False, keypoints_np=None, new_stats=None,
             dtype=torch.float32):
    """feature_detection_module, models/feat3dnet.py:90-151 (compute_det_gradients=False semantics)."""
    xyz_np = np.ascontiguousarray(xyz_np[:, :, :3], np.float32)
    if keypoints_np is not None:  # inference.py feeds end_points['keypoints'] directly (:128-131)
        new_xyz = np.ascontiguousarray(keypoints_np, np.float32)
        fps_idx = None
    elif num_clusters <= 0:  # pointnet_common.py:24-25
        new_xyz, fps_idx = xyz_np.copy(), None
    else:
        fps_idx = ops.farthest_point_sample(num_clusters, xyz_np)
        new_xyz = ops.gather_point(xyz_np, fps_idx)
    g, idx, cnt = group_normalised(xyz_np, new_xyz, radius, nsample, dtype)
    h = g
    for i in range(3):
        h = conv2d(h, P, "detection/conv%d" % i, True, "relu", is_training, new_stats)
    h = h.amax(dim=2, keepdim=True)  # tf.reduce_max, :130 (amax splits grads equally among ties, like TF)
    for i in range(2):
        h = conv2d(h, P, "detection/conv_post_%d" % i, True, "relu", is_training, new_stats)
    att = conv2d(h, P, "detection/attention", False, "softplus")[:, :, 0, 0]
    oxy_raw = conv2d(h, P, "detection/orientation", False, None)[:, :, 0, :]
    oxy = l2_normalize(oxy_raw, 2, 1e-8)
    ori = torch.atan2(oxy[:, :, 1], oxy[:, :, 0])
    # orientation_xy: the head output BEFORE l2_normalize / atan2 -- its norm is the conditioning of the angle (oracle/parity.py)
    return dict(new_xyz=new_xyz, fps_idx=fps_idx, idx=idx, pts_cnt=cnt, attention=att, orientation=ori,
                orientation_xy=oxy_raw, grouped_xyz=g)


def descriptor(xyz_np, P, keypoints_np, orientation, radius, nsample, feature_dim=32, is_training=False,
               new_stats=None, dtype=torch.float32):
    """feature_extraction_module / pointnet_sa_module, models/feat3dnet.py:9-87,154-187, with
    sample_and_group (pointnet_common.py:69-135): keypoints given, rotation by `orientation`."""
    xyz_np = np.ascontiguousarray(xyz_np[:, :, :3], np.float32)
    g, idx, cnt = group_normalised(xyz_np, np.ascontiguousarray(keypoints_np, np.float32), radius, nsample, dtype)
    if orientation is not None:  # pointnet_common.py:110-120: x' = x c - y s ; y' = x s + y c
        c, s = torch.cos(orientation)[:, :, None], torch.sin(orientation)[:, :, None]
        g = torch.stack([g[..., 0] * c - g[..., 1] * s, g[..., 0] * s + g[..., 1] * c, g[..., 2]], dim=-1)
    sc = "description/layer1/"
    h = g
    for i in range(2):
        h = conv2d(h, P, sc + "conv%d" % i, True, "relu", is_training, new_stats)
    pooled = h.amax(dim=2, keepdim=True)
    h = torch.cat([h, pooled.expand(-1, -1, h.shape[2], -1)], dim=3)  # :60-64
    h = conv2d(h, P, sc + "conv_mid_0", True, None, is_training, new_stats)  # final_relu=False, :67-72
    h = h.amax(dim=2, keepdim=True)  # :75
    h = conv2d(h, P, sc + "conv_post_0", True, None, is_training, new_stats)  # :79-84
    feat = l2_normalize(h[:, :, 0, :], 2, 1e-8)  # :185
    return dict(features=feat, idx=idx, pts_cnt=cnt, rotated_xyz=g)


def inference_model(xyz_np, P, num_clusters=512, radius=2.0, nsample=64, feature_dim=32, no_regress=False,
                    is_training=False, keypoints_np=None, new_stats=None, dtype=torch.float32):
    """Feat3dNet.get_inference_model, models/feat3dnet.py:258-313."""
    det = detector(xyz_np, P, num_clusters, radius, nsample, is_training, keypoints_np, new_stats, dtype)
    ori = None if no_regress else det["orientation"]
    desc = descriptor(xyz_np, P, det["new_xyz"], ori, radius, nsample, feature_dim, is_training, new_stats, dtype)
    return dict(xyz=det["new_xyz"], features=desc["features"], attention=det["attention"],
                orientation=det["orientation"], orientation_xy=det["orientation_xy"], idx=det["idx"], pts_cnt=det["pts_cnt"],
                fps_idx=det["fps_idx"])


def pairwise_dist(A, B):
    """models/layers.py:49-62."""
    return ((A[:, :, None, :] - B[:, None, :, :]) ** 2).sum(3)


def triplet_loss(fa, fp, fn, att_a, margin=0.2, use_attention=True):
    """Feat3dNet.get_loss, models/feat3dnet.py:315-357."""
    best_p = pairwise_dist(fa, fp).amin(dim=2)
    best_n = pairwise_dist(fa, fn).amin(dim=2)
    if not use_attention:
        sp, sn = best_p.mean(1), best_n.mean(1)
    else:
        w = att_a / att_a.sum(1, keepdim=True)
        sp, sn = (w * best_p).sum(1), (w * best_n).sum(1)
    return torch.clamp(sp - sn + margin, min=0).mean()


def adam_step(P, grads, state, lr=1e-5, b1=0.9, b2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (feat3dnet.py:359-375): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; theta -= lr_t * m / (sqrt(v) + eps)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    for k, g in grads.items():
        m = state.setdefault("m/" + k, torch.zeros_like(g))
        v = state.setdefault("v/" + k, torch.zeros_like(g))
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        with torch.no_grad():
            P[k] -= lr_t * m / (v.sqrt() + eps)


def train_step(anchors, positives, negatives, P, state, num_clusters=512, radius=2.0, nsample=64, feature_dim=32,
               margin=0.2, attention=True, no_regress=False, lr=1e-5, dtype=torch.float32):
    """get_train_model (feat3dnet.py:227-256) + get_loss + get_train_op on one triplet batch.
    P: dict of torch tensors (to_torch(..., requires_grad=True)); updated in place.  Returns loss, grads."""
    clouds = np.concatenate([anchors, positives, negatives], axis=0)
    new_stats = {}
    out = inference_model(clouds, P, num_clusters, radius, nsample, feature_dim, no_regress, True, None, new_stats,
                          dtype)
    fa, fp, fn = torch.chunk(out["features"], 3, dim=0)
    att_a = torch.chunk(out["attention"], 3, dim=0)[0]
    loss = triplet_loss(fa, fp, fn, att_a, margin, attention)
    names = [k for k, v in P.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [P[k] for k in names], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(P[k])) for k, g in zip(names, gs)}
    adam_step(P, grads, state, lr)
    with torch.no_grad():
        for k, v in new_stats.items():
            P[k].copy_(v)
    return loss.detach(), grads, out

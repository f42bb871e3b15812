"""3DFeat-Net detector + descriptor -- the reference's models/feat3dnet.py:9-375 on torch CUDA tensors.

Same public names and argument meaning as the reference (`pointnet_sa_module`, `feature_detection_module`,
`feature_extraction_module`, `Feat3dNet.{get_placeholders,get_train_model,get_inference_model,get_loss,
get_train_op}`); TF placeholders/sessions disappear (functions run eagerly on tensors), `is_training` is a Python
bool, TF variable scopes are key prefixes of one flat parameter dict (see models/layers.py).

Eval-mode forwards (`is_training=False`) run the FUSED CUDA path: FPS -> gather -> ball query (once, shared by
detector and descriptor: pointnet_common.py:39 and :102 issue the identical query) -> f3d_detector_forward ->
f3d_descriptor_forward.  Training-mode forwards use the unfused differentiable layers of models/layers.py on top
of the same CUDA sampling/grouping operators.
"""
import ctypes
import importlib
import logging
import math

import numpy as np
import torch

_ROOT = __name__.split(".")[0]
_pfx = "3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else ""
_lib = importlib.import_module(_pfx + "_lib")
_layers = importlib.import_module(_pfx + "models.layers")
_pc = importlib.import_module(_pfx + "models.pointnet_common")
conv2d, pairwise_dist = _layers.conv2d, _layers.pairwise_dist
sample_points, sample_and_group = _pc.sample_points, _pc.sample_and_group
sample_and_group_all, query_and_group_points = _pc.sample_and_group_all, _pc.query_and_group_points

PRECISIONS = {"fp32": 0, "bf16x3": 2}

# (scope, Cin, Cout, has_bn) -- feat3dnet.py:277-284 (detector) and :297-310 (descriptor)
DET_LAYERS = [
    ("detection/conv0", 3, 64, True), ("detection/conv1", 64, 128, True), ("detection/conv2", 128, 256, True),
    ("detection/conv_post_0", 256, 128, True), ("detection/conv_post_1", 128, 64, True),
    ("detection/attention", 64, 1, False), ("detection/orientation", 64, 2, False),
]


def desc_layers(feature_dim=32):
    mid = 128 if feature_dim <= 64 else 256  # feat3dnet.py:300
    return [("description/layer1/conv0", 3, 32, True), ("description/layer1/conv1", 32, 64, True),
            ("description/layer1/conv_mid_0", 128, mid, True), ("description/layer1/conv_post_0", mid, feature_dim, True)]


def init_params(seed=0, feature_dim=32, device="cuda"):
    """Random-init state of the TF graph: slim.variance_scaling_initializer() (factor 2, FAN_IN, truncated normal,
    layers.py:35), zero biases, gamma=1, beta=0, EMA shadows 0.  Returns {name: torch tensor on `device`}."""
    rng = np.random.default_rng(seed)
    p = {}
    for scope, cin, cout, bn in DET_LAYERS + desc_layers(feature_dim):
        std = math.sqrt(1.3 * 2.0 / cin)
        w = np.clip(rng.standard_normal((cin, cout)) * std, -2 * std, 2 * std).astype(np.float32)
        p[scope + "/conv2d/weights"] = w
        p[scope + "/conv2d/biases"] = np.zeros(cout, np.float32)
        if bn:
            p[scope + "/bn/beta"] = np.zeros(cout, np.float32)
            p[scope + "/bn/gamma"] = np.ones(cout, np.float32)
            p[scope + "/bn/moving_mean"] = np.zeros(cout, np.float32)
            p[scope + "/bn/moving_variance"] = np.zeros(cout, np.float32)
    return {k: torch.as_tensor(v).to(device) for k, v in p.items()}


def params_to_device(params, device="cuda", requires_grad=False):
    out = {}
    for k, v in params.items():
        t = torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).to(device=device, dtype=torch.float32).clone()
        if requires_grad and not k.endswith(("moving_mean", "moving_variance")):
            t.requires_grad_(True)
        out[k] = t
    return out


def fold_params(params, feature_dim=32):
    """Fold every conv+BN pair (eval mode) into W' = W*s, b' = (b-mean)*s + beta, s = gamma*rsqrt(var+1e-3)
    (layers.py:244-272) and pack all layers into the single device buffer the fused kernels read
    (csrc/weights_layout.h; offsets come from the library itself)."""
    L = _lib.lib()
    nb = L.f3d_packed_weights_num_blocks()
    offs = (ctypes.c_int * nb)()
    sizes = (ctypes.c_int * nb)()
    _lib.check(L.f3d_packed_weights_offsets(feature_dim, offs, sizes), "packed_weights_offsets")
    total = L.f3d_packed_weights_floats(feature_dim)
    any_t = next(iter(params.values()))
    packed = torch.zeros((total,), dtype=torch.float32, device=any_t.device)
    layers = DET_LAYERS + desc_layers(feature_dim)
    if nb != 2 * len(layers):
        raise _lib.F3DError("weight layout mismatch between host (%d blocks) and library (%d)" % (2 * len(layers), nb))
    for i, (scope, cin, cout, bn) in enumerate(layers):
        w = params[scope + "/conv2d/weights"].detach().to(torch.float64).reshape(cin, cout)
        b = params[scope + "/conv2d/biases"].detach().to(torch.float64)
        if bn:
            s = params[scope + "/bn/gamma"].detach().to(torch.float64) * torch.rsqrt(
                params[scope + "/bn/moving_variance"].detach().to(torch.float64) + _layers.BN_EPS)
            w = w * s
            b = (b - params[scope + "/bn/moving_mean"].detach().to(torch.float64)) * s + \
                params[scope + "/bn/beta"].detach().to(torch.float64)
        if sizes[2 * i] != cin * cout or sizes[2 * i + 1] != cout:
            raise _lib.F3DError("weight layout mismatch at %s" % scope)
        packed[offs[2 * i]:offs[2 * i] + cin * cout] = w.to(torch.float32).reshape(-1)
        packed[offs[2 * i + 1]:offs[2 * i + 1] + cout] = b.to(torch.float32)
    return packed


def _workspace(b, m, feature_dim, device):
    nbytes = _lib.lib().f3d_forward_workspace_bytes(b, m, feature_dim)
    return torch.empty((nbytes,), dtype=torch.uint8, device=device), nbytes


def detector_forward_fused(xyz, new_xyz, idx, radius, packed, precision="fp32"):
    """f3d_detector_forward: (attention (B,M), orientation (B,M)) for given centres and ball-query indices."""
    _lib.require_cuda(xyz, new_xyz, idx, packed)
    b, n, _ = xyz.shape
    m, ns = idx.shape[1], idx.shape[2]
    att = torch.empty((b, m), dtype=torch.float32, device=xyz.device)
    ori = torch.empty((b, m), dtype=torch.float32, device=xyz.device)
    ws, nbytes = _workspace(b, m, 32, xyz.device)
    _lib.check(_lib.lib().f3d_detector_forward(b, n, m, ns, float(radius), _lib.ptr(xyz), _lib.ptr(new_xyz), _lib.ptr(idx),
                                               _lib.ptr(packed), _lib.ptr(att), _lib.ptr(ori), PRECISIONS[precision],
                                               _lib.ptr(ws), nbytes, _lib.stream()), "detector_forward")
    return att, ori


def descriptor_forward_fused(xyz, new_xyz, idx, orientation, radius, packed, feature_dim=32, precision="fp32"):
    """f3d_descriptor_forward: l2-normalised features (B,M,feature_dim)."""
    _lib.require_cuda(xyz, new_xyz, idx, packed, orientation)
    b, n, _ = xyz.shape
    m, ns = idx.shape[1], idx.shape[2]
    feat = torch.empty((b, m, feature_dim), dtype=torch.float32, device=xyz.device)
    ws, nbytes = _workspace(b, m, feature_dim, xyz.device)
    _lib.check(_lib.lib().f3d_descriptor_forward(b, n, m, ns, float(radius), feature_dim, _lib.ptr(xyz), _lib.ptr(new_xyz),
                                                 _lib.ptr(idx), _lib.ptr(orientation), _lib.ptr(packed), _lib.ptr(feat),
                                                 PRECISIONS[precision], _lib.ptr(ws), nbytes, _lib.stream()),
               "descriptor_forward")
    return feat


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, mlp3, is_training, scope, bn=True, bn_decay=None,
                       tnet_spec=None, knn=False, use_xyz=True,
                       keypoints=None, orientations=None, normalize_radius=True, final_relu=True,
                       params=None, new_stats=None, neighbours=None, after_mid_hook=None):
    """ PointNet Set Abstraction (SA) Module (feat3dnet.py:9-87), unfused differentiable statement.

    Returns:
        new_xyz: (batch_size, npoint, 3), new_points: (batch_size, npoint, mlp3[-1]), idx, end_points
    """
    if npoint is None:
        nsample = xyz.shape[1]
        new_xyz, new_points, idx, grouped_xyz = sample_and_group_all(xyz, points, use_xyz)
        end_points = {}
    else:
        new_xyz, new_points, idx, grouped_xyz, end_points = sample_and_group(
            npoint, radius, nsample, xyz, points, tnet_spec, knn, use_xyz, keypoints=keypoints,
            orientations=orientations, normalize_radius=normalize_radius, neighbours=neighbours)

    # training on the tensor cores: the activations of the chain conv0 -> conv1 -> conv_mid stay unmaterialised (layers.DeferredActivation);
    # the last layer of `mlp` feeds the pool AND conv_mid, so it hands back its pooled maximum next to the deferred activation
    chain = bool(is_training) and bn and mlp and mlp2 and new_points.is_cuda and new_points.shape[-1] == 3 and _layers.chain_supported() \
        and all(c % 16 == 0 and c & (c - 1) == 0 and c <= 128 for c in mlp)
    pooled = None
    for i, num_out_channel in enumerate(mlp):
        last = i == len(mlp) - 1
        new_points = conv2d(new_points, num_out_channel, [1, 1], stride=[1, 1], padding='VALID', bn=bn,
                            is_training=is_training, scope=scope + '/conv%d' % i, params=params, new_stats=new_stats,
                            defer=chain, also_pool=chain and last)
        if chain and last:
            new_points, pooled = new_points

    if pooled is None:
        pooled = _layers.max_pool_samples(new_points)  # tf.reduce_max; the gradient is shared among ties like TF
    # the reference tiles `pooled` over the samples and concatenates it to new_points (:60-66); conv2d(concat_pooled=) states
    # the same layer without building the (B,M,S,2C) tensor
    for i, num_out_channel in enumerate(mlp2 or []):
        act = _layers.relu if (final_relu or i < len(mlp2) - 1) else None
        new_points = conv2d(new_points, num_out_channel, [1, 1], padding='VALID', stride=[1, 1], bn=bn,
                            is_training=is_training, scope=scope + '/conv_mid_%d' % i, bn_decay=bn_decay, activation=act,
                            params=params, new_stats=new_stats, pool_samples=(i == len(mlp2) - 1),
                            concat_pooled=pooled if i == 0 else None)
    if not mlp2:
        new_points = torch.cat((new_points, pooled.expand(-1, -1, new_points.shape[2], -1)), dim=3)
        new_points = _layers.max_pool_samples(new_points)
    if after_mid_hook is not None:
        after_mid_hook()  # the last per-point layer has been launched: only per-cluster work follows until the backward returns to it

    for i, num_out_channel in enumerate(mlp3 or []):
        act = _layers.relu if (final_relu or i < len(mlp3) - 1) else None
        new_points = conv2d(new_points, num_out_channel, [1, 1], padding='VALID', stride=[1, 1], bn=bn,
                            is_training=is_training, scope=scope + '/conv_post_%d' % i, bn_decay=bn_decay, activation=act,
                            params=params, new_stats=new_stats)
    new_points = new_points.squeeze(2)
    return new_xyz, new_points, idx, end_points


def feature_detection_module(xyz, points, num_clusters, radius, is_training, mlp, mlp2, num_samples=64, use_bn=True,
                             compute_det_gradients=False, params=None, new_stats=None, scope="detection",
                             keypoints=None, neighbours=None):
    """ Detect features in point cloud (feat3dnet.py:90-151), unfused differentiable statement.

    compute_det_gradients: the reference's default (True) raises KeyError while building its graph
    (feat3dnet.py:112,125-127); the intended per-layer saliency sum(y * dy/dxyz) is implemented here with autograd
    through the GroupPoint / GatherPoint gradients and returned in end_points['gradients']['det'].

    Returns: new_xyz, idx, attention, orientation, end_points
    """
    end_points = {}
    if compute_det_gradients:
        xyz = xyz.detach().requires_grad_(True)
        end_points['gradients'] = {'det': {}}
    new_xyz = sample_points(xyz, num_clusters) if keypoints is None else keypoints
    new_points, idx = query_and_group_points(xyz, points, new_xyz, num_samples, radius, knn=False, use_xyz=True,
                                             normalize_radius=True, orientations=None, end_points=end_points, neighbours=neighbours)

    # training on the tensor cores: the activations between the layers of the chain stay unmaterialised (layers.DeferredActivation)
    chain = bool(is_training) and use_bn and not compute_det_gradients and new_points.is_cuda and new_points.shape[-1] == 3 \
        and _layers.chain_supported() and all(c % 16 == 0 and c & (c - 1) == 0 for c in mlp) and all(c <= 128 for c in mlp[:-1])
    for i, num_out_channel in enumerate(mlp):
        # the last layer's activation only feeds the max-pool: conv2d(pool_samples=True) pools in the same call
        pool_here = (i == len(mlp) - 1) and not compute_det_gradients
        new_points = conv2d(new_points, num_out_channel, [1, 1], stride=[1, 1], padding='VALID', bn=use_bn,
                            is_training=is_training, scope=scope + '/conv%d' % i, params=params, new_stats=new_stats,
                            pool_samples=pool_here, defer=chain and not pool_here)
        if compute_det_gradients:
            (g,) = torch.autograd.grad(new_points, xyz, grad_outputs=new_points.detach(), retain_graph=True)
            end_points['gradients']['det']['mlp_{}'.format(i)] = g
    if compute_det_gradients or not mlp:
        new_points = _layers.max_pool_samples(new_points)

    for i, num_out_channel in enumerate(mlp2 or []):
        new_points = conv2d(new_points, num_out_channel, [1, 1], padding='VALID', stride=[1, 1], bn=use_bn,
                            is_training=is_training, scope=scope + '/conv_post_%d' % i, params=params,
                            new_stats=new_stats)

    fused = _layers.detector_heads(new_points, params, scope) if (is_training and not compute_det_gradients) else None
    if fused is not None:  # both heads, forward and backward, as one CUDA op each (csrc/train.cu)
        return new_xyz, idx, fused[0], fused[1], end_points
    attention = conv2d(new_points, 1, [1, 1], stride=[1, 1], padding='VALID', activation=_layers.softplus, bn=False,
                       scope=scope + '/attention', params=params)
    attention = attention.squeeze(3).squeeze(2)

    orientation_xy = conv2d(new_points, 2, [1, 1], stride=[1, 1], padding='VALID', activation=None, bn=False,
                            scope=scope + '/orientation', params=params)
    orientation_xy = orientation_xy.squeeze(2)
    ss = (orientation_xy * orientation_xy).sum(2, keepdim=True)
    orientation_xy = orientation_xy * torch.rsqrt(torch.clamp(ss, min=1e-8))  # tf.nn.l2_normalize(eps=1e-8)
    orientation = torch.atan2(orientation_xy[:, :, 1], orientation_xy[:, :, 0])
    return new_xyz, idx, attention, orientation, end_points


def feature_extraction_module(l0_xyz, l0_points, is_training, mlp, mlp2, mlp3, keypoints, orientations, radius=2.0,
                              num_samples=64, use_bn=True, params=None, new_stats=None, scope="description", neighbours=None,
                              after_mid_hook=None):
    """ Extract feature descriptors (feat3dnet.py:154-187), unfused differentiable statement.
    neighbours: optional (idx, pts_cnt) of the ball query the detector ran on the same (l0_xyz, keypoints, radius, num_samples)
    -- the reference issues that query twice (pointnet_common.py:39 and :102); handing it over runs it once, with no hidden state.
    Returns: xyz, features, end_points """
    l1_xyz, l1_points, l1_idx, end_points = pointnet_sa_module(
        l0_xyz, l0_points, 512, radius, num_samples, mlp=mlp, mlp2=mlp2, mlp3=mlp3, is_training=is_training,
        scope=scope + '/layer1', bn=use_bn, bn_decay=None, keypoints=keypoints, orientations=orientations,
        normalize_radius=True, final_relu=False, params=params, new_stats=new_stats, neighbours=neighbours, after_mid_hook=after_mid_hook)
    ss = (l1_points * l1_points).sum(2, keepdim=True)
    features = l1_points * torch.rsqrt(torch.clamp(ss, min=1e-8))
    return l1_xyz, features, end_points


class Feat3dNet:

    DEFAULT_PARAM = {'NoRegress': False, 'BaseScale': 2.0, 'Attention': True, 'num_clusters': 512, 'num_samples': 64,
                     'margin': 0.2, 'feature_dim': 32, 'freeze_scopes': None}

    def __init__(self, param=None, weights=None, device="cuda", precision="fp32", seed=0):
        """ Constructor (feat3dnet.py:192-209).

        param: dict with the reference's keys ('NoRegress', 'BaseScale', 'Attention', 'num_clusters',
               'num_samples', 'margin', 'feature_dim', 'freeze_scopes').
        weights: flat {TF variable name: array}; random-init state when None.
        precision: 'fp32' (CUDA-core FFMA) or 'bf16x3' (tcgen05 tensor cores, split-bf16) for the fused eval-mode kernels.
        """
        self.logger = logging.getLogger(self.__class__.__name__)
        self.param = dict(self.DEFAULT_PARAM)
        self.param.update(param or {})
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        self.precision = precision
        self.device = torch.device(device)
        if weights is None:
            self.weights = init_params(seed, self.param['feature_dim'], self.device)
        else:
            self.weights = params_to_device(weights, self.device)
        self._packed = None
        self._adam = None
        self.logger.info('Model parameters: %s', self.param)

    # -- parameters --------------------------------------------------------------------------------
    def invalidate(self):
        """Call after changing self.weights in place: the folded/packed copy is rebuilt on the next forward."""
        self._packed = None

    def packed_weights(self):
        if self._packed is None:
            self._packed = fold_params(self.weights, self.param['feature_dim'])
        return self._packed

    def trainable_variables(self):
        ex = tuple(self.param.get('freeze_scopes') or ())
        return {k: v for k, v in self.weights.items()
                if not k.endswith(("moving_mean", "moving_variance")) and not (ex and k.startswith(ex))}

    def get_placeholders(self, data_dim):
        """TF placeholders have no counterpart: inputs are passed as tensors (feat3dnet.py:211-225)."""
        return None, None, None

    # -- graphs ------------------------------------------------------------------------------------
    def sample_clusters(self, point_clouds):
        """(keypoints, idx, pts_cnt) of the model's clusters for a batch of clouds: farthest point sampling + ball query, exactly what
        get_inference_model computes at its start.  Handing the result back as `presampled=` skips that part of the forward (the pipelined
        training step computes it for the next batch beside the backward pass of the current one)."""
        with torch.no_grad():
            l0_xyz = point_clouds[:, :, :3].contiguous()
            kp = sample_points(l0_xyz, self.param['num_clusters'])
            idx, pts_cnt = _pc.query_ball_point(self.param['BaseScale'], self.param['num_samples'], l0_xyz, kp)
        return kp, idx, pts_cnt

    def get_train_model(self, anchors, positives, negatives, is_training, use_bn=True, presampled=None, after_mid_hook=None):
        """feat3dnet.py:227-256: concatenates the triplet on the batch axis and calls get_inference_model.
        presampled: sample_clusters() of the concatenated batch (optional); after_mid_hook: see capture_train_step(pipelined=True)."""
        end_points = {}
        point_clouds = torch.cat([anchors, positives, negatives], dim=0)
        end_points['input_pointclouds'] = point_clouds
        xyz, features, attention, endpoints_temp = self.get_inference_model(point_clouds, is_training, use_bn, presampled=presampled,
                                                                            after_mid_hook=after_mid_hook)
        end_points['output_xyz'] = xyz
        end_points['output_features'] = features
        end_points.update(endpoints_temp)
        xyz = torch.chunk(xyz, 3, dim=0)
        features = torch.chunk(features, 3, dim=0)
        anchor_attention = torch.chunk(attention, 3, dim=0)[0] if attention is not None else None
        return xyz, features, anchor_attention, end_points

    def get_inference_model(self, point_cloud, is_training, use_bn=True, compute_det_gradients=False, keypoints=None,
                            fetch_features=True, presampled=None, after_mid_hook=None, ball_grid=None):
        """ The core 3DFeat-Net model (feat3dnet.py:258-313).

        point_cloud: (B,N,>=3) CUDA float32.  keypoints: optional (B,M,3) cluster centres (what inference.py feeds
        through end_points['keypoints'], :128-131); otherwise FPS picks param['num_clusters'] (or every point when <=0).
        fetch_features=False is `sess.run([xyz_op, attention_op])` (inference.py:128): TF prunes the descriptor sub-graph
        when it is not fetched, and so does the fused eval path (features come back as None).
        Returns: xyz, features, attention, end_points
        """
        _lib.require_cuda(point_cloud)
        l0_xyz = point_cloud[:, :, :3].contiguous()
        radius, ns = self.param['BaseScale'], self.param['num_samples']
        fdim = self.param['feature_dim']
        end_points = {}
        if not is_training and use_bn and not compute_det_gradients:
            with torch.no_grad():
                kp = sample_points(l0_xyz, self.param['num_clusters']) if keypoints is None else keypoints.contiguous()
                # ball_grid: a tf_grouping.BallGrid of l0_xyz (same tensor, same radius), for callers that query one cloud repeatedly
                idx, pts_cnt = _pc.query_ball_point(radius, ns, l0_xyz, kp, grid=ball_grid) if ball_grid is not None else \
                    _pc.query_ball_point(radius, ns, l0_xyz, kp)
                packed = self.packed_weights()
                attention, orientation = detector_forward_fused(l0_xyz, kp, idx, radius, packed, self.precision)
                ori = None if self.param['NoRegress'] else orientation
                features = descriptor_forward_fused(l0_xyz, kp, idx, ori, radius, packed, fdim, self.precision) if fetch_features else None
            end_points.update(keypoints=kp, attention=attention, orientation=orientation, idx=idx, pts_cnt=pts_cnt)
            return kp, features, (attention if self.param['Attention'] else None), end_points

        new_stats = {} if is_training else None
        neighbours = None
        if presampled is not None:  # (keypoints, idx, pts_cnt) of sample_clusters() on this batch
            keypoints, neighbours = presampled[0], (presampled[1], presampled[2])
        kp, idx, attention, orientation, ep = feature_detection_module(
            l0_xyz, None, self.param['num_clusters'], radius, is_training, [64, 128, 256], [128, 64], num_samples=ns,
            use_bn=use_bn, compute_det_gradients=compute_det_gradients, params=self.weights, new_stats=new_stats,
            keypoints=keypoints, neighbours=neighbours)
        end_points.update(ep)
        end_points.update(keypoints=kp, attention=attention, orientation=orientation, idx=idx)
        keypoint_orientation = None if self.param['NoRegress'] else orientation
        mlp, mlp2, mlp3 = [32, 64], ([128] if fdim <= 64 else [256]), [fdim]
        xyz, features, ep2 = feature_extraction_module(
            l0_xyz, None, is_training, mlp, mlp2, mlp3, keypoints=kp, orientations=keypoint_orientation, radius=radius,
            num_samples=ns, use_bn=use_bn, params=self.weights, new_stats=new_stats, neighbours=(idx, ep['pts_cnt']),
            after_mid_hook=after_mid_hook)
        end_points.update(ep2)
        end_points['bn_updates'] = new_stats
        return xyz, features, (attention if self.param['Attention'] else None), end_points

    def get_loss(self, xyz, features, anchor_attention, end_points):
        """ Attention weighted alignment loss (feat3dnet.py:315-357). """
        anchors, positives, negatives = features
        if self.param.get('fused_loss', True) and anchors.is_cuda:
            # one CUDA op for the forward and the backward (csrc/train.cu); the per-cloud terms the reference only
            # logs (sum_positive / sum_negative) are produced by the unfused statement below (fused_loss=False)
            if self.param['Attention']:
                end_points['normalized_attention'] = (anchor_attention / anchor_attention.sum(dim=1)[:, None]).detach()
            return _layers.triplet_loss(anchors, positives, negatives, anchor_attention if self.param['Attention'] else None,
                                        self.param['margin']), end_points
        best_positive = pairwise_dist(anchors, positives).amin(dim=2)
        best_negative = pairwise_dist(anchors, negatives).amin(dim=2)
        if not self.param['Attention']:
            sum_positive = best_positive.mean(1)
            sum_negative = best_negative.mean(1)
        else:
            attention_sm = anchor_attention / anchor_attention.sum(dim=1)[:, None]
            sum_positive = (attention_sm * best_positive).sum(1)
            sum_negative = (attention_sm * best_negative).sum(1)
            end_points['normalized_attention'] = attention_sm
        end_points['sum_positive'] = sum_positive
        end_points['sum_negative'] = sum_negative
        triplet_cost = torch.clamp(sum_positive - sum_negative + self.param['margin'], min=0.)
        return triplet_cost.mean(), end_points

    def get_train_op(self, loss_op, lr=1e-5, global_step=None, end_points=None, grad_hook=None, grad_scale=1.0):
        """ One optimiser step (feat3dnet.py:359-375): TF-1 Adam, lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
        theta -= lr_t*m/(sqrt(v)+1e-8), over the trainable variables minus param['freeze_scopes'].
        grad_hook(flat_grad) runs between backward and the update (the data-parallel all-reduce goes there); grad_scale
        multiplies the gradient inside the Adam kernel (1/world_size after a SUM all-reduce).
        Also applies the BN EMA updates collected by the forward (end_points['bn_updates'])."""
        var = self.trainable_variables()
        names = [k for k, v in var.items() if v.requires_grad]
        if not names:
            raise _lib.F3DError("get_train_op: no variable requires grad; build the model with train_mode()")
        grads = torch.autograd.grad(loss_op, [var[k] for k in names], allow_unused=True)
        flat = torch.cat([(g if g is not None else torch.zeros_like(var[k])).reshape(-1) for k, g in zip(names, grads)])
        if self._adam is None or self._adam["names"] != names:
            # flat first/second moments + a device table of {param*, grad*, m*, v*, n} records: ONE launch per update
            offs = np.cumsum([0] + [var[k].numel() for k in names])
            mflat, vflat = torch.zeros_like(flat), torch.zeros_like(flat)
            gbuf = torch.empty_like(flat)
            rec = np.zeros((len(names), 5), np.int64)
            for r, k in enumerate(names):
                if not var[k].is_contiguous():
                    raise _lib.F3DError("get_train_op: variable %s is not contiguous" % k)
                rec[r] = (var[k].data_ptr(), gbuf.data_ptr() + 4 * int(offs[r]), mflat.data_ptr() + 4 * int(offs[r]),
                          vflat.data_ptr() + 4 * int(offs[r]), var[k].numel())
            self._adam = {"t": 0, "names": names, "m": mflat, "v": vflat, "g": gbuf, "records": torch.as_tensor(rec).to(flat.device),
                          "max_n": int(rec[:, 4].max()), "ptrs": [var[k].data_ptr() for k in names],
                          "t_dev": torch.zeros(1, dtype=torch.int64, device=flat.device)}  # the update count lives on the device
        st = self._adam
        if st["ptrs"] != [var[k].data_ptr() for k in names]:
            raise _lib.F3DError("get_train_op: a variable was re-allocated since the optimiser state was built")
        if grad_hook is not None:
            flat = grad_hook(flat)
        st["g"].copy_(flat)
        st["t"] += 1
        _lib.require_cuda(st["g"])
        with torch.no_grad():
            _lib.check(_lib.lib().f3d_adam_step(len(names), _lib.ptr(st["records"]), st["max_n"], float(lr), 0.9, 0.999, 1e-8,
                                                st["t"], float(grad_scale), _lib.ptr(st["t_dev"]), _lib.stream()), "adam_step")
            if end_points is not None and end_points.get('bn_updates'):
                _layers.apply_ema_updates(self.weights, end_points['bn_updates'])
        self.invalidate()
        return flat

    def capture_train_step(self, anchors, positives, negatives, lr=1e-5, grad_hook=None, grad_scale=1.0, warmup=3, pipelined=False):
        """Capture one whole training step (get_train_model -> get_loss -> get_train_op: ~180 launches) into a CUDA graph.
        Returns replay(anchors=None, positives=None, negatives=None) -> loss tensor (device, updated by every replay); new
        triplets are copied into the static input buffers first.  The Adam step count lives on the device, so a replay is
        a real optimiser step.  `warmup` eager steps run first (they are real steps too).

        pipelined=True: the SOFTWARE-PIPELINED step.  Farthest point sampling + ball query (0.2 ms at C4: 18 CTAs busy, 130 SMs idle)
        depend on the input clouds only, so the clusters of the NEXT batch are computed on a side stream beside the per-cluster part of
        the current step (forked after the forward's last per-point layer, where tails, loss and their gradients leave most SMs idle;
        joined at the end of the step).
        replay(a, p, n) then takes the triplets of the step AFTER the one it runs: it runs the step on the batch staged by the previous
        call (the first one: the batch given here) and stages (a, p, n) -- copies them in and samples their clusters -- for the next call;
        None re-stages the same batch.  Same arithmetic, same bits as the serial step on the same sequence of batches."""
        if not pipelined:
            static = [t.detach().clone() for t in (anchors, positives, negatives)]

            def step():
                xyz, feats, att, ep = self.get_train_model(static[0], static[1], static[2], True)
                loss, ep = self.get_loss(xyz, feats, att, ep)
                self.get_train_op(loss, lr=lr, end_points=ep, grad_hook=grad_hook, grad_scale=grad_scale)
                return loss.detach()

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                loss = step()

            def replay(anchors=None, positives=None, negatives=None):
                for dst, src in zip(static, (anchors, positives, negatives)):
                    if src is not None:
                        dst.copy_(src, non_blocking=True)
                graph.replay()
                self._adam["t"] += 1
                self.invalidate()  # the replay moved the weights and the BN shadows: the folded eval copy is stale
                return loss

            replay.graph = graph
            replay.static_inputs = static  # (anchors, positives, negatives) buffers the graph reads
            replay.loss = loss
            return replay

        # two input sets and two sets of sampled clusters, used alternately: step k runs on set k & 1 and samples set (k + 1) & 1
        inputs = [[t.detach().clone() for t in (anchors, positives, negatives)] for _ in range(2)]
        sampler = torch.cuda.Stream(priority=-1)  # high priority: its few CTAs take SMs as soon as blocks of the BN passes retire

        def sample_into(dst, src_inputs):
            kp, idx, cnt = self.sample_clusters(torch.cat(src_inputs, dim=0))
            if dst is None:
                return [kp, idx, cnt]
            for d, v in zip(dst, (kp, idx, cnt)):
                d.copy_(v)
            return dst

        def step(cur, samples, fork):
            hook = None
            if fork:
                main = torch.cuda.current_stream()

                def hook():
                    # the forward has launched its last per-point kernel (descriptor conv_mid): until the backward pass returns to that
                    # layer only per-cluster kernels run (tails, loss, their gradients: ~0.25 ms in which most SMs idle).  The sampling
                    # kernels of the next batch (whole-SM CTAs, 0.2 ms) go beside them on a high-priority stream.
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream())
                    sampler.wait_event(ev)
                    with torch.cuda.stream(sampler):
                        sample_into(samples[1 - cur], inputs[1 - cur])
            xyz, feats, att, ep = self.get_train_model(inputs[cur][0], inputs[cur][1], inputs[cur][2], True, presampled=samples[cur],
                                                       after_mid_hook=hook)
            loss, ep = self.get_loss(xyz, feats, att, ep)
            self.get_train_op(loss, lr=lr, end_points=ep, grad_hook=grad_hook, grad_scale=grad_scale)
            if fork:
                torch.cuda.current_stream().wait_stream(sampler)  # join: the next batch's clusters are part of this step
            return loss.detach()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            samples = [sample_into(None, inputs[0]), sample_into(None, inputs[1])]
            for _ in range(max(1, warmup)):
                step(0, samples, False)
        torch.cuda.current_stream().wait_stream(side)
        graphs, losses = [], []
        for cur in (0, 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                losses.append(step(cur, samples, True))
            graphs.append(g)
        state = {"cur": 0}
        loss = torch.zeros_like(losses[0])

        def replay(anchors=None, positives=None, negatives=None):
            cur = state["cur"]
            # stage the batch of the NEXT step: the graph about to run samples its clusters beside its own backward pass
            for dst, src, same in zip(inputs[1 - cur], (anchors, positives, negatives), inputs[cur]):
                dst.copy_(src if src is not None else same, non_blocking=True)
            graphs[cur].replay()
            loss.copy_(losses[cur])
            state["cur"] = 1 - cur
            self._adam["t"] += 1
            self.invalidate()
            return loss

        replay.graph = graphs
        replay.static_inputs = inputs[0]
        replay.next_inputs = lambda: inputs[1 - state["cur"]]  # where replay() stages the next batch
        replay.loss = loss
        replay.pipelined = True
        return replay

    def train_mode(self):
        """Mark the trainable variables as requiring grad (TF: GraphKeys.TRAINABLE_VARIABLES)."""
        for k, v in self.trainable_variables().items():
            v.requires_grad_(True)
        return self

"""Layer library -- the reference's models/layers.py:11-62,225-272 on torch CUDA tensors.

These are the UNFUSED, differentiable statements of the layers (one op per call, like the TF graph).  They exist
so that code written against `models.layers` keeps working and so that the training graph has an autograd path;
the inference hot path does not go through them: `models.feat3dnet` dispatches eval-mode forwards to the fused
CUDA kernels (csrc/mlp_fp32.cu, csrc/mlp_tc.cu) where conv+BN+ReLU+max-pool never leave the SM.

TF variable scopes become key prefixes of a flat parameter dict (`params`):
    <scope>/conv2d/weights (Cin,Cout)   <scope>/conv2d/biases (Cout)
    <scope>/bn/beta  <scope>/bn/gamma  <scope>/bn/moving_mean  <scope>/bn/moving_variance
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3   # layers.py:271
BN_DECAY = 0.9  # layers.py:251 (bn_decay=None everywhere in the model)


class PendingEma(object):
    """A moving-average update waiting for the optimiser step (layers.py:252-269: shadow <- shadow - (1 - decay)(shadow - batch)).
    The forward only records (batch statistic, decay); apply_ema_updates() performs all of them in three multi-tensor launches
    instead of four tiny kernels per statistic (72 per step for the 18 statistics of the model)."""
    __slots__ = ("stat", "decay")

    def __init__(self, stat, decay):
        self.stat, self.decay = stat, float(decay)

    def value(self, shadow):
        """the updated shadow as a new tensor (the op-by-op statement)"""
        return shadow - (1 - self.decay) * (shadow - self.stat)


def apply_ema_updates(weights, updates):
    """weights[k] <- weights[k] - (1 - decay)(weights[k] - batch statistic) for every k in `updates` ({name: PendingEma}), in
    place; the arithmetic (sub, mul by the Python scalar, sub) is the op-by-op statement's, so the results are bit-identical."""
    by_decay = {}
    for k, u in updates.items():
        by_decay.setdefault(u.decay, ([], []))
        by_decay[u.decay][0].append(weights[k])
        by_decay[u.decay][1].append(u.stat.to(weights[k].dtype))
    with torch.no_grad():
        for decay, (dst, src) in by_decay.items():
            diffs = torch._foreach_sub(dst, src)
            torch._foreach_mul_(diffs, 1 - decay)
            torch._foreach_sub_(dst, diffs)
FUSED_TRAINING = True  # training-mode conv+BN(+ReLU) through csrc/train_layers.cu (False: the op-by-op torch statement)
CHAIN_ACTIVATIONS = True  # the per-point MLP chains leave their intermediate activations unmaterialised (DeferredActivation)
TRAIN_PRECISION = "bf16x3"  # contractions of the training layers: "bf16x3" = tcgen05 tensor cores, "fp32" = FFMA kernels
_PRECISION_CODE = {"fp32": 0, "bf16x3": 2}

relu = torch.relu
softplus = F.softplus


def _native():
    import importlib
    return importlib.import_module(("3dfeatnet_b200." if __name__.split(".")[0] == "3dfeatnet_b200" else "") + "_lib")


class DeferredActivation(object):
    """The output of a training-mode conv2d whose activation is NOT materialised: y = act(z * scale + shift) with z the layer's pre-BN
    tensor (B,M,S,C), coef = [scale | shift] (2C,) and relu the activation flag.  The next conv2d of the chain takes it as its input
    and forms y inside its contractions (csrc/dz_source.cuh, XSource).  `z` is the autograd handle of y: the gradient that flows into
    it is dL/dy (the producing layer's backward recomputes y from z)."""
    __slots__ = ("z", "coef", "relu")

    def __init__(self, z, coef, relu):
        self.z, self.coef, self.relu = z, coef, bool(relu)

    @property
    def shape(self):
        return self.z.shape

    @property
    def is_cuda(self):
        return self.z.is_cuda

    def dim(self):
        return self.z.dim()

    def materialize(self):
        """the activation as an ordinary tensor (op-by-op statement; differentiable through z)"""
        c = self.z.shape[-1]
        y = self.z * self.coef[:c] + self.coef[c:]
        return torch.relu(y) if self.relu else y


def _conv_bn_forward(x, w, b, gamma, beta, use_relu, precision, gbias=None, gs=0, pool_s=0, x_coef=None, x_relu=False, want_y=True,
                     want_coef=False):
    """-> dict(x, w, gamma, beta, z, y, mean, var, pooled, inv, coef): y when want_y; pooled / inv (max over groups of pool_s rows and
    1 / ties) when pool_s > 0; coef = this layer's BN scale / shift when want_coef.  x_coef: x holds the previous layer's z (chain)."""
    _lib = _native()
    L = _lib.lib()
    x2, w2 = x.detach().contiguous().float(), w.detach().contiguous().float()
    b2, g2, be2 = b.detach().contiguous().float(), gamma.detach().contiguous().float(), beta.detach().contiguous().float()
    gb2 = gbias.detach().contiguous().float() if gbias is not None else None
    _lib.require_cuda(x2, w2, b2, g2, be2)
    rows, cin, cout = x2.shape[0], x2.shape[1], w2.shape[1]
    nbytes = L.f3d_conv_bn_train_workspace_bytes(rows, cin, cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
    z = torch.empty((rows, cout), dtype=torch.float32, device=x2.device)
    mean = torch.empty(cout, dtype=torch.float32, device=x2.device)
    var = torch.empty_like(mean)
    pooled = inv = y = coef = None
    if pool_s:
        if rows % pool_s:
            raise ValueError("pool_s must divide the number of rows")
        pooled = torch.empty((rows // pool_s, cout), dtype=torch.float32, device=x2.device)
        inv = torch.empty_like(pooled)
    if want_y:
        y = torch.empty_like(z)
    if want_coef:
        coef = torch.empty(2 * cout, dtype=torch.float32, device=x2.device)
    xc = x_coef.detach().contiguous().float() if x_coef is not None else None
    p = _lib.ptr
    _lib.check(L.f3d_conv_bn_train_forward_chain(rows, cin, cout, p(x2), p(xc) if xc is not None else None, int(bool(x_relu)), p(w2), p(b2),
                                                 p(gb2) if gb2 is not None else None, int(gs), p(g2), p(be2), int(use_relu), BN_EPS, p(z),
                                                 p(y) if y is not None else None, int(pool_s), p(pooled) if pooled is not None else None,
                                                 p(inv) if inv is not None else None, p(coef) if coef is not None else None, p(mean), p(var),
                                                 precision, p(ws), nbytes, _lib.stream()), "conv_bn_train_forward_chain")
    return dict(x=x2, w=w2, gamma=g2, beta=be2, z=z, y=y, mean=mean, var=var, pooled=pooled, inv=inv, coef=coef, x_coef=xc)


def _conv_bn_backward(saved, use_relu, precision, gy, need_dx, pool_s=0, gpool=None, gs=0, x_relu=False):
    """saved: the dict of _conv_bn_forward.  gy: dense gradient or None; gpool: gradient of the pooled tensor or None.
    -> (dx, dW, db, dgamma, dbeta, dgroup_bias); dgroup_bias only when gs > 0."""
    _lib = _native()
    L = _lib.lib()
    x2, w2, g2, be2, z, mean, var = (saved[k] for k in ("x", "w", "gamma", "beta", "z", "mean", "var"))
    pooled, inv, xc = saved["pooled"], saved["inv"], saved["x_coef"]
    rows, cin, cout = x2.shape[0], x2.shape[1], w2.shape[1]
    gy = gy.contiguous().float() if gy is not None else None
    gpool = gpool.contiguous().float() if gpool is not None else None
    if gpool is None:
        pool_s = 0
    nbytes = L.f3d_conv_bn_train_workspace_bytes(rows, cin, cout)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
    dx = torch.empty_like(x2) if need_dx else None
    dw = torch.empty_like(w2)
    db, dg, dbe = (torch.empty(cout, dtype=torch.float32, device=x2.device) for _ in range(3))
    dgb = torch.empty((rows // gs, cout), dtype=torch.float32, device=x2.device) if gs > 0 else None
    p = _lib.ptr
    opt = lambda t: p(t) if t is not None else None
    _lib.check(L.f3d_conv_bn_train_backward_chain(rows, cin, cout, p(x2), opt(xc), int(bool(x_relu)), p(w2), p(g2), p(be2), p(z), p(mean), p(var),
                                                  int(use_relu), BN_EPS, opt(gy), int(pool_s), opt(pooled) if pool_s else None, opt(gpool),
                                                  opt(inv) if pool_s else None, opt(dx), p(dw), p(db), p(dg), p(dbe), opt(dgb), int(gs), precision,
                                                  p(ws), nbytes, _lib.stream()), "conv_bn_train_backward_chain")
    return dx, dw, db, dg, dbe, dgb


def _max_pool_forward(x3):
    """x3 (groups, s, c) contiguous fp32 -> pooled (groups, c), 1/ties (groups, c)."""
    _lib = _native()
    groups, s, c = x3.shape
    out = torch.empty((groups, c), dtype=torch.float32, device=x3.device)
    inv = torch.empty_like(out)
    _lib.check(_lib.lib().f3d_maxpool_samples_forward(groups, s, c, _lib.ptr(x3), _lib.ptr(out), _lib.ptr(inv), _lib.stream()),
               "maxpool_samples_forward")
    return out, inv


class _ConvBnTrain(torch.autograd.Function):
    """conv 1x1 + bias + batch-norm with BATCH statistics + optional ReLU as one differentiable CUDA op
    (csrc/train_layers.cu, contractions in csrc/train_tc.cu).  Returns (out, pooled, coef, batch_mean, batch_var); coef and the
    moments are not differentiable (the moments only feed the EMA shadows).  Saves x and z (pre-BN); the backward recomputes y.

    mode "y":      out = the activation (rows, cout).
    mode "pool":   the layer is followed by tf.reduce_max over groups of pool_s consecutive rows (the sample axis) and ONLY feeds that
                   pool (detector conv2, descriptor conv_mid): out = pooled (rows/pool_s, cout); in the backward the dense (rows, cout)
                   gradient is never materialised: it is rebuilt on the fly from the pooled maxima, the pooled gradient and the tie counts.
    mode "defer":  the activation is not materialised: out = z, standing for y in the autograd graph (see DeferredActivation), and
                   coef = the BN scale / shift the consumer needs.  With pool_s > 0 `pooled` = max of the activation over the groups is
                   returned as well (descriptor conv1 feeds the pool and conv_mid); the backward then sums both gradients on the fly.
    x_coef:        x is itself a deferred activation (the previous layer's z) with this scale / shift and x_relu.
    gbias (rows/gs, cout) or None: a per-group additive term of the pre-BN activation (the pooled half of a
    concat([x, tile(pooled)]) input, see conv2d(concat_pooled=)); its gradient is the per-group row sum of dz."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, use_relu, pool_s, gbias, gs, mode, x_coef, x_relu):
        ctx.precision = _PRECISION_CODE[TRAIN_PRECISION]
        ctx.set_materialize_grads(False)  # no zero tensors for the outputs nobody differentiates (moments, coef, unused pooled)
        sv = _conv_bn_forward(x, w, b, gamma, beta, use_relu, ctx.precision, gbias, gs if gbias is not None else 0, int(pool_s),
                              x_coef=x_coef, x_relu=x_relu, want_y=(mode == "y"), want_coef=(mode == "defer"))
        ctx.use_relu, ctx.pool_s, ctx.gs, ctx.mode, ctx.x_relu = bool(use_relu), int(pool_s), int(gs) if gbias is not None else 0, mode, bool(x_relu)
        ctx.keys = [k for k in ("x", "w", "gamma", "beta", "z", "mean", "var", "pooled", "inv", "x_coef") if sv[k] is not None]
        ctx.save_for_backward(*[sv[k] for k in ctx.keys])
        mean, var, coef, pooled = sv["mean"], sv["var"], sv["coef"], sv["pooled"]
        ctx.mark_non_differentiable(*[t for t in (mean, var, coef) if t is not None])
        if mode == "y":
            return sv["y"], None, None, mean, var
        if mode == "pool":
            return pooled, None, None, mean, var
        return sv["z"], pooled, coef, mean, var

    @staticmethod
    def backward(ctx, gout, gpooled, _gc, _gm, _gv):
        saved = dict.fromkeys(("x", "w", "gamma", "beta", "z", "mean", "var", "pooled", "inv", "x_coef"))
        saved.update(zip(ctx.keys, ctx.saved_tensors))
        if ctx.mode == "pool":
            gy, gpool = None, (gout if gout is not None else torch.zeros_like(saved["pooled"]))
        else:
            gy, gpool = gout, (gpooled if ctx.pool_s else None)
            if gy is None and gpool is None:
                gy = torch.zeros_like(saved["z"])
        dx, dw, db, dg, dbe, dgb = _conv_bn_backward(saved, ctx.use_relu, ctx.precision, gy, ctx.needs_input_grad[0], ctx.pool_s, gpool, ctx.gs,
                                                     ctx.x_relu)
        return dx, dw, db, dg, dbe, None, None, dgb, None, None, None, None


class _LinearTC(torch.autograd.Function):
    """out = x @ W for row-major fp32 (rows, k) x (k, nout) on the tensor-core contractions of the training layers (csrc/train_tc.cu through
    f3d_linear_forward / _backward): the per-cluster term of conv_mid, which used to be the one cuBLAS call of the training step."""

    @staticmethod
    def forward(ctx, x, w):
        _lib = _native()
        L = _lib.lib()
        x2, w2 = x.detach().contiguous().float(), w.detach().contiguous().float()
        _lib.require_cuda(x2, w2)
        rows, k, nout = x2.shape[0], x2.shape[1], w2.shape[1]
        nbytes = L.f3d_linear_workspace_bytes(rows, k, nout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
        out = torch.empty((rows, nout), dtype=torch.float32, device=x2.device)
        _lib.check(L.f3d_linear_forward(rows, k, nout, _lib.ptr(x2), _lib.ptr(w2), _lib.ptr(out), _lib.ptr(ws), nbytes, _lib.stream()), "linear_forward")
        ctx.save_for_backward(x2, w2)
        return out

    @staticmethod
    def backward(ctx, g):
        _lib = _native()
        L = _lib.lib()
        x2, w2 = ctx.saved_tensors
        rows, k, nout = x2.shape[0], x2.shape[1], w2.shape[1]
        g2 = g.contiguous().float()
        nbytes = L.f3d_linear_workspace_bytes(rows, k, nout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w2) if ctx.needs_input_grad[1] else None
        _lib.check(L.f3d_linear_backward(rows, k, nout, _lib.ptr(x2), _lib.ptr(w2), _lib.ptr(g2), _lib.ptr(dx) if dx is not None else None,
                                         _lib.ptr(dw) if dw is not None else None, _lib.ptr(ws), nbytes, _lib.stream()), "linear_backward")
        return dx, dw


def linear_rows(x, w):
    """(rows, k) @ (k, nout): the tensor-core contraction where its shapes are covered (training on the GPU with TRAIN_PRECISION "bf16x3"),
    torch.matmul otherwise."""
    k, nout = w.shape
    if (FUSED_TRAINING and TRAIN_PRECISION == "bf16x3" and x.is_cuda and x.dim() == 2 and k % 8 == 0 and k <= 128 and nout % 16 == 0 and nout <= 256):
        return _LinearTC.apply(x, w)
    return torch.matmul(x, w)


class _MaxPoolSamples(torch.autograd.Function):
    """tf.reduce_max over the sample axis of a (B,M,S,C) tensor with TensorFlow's tie-sharing gradient (csrc/train_layers.cu)."""

    @staticmethod
    def forward(ctx, x):
        _lib = _native()
        x4 = x.detach().contiguous().float()
        _lib.require_cuda(x4)
        b, m, s, c = x4.shape
        out, inv = _max_pool_forward(x4.view(b * m, s, c))
        ctx.save_for_backward(x4, out, inv)
        return out.view(b, m, 1, c)

    @staticmethod
    def backward(ctx, g):
        _lib = _native()
        x4, out, inv = ctx.saved_tensors
        b, m, s, c = x4.shape
        g = g.contiguous().float()
        dx = torch.empty_like(x4)
        _lib.check(_lib.lib().f3d_maxpool_samples_backward(b * m, s, c, _lib.ptr(x4), _lib.ptr(out), _lib.ptr(inv), _lib.ptr(g), _lib.ptr(dx),
                                                           _lib.stream()), "maxpool_samples_backward")
        return dx


class _DetectorHeads(torch.autograd.Function):
    """attention = softplus(h w_att + b_att), orientation = atan2(l2_normalize(h w_ori + b_ori)) over the rows of h (R,k):
    the detector's heads (feat3dnet.py:142-149) as one forward and one backward CUDA op (csrc/train.cu)."""

    @staticmethod
    def forward(ctx, h, w_att, b_att, w_ori, b_ori):
        _lib = _native()
        t = [v.detach().contiguous().float() for v in (h, w_att, b_att, w_ori, b_ori)]
        _lib.require_cuda(*t)
        rows, k = t[0].shape
        att = torch.empty(rows, dtype=torch.float32, device=t[0].device)
        ori = torch.empty_like(att)
        _lib.check(_lib.lib().f3d_detector_heads_forward(rows, k, *(_lib.ptr(v) for v in t), _lib.ptr(att), _lib.ptr(ori), _lib.stream()),
                   "detector_heads_forward")
        ctx.save_for_backward(*t)
        return att, ori

    @staticmethod
    def backward(ctx, g_att, g_ori):
        _lib = _native()
        L = _lib.lib()
        h, w_att, b_att, w_ori, b_ori = ctx.saved_tensors
        rows, k = h.shape
        ga = g_att.contiguous().float() if g_att is not None else None
        go = g_ori.contiguous().float() if g_ori is not None else None
        dh = torch.empty_like(h) if ctx.needs_input_grad[0] else None
        dwa, dba, dwo, dbo = torch.empty_like(w_att), torch.empty_like(b_att), torch.empty_like(w_ori), torch.empty_like(b_ori)
        nbytes = L.f3d_detector_heads_workspace_bytes(k)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=h.device)
        _lib.check(L.f3d_detector_heads_backward(rows, k, _lib.ptr(h), _lib.ptr(w_att), _lib.ptr(b_att), _lib.ptr(w_ori), _lib.ptr(b_ori),
                                                 _lib.ptr(ga), _lib.ptr(go), _lib.ptr(dh), _lib.ptr(dwa), _lib.ptr(dba), _lib.ptr(dwo),
                                                 _lib.ptr(dbo), _lib.ptr(ws), nbytes, _lib.stream()), "detector_heads_backward")
        return dh, dwa, dba, dwo, dbo


def detector_heads(new_points, params, scope):
    """(B,M,1,k) pooled detector features -> (attention (B,M), orientation (B,M)) through the fused heads op, or None when the
    shape / device is not covered (the caller then states the heads with conv2d)."""
    if not (FUSED_TRAINING and new_points.is_cuda and new_points.dim() == 4 and new_points.shape[2] == 1
            and new_points.shape[3] % 32 == 0 and new_points.shape[3] <= 128):
        return None
    b, m, _, k = new_points.shape
    wa, ba = params[scope + "/attention/conv2d/weights"], params[scope + "/attention/conv2d/biases"]
    wo, bo = params[scope + "/orientation/conv2d/weights"], params[scope + "/orientation/conv2d/biases"]
    att, ori = _DetectorHeads.apply(new_points.reshape(b * m, k), wa.reshape(k), ba.reshape(1), wo.reshape(k, 2), bo.reshape(2))
    return att.view(b, m), ori.view(b, m)


def max_pool_samples(x):
    """(B,M,S,C) -> (B,M,1,C): tf.reduce_max(x, axis=[2], keep_dims=True).  CUDA op when x is on the GPU and C % 4 == 0."""
    if x.is_cuda and x.dim() == 4 and x.shape[3] % 4 == 0:
        return _MaxPoolSamples.apply(x)
    return x.amax(dim=2, keepdim=True)  # amax shares the gradient among ties like TF


def conv_bn_train(x, w, b, gamma, beta, use_relu=True, pool_s=0, gbias=None, gs=0):
    """(rows,cin) x (cin,cout) -> (out, batch_mean, batch_var): the training-mode layer as one CUDA op (see _ConvBnTrain)."""
    out, _, _, mean, var = _ConvBnTrain.apply(x, w, b, gamma, beta, use_relu, pool_s, gbias, gs, "pool" if pool_s else "y", None, False)
    return out, mean, var


def conv_bn_train_chain(x, w, b, gamma, beta, use_relu=True, pool_s=0, gbias=None, gs=0, mode="y", x_coef=None, x_relu=False):
    """The layer inside a chain of unmaterialised activations -> (out, pooled, coef, batch_mean, batch_var); see _ConvBnTrain."""
    return _ConvBnTrain.apply(x, w, b, gamma, beta, use_relu, pool_s, gbias, gs, mode, x_coef, x_relu)


def batch_norm_template(inputs, is_training, scope, moments_dims, bn_decay, params, new_stats=None):
    """layers.py:225-272.  Training: batch moments over `moments_dims` (population variance) and EMA update of the
    shadows (returned through `new_stats`, applied by the caller after the step); eval: the EMA shadows."""
    gamma, beta = params[scope + "/gamma"], params[scope + "/beta"]
    if is_training:
        mean = inputs.mean(dim=moments_dims)
        var = inputs.var(dim=moments_dims, unbiased=False)
        if new_stats is not None:
            decay = bn_decay if bn_decay is not None else BN_DECAY
            mm, mv = params[scope + "/moving_mean"], params[scope + "/moving_variance"]
            new_stats[scope + "/moving_mean"] = PendingEma(mean.detach(), decay)
            new_stats[scope + "/moving_variance"] = PendingEma(var.detach(), decay)
    else:
        mean, var = params[scope + "/moving_mean"], params[scope + "/moving_variance"]
    inv = torch.rsqrt(var + BN_EPS) * gamma  # tf.nn.batch_normalization
    return inputs * inv + (beta - mean * inv)


def batch_norm_for_fc(inputs, is_training, bn_decay, scope, params, new_stats=None):
    return batch_norm_template(inputs, is_training, scope, [0], bn_decay, params, new_stats)


def batch_norm_for_conv1d(inputs, is_training, bn_decay, scope, params, new_stats=None):
    return batch_norm_template(inputs, is_training, scope, [0, 1], bn_decay, params, new_stats)


def batch_norm_for_conv2d(inputs, is_training, bn_decay, scope, params, new_stats=None):
    return batch_norm_template(inputs, is_training, scope, [0, 1, 2], bn_decay, params, new_stats)


def batch_norm_for_conv3d(inputs, is_training, bn_decay, scope, params, new_stats=None):
    return batch_norm_template(inputs, is_training, scope, [0, 1, 2, 3], bn_decay, params, new_stats)


def dropout(inputs, is_training, scope=None, keep_prob=0.5, noise_shape=None):
    """layers.py:107-127 (not called by the model): tf.nn.dropout when training -- kept elements scaled by 1/keep_prob,
    the mask drawn with `noise_shape` and broadcast -- identity otherwise."""
    if not is_training:
        return inputs
    if not 0.0 < keep_prob <= 1.0:
        raise ValueError("dropout: keep_prob must be in (0, 1], got %r" % (keep_prob,))
    shape = list(inputs.shape) if noise_shape is None else list(noise_shape)
    keep = torch.rand(shape, device=inputs.device, dtype=inputs.dtype) < keep_prob
    return inputs * keep.to(inputs.dtype) / keep_prob


def fully_connected(inputs, num_outputs, scope, use_xavier=True, stddev=1e-3, weight_decay=0.0, activation_fn=relu, bn=False,
                    bn_decay=None, is_training=None, params=None, new_stats=None):
    """layers.py:130-171 (not called by the model): (B,Cin) @ <scope>/weights (Cin,num_outputs) + <scope>/biases, then
    optional BN over the batch axis (<scope>/bn/...) and the activation.  The initialiser arguments are accepted and unused:
    variables come from `params`."""
    if params is None or scope is None:
        raise ValueError("fully_connected needs params= (flat dict keyed by TF scope names) and scope=")
    if inputs.dim() != 2:
        raise ValueError("fully_connected: inputs must be (B, Cin), got %s" % (tuple(inputs.shape),))
    w, b = params[scope + "/weights"], params[scope + "/biases"]
    if tuple(w.shape) != (inputs.shape[1], num_outputs):
        raise ValueError("fully_connected: weight shape %s does not match (%d -> %d)" % (tuple(w.shape), inputs.shape[1], num_outputs))
    outputs = torch.matmul(inputs, w) + b
    if bn:
        outputs = batch_norm_for_fc(outputs, bool(is_training), bn_decay, scope + "/bn", params, new_stats)
    if activation_fn is not None:
        outputs = activation_fn(outputs)
    return outputs


def chain_supported():
    """True when the training layers may leave their activations unmaterialised (DeferredActivation): the tensor-core contractions form
    them on the fly; the fp32 FFMA path (TRAIN_PRECISION = "fp32") and the op-by-op statement read a stored tensor."""
    return FUSED_TRAINING and CHAIN_ACTIVATIONS and TRAIN_PRECISION == "bf16x3"


def conv2d(inputs, num_outputs, kernel_size, stride=[1, 1], padding='SAME', activation=relu, bn=True, bn_decay=None,
           is_training=None, scope=None, reuse=None, params=None, new_stats=None, pool_samples=False, concat_pooled=None,
           defer=False, also_pool=False):
    """ 2D convolution with non-linear operation (layers.py:11-46): slim.conv2d WITH bias -> BN -> activation.

    inputs: (B,H,W,C).  Only the 1x1 / stride-1 kernels the model uses are implemented.
    pool_samples=True additionally applies tf.reduce_max(axis=[2], keep_dims=True) to the result -- the callers that pool
    right after the layer say so here, which lets the training path fuse the pool's gradient into the layer's backward.
    concat_pooled (B,H,1,C2) or None: the layer's input is concat([inputs, tile(concat_pooled, W)], -1) (the reference's
    feat3dnet.py:60-66); weights are (C+C2, Cout).  The training path never materialises the concatenation.
    Chains (training mode on the GPU, chain_supported()): `inputs` may be a DeferredActivation (the previous layer's output, not
    materialised); defer=True returns one instead of a tensor -- for the next conv2d of the chain only; also_pool=True (with defer)
    returns (DeferredActivation, tf.reduce_max(activation, axis=[2], keep_dims=True)) for a layer that feeds the pool and a conv2d."""
    if list(kernel_size) != [1, 1] or list(stride) != [1, 1]:
        raise ValueError("conv2d: only kernel_size=[1,1], stride=[1,1] is supported (all the model uses)")
    if params is None or scope is None:
        raise ValueError("conv2d needs params= (flat dict keyed by TF scope names) and scope=")
    w, b = params[scope + "/conv2d/weights"], params[scope + "/conv2d/biases"]
    cin_total = inputs.shape[-1] + (concat_pooled.shape[-1] if concat_pooled is not None else 0)
    if w.shape[-1] != num_outputs or w.shape[-2] != cin_total:
        raise ValueError("conv2d: weight shape %s does not match (%d -> %d)" % (tuple(w.shape), cin_total, num_outputs))
    fused = (bn and is_training and inputs.is_cuda and FUSED_TRAINING and (activation is relu or activation is None)
             and num_outputs % 16 == 0 and num_outputs & (num_outputs - 1) == 0)
    deferred_in = isinstance(inputs, DeferredActivation)
    if (deferred_in or defer) and not (fused and chain_supported() and inputs.dim() == 4):
        raise ValueError("conv2d: deferred activations need the fused tensor-core training path (see chain_supported())")
    if also_pool and not defer:
        raise ValueError("conv2d: also_pool comes with defer=True (use pool_samples=True for a layer that only feeds the pool)")
    if fused:
        # training mode on the GPU: conv + bias + batch-statistics BN + ReLU, forward and backward, in csrc/train_layers.cu
        fuse_pool = pool_samples and inputs.dim() == 4
        w2 = w.reshape(w.shape[-2], w.shape[-1])
        gbias, gs = None, 0
        x_rows = inputs.z if deferred_in else inputs
        if concat_pooled is not None:
            # input = concat([inputs, tile(concat_pooled)], -1) without building it: the pooled half contributes
            # concat_pooled @ W_bottom once per cluster (the split-weight identity)
            c1 = inputs.shape[-1]
            gbias = linear_rows(concat_pooled.reshape(-1, concat_pooled.shape[-1]), w2[c1:])
            gs, w2 = inputs.shape[2], w2[:c1]
        mode = "defer" if defer else "pool" if fuse_pool else "y"
        pool_s = inputs.shape[2] if (fuse_pool or also_pool) else 0
        y, pooled, coef, mean, var = conv_bn_train_chain(
            x_rows.reshape(-1, x_rows.shape[-1]), w2, b, params[scope + "/bn/gamma"], params[scope + "/bn/beta"], activation is relu, pool_s,
            gbias, gs, mode, inputs.coef if deferred_in else None, inputs.relu if deferred_in else False)
        if new_stats is not None:
            decay = bn_decay if bn_decay is not None else BN_DECAY
            mm, mv = params[scope + "/bn/moving_mean"], params[scope + "/bn/moving_variance"]
            new_stats[scope + "/bn/moving_mean"] = PendingEma(mean.detach(), decay)
            new_stats[scope + "/bn/moving_variance"] = PendingEma(var.detach(), decay)
        if defer:
            out = DeferredActivation(y.reshape(*inputs.shape[:-1], num_outputs), coef, activation is relu)
            return (out, pooled.reshape(inputs.shape[0], inputs.shape[1], 1, num_outputs)) if also_pool else out
        if fuse_pool:
            return y.reshape(inputs.shape[0], inputs.shape[1], 1, num_outputs)
        y = y.reshape(*inputs.shape[:-1], num_outputs)
        return max_pool_samples(y) if pool_samples else y
    if concat_pooled is not None:
        inputs = torch.cat((inputs, concat_pooled.expand(-1, -1, inputs.shape[2], -1)), dim=3)
    net = torch.matmul(inputs, w.reshape(w.shape[-2], w.shape[-1])) + b
    if bn:
        net = batch_norm_for_conv2d(net, bool(is_training), bn_decay, scope + "/bn", params, new_stats)
    if activation is not None:
        net = activation(net)
    return max_pool_samples(net) if pool_samples else net


def pairwise_dist(A, B):
    ''' Computes pairwise distance (layers.py:49-62)

    :param A: (B x N x D) containing descriptors of A
    :param B: (B x N x D) containing descriptors of B
    :return: (B x N x N) tensor. Element[i,j,k] denotes the distance between the jth descriptor in ith model of A,
             and kth descriptor in ith model of B
    '''
    return ((A.unsqueeze(2) - B.unsqueeze(1)) ** 2).sum(3)


class _TripletLoss(torch.autograd.Function):
    """Fused CUDA loss of Feat3dNet.get_loss (csrc/train.cu): the row minima of the two pairwise-distance matrices, the
    attention-normalised hinge and -- in the same call -- the gradients w.r.t. the three descriptor sets and the
    attention, so that no (B,M,M[,F]) tensor is ever materialised."""

    @staticmethod
    def forward(ctx, anchors, positives, negatives, attention, margin):
        import importlib
        root = __name__.split(".")[0]
        _lib = importlib.import_module(("3dfeatnet_b200." if root == "3dfeatnet_b200" else "") + "_lib")
        fa, fp, fn = (t.detach().contiguous().float() for t in (anchors, positives, negatives))
        att = attention.detach().contiguous().float() if attention is not None else None
        _lib.require_cuda(fa, fp, fn)
        b, m, f = fa.shape
        L = _lib.lib()
        nbytes = L.f3d_triplet_loss_workspace_bytes(b, m)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=fa.device)
        loss = torch.empty(1, dtype=torch.float32, device=fa.device)
        dfa, dfp, dfn = torch.empty_like(fa), torch.empty_like(fp), torch.empty_like(fn)
        datt = torch.empty_like(att) if att is not None else None
        _lib.check(L.f3d_triplet_loss(b, m, f, float(margin), _lib.ptr(fa), _lib.ptr(fp), _lib.ptr(fn),
                                      _lib.ptr(att) if att is not None else None, _lib.ptr(loss), _lib.ptr(dfa), _lib.ptr(dfp),
                                      _lib.ptr(dfn), _lib.ptr(datt) if datt is not None else None, _lib.ptr(ws), nbytes,
                                      _lib.stream()), "triplet_loss")
        ctx.save_for_backward(dfa, dfp, dfn, datt if datt is not None else torch.empty(0, device=fa.device))
        ctx.has_att = datt is not None
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        dfa, dfp, dfn, datt = ctx.saved_tensors
        return g * dfa, g * dfp, g * dfn, (g * datt if ctx.has_att else None), None


def triplet_loss(anchors, positives, negatives, attention=None, margin=0.2):
    """mean_b max(0, sum_i w_bi (min_k |a_bi-p_bk|^2 - min_k |a_bi-n_bk|^2) + margin), w = attention / sum(attention)
    (uniform when attention is None): models/feat3dnet.py:315-357 as one differentiable CUDA op."""
    return _TripletLoss.apply(anchors, positives, negatives, attention, margin)

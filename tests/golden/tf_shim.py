"""An eager stand-in for the slice of TensorFlow 1.x that the reference's model files touch, so that
models/layers.py, models/pointnet_common.py and models/feat3dnet.py can be IMPORTED AND EXECUTED UNMODIFIED from
/root/reference without TensorFlow (build container only; test infrastructure, never shipped).

What is the reference's: every line of graph construction -- which layer follows which, scopes and variable names, the
activation / BN flags, pooling axes, tile + concat, the rotation matrix and its transpose, squeeze / l2_normalize / atan2,
the loss.  What is restated here: the primitive ops those lines call (tf.expand_dims, tf.tile, tf.reduce_max,
tf.nn.moments, tf.nn.batch_normalization, tf.train.ExponentialMovingAverage, slim.conv2d with a 1x1 kernel, ...), each by
its documented TF 1.15 semantics, on float64 torch tensors.  Variables are not created but looked up by their full
TF name ("detection/conv0/conv2d/weights", ".../bn/gamma", ...) in `STORE.params`; the two EMA shadows of a BN layer are
".../bn/moving_mean" and ".../bn/moving_variance" (SURVEY.md appendix B).  The custom ops (tf_ops.*) are bound to the C
oracle, which is pinned to the reference's CUDA kernels (tests/golden/ref_gpu_ops.npz).

    install(reference_root)  ->  puts `tensorflow`, `tensorflow.contrib.slim`, `tf_ops.*` in sys.modules and the reference
                                 root on sys.path; afterwards `import models.feat3dnet` is the reference's own file.
"""
import contextlib
import sys
import types

import numpy as np
import torch

DTYPE = torch.float64


class _Dim(object):
    def __init__(self, v):
        self.value = int(v)


class T(torch.Tensor):
    """torch tensor with the two TF-1 Tensor methods the reference calls."""

    def get_shape(self):
        return [_Dim(s) for s in self.shape]


def t(x, dtype=DTYPE):
    if isinstance(x, (list, tuple)):
        return torch.stack([t(e, dtype) for e in x], dim=0).as_subclass(T)
    return torch.as_tensor(x).to(dtype).as_subclass(T)


class Store(object):
    """Variables by TF name, the scope stack, EMA updates recorded during a training-mode pass."""

    def __init__(self):
        self.params, self.updates, self.scope, self.touched = {}, {}, [], set()

    def name(self, leaf):
        return "/".join(self.scope + [leaf])

    def get(self, leaf):
        n = self.name(leaf)
        self.touched.add(n)
        return t(self.params[n])


STORE = Store()


# ------------------------------------------------------------------------------------------------ tf.*
@contextlib.contextmanager
def variable_scope(scope, reuse=None):
    pushed = isinstance(scope, str)
    if pushed:
        STORE.scope.append(scope)
    try:
        yield "/".join(STORE.scope)
    finally:
        if pushed:
            STORE.scope.pop()


def get_variable_scope():
    return None  # only ever passed back to variable_scope(..., reuse=False): re-enters the current scope


@contextlib.contextmanager
def control_dependencies(ops):
    yield


def _axes(axis):
    return list(axis) if isinstance(axis, (list, tuple)) else [axis]


def reduce_max(x, axis=None, keep_dims=False):
    return torch.amax(x, dim=_axes(axis), keepdim=keep_dims)


def reduce_min(x, axis=None, keep_dims=False):
    return torch.amin(x, dim=_axes(axis), keepdim=keep_dims)


def reduce_sum(x, axis=None, keep_dims=False):
    return x.sum() if axis is None else torch.sum(x, dim=_axes(axis), keepdim=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False):
    return x.mean() if axis is None else torch.mean(x, dim=_axes(axis), keepdim=keep_dims)


def squeeze(x, axis=None):
    for a in sorted(_axes(axis), reverse=True):
        assert x.shape[a] == 1
        x = x.squeeze(a)
    return x


def cond(pred, true_fn, false_fn):
    assert isinstance(pred, bool), "the shim runs eagerly: is_training is a Python bool"
    return true_fn() if pred else false_fn()


def Variable(initial_value, name=None, trainable=True):
    return STORE.get(name)


def constant(value, shape=None):
    return t(np.full(shape, value) if shape is not None else value)


class ExponentialMovingAverage(object):
    """tf.train.ExponentialMovingAverage on the two tensors tf.nn.moments returned for the current BN scope: apply() is
    shadow <- shadow - (1 - decay) (shadow - value) (no zero-debias for plain tensors, num_updates=None); average() reads the
    shadow.  The shadows live in the store as <scope>/moving_mean and <scope>/moving_variance."""

    def __init__(self, decay):
        self.decay = decay

    def apply(self, tensors):
        for v in tensors:
            shadow = t(STORE.params[v._ema_name])
            STORE.touched.add(v._ema_name)
            STORE.updates[v._ema_name] = (shadow - (1 - self.decay) * (shadow - v)).detach().numpy()
        return None

    def average(self, v):
        STORE.touched.add(v._ema_name)
        return t(STORE.params[v._ema_name])


def moments(x, axes, name=None):
    """tf.nn.moments: mean and POPULATION variance over `axes`."""
    mean = torch.mean(x, dim=list(axes))
    var = torch.mean((x - mean) ** 2, dim=list(axes))
    mean._ema_name, var._ema_name = STORE.name("moving_mean"), STORE.name("moving_variance")
    return mean, var


def batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
    """tf.nn.batch_normalization: inv = rsqrt(var + eps) * scale; x * inv + (offset - mean * inv)."""
    inv = torch.rsqrt(variance + variance_epsilon) * scale
    return x * inv + (offset - mean * inv)


def l2_normalize(x, dim=None, epsilon=1e-12, axis=None):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), epsilon))."""
    d = dim if dim is not None else axis
    return x * torch.rsqrt(torch.clamp(torch.sum(x * x, dim=_axes(d), keepdim=True), min=epsilon))


def conv2d_slim(inputs, num_outputs, kernel_size, stride=1, padding='SAME', activation_fn=torch.relu, weights_initializer=None,
                reuse=None, scope=None):
    """slim.conv2d with a 1x1 kernel and stride 1 on NHWC: inputs @ weights[0,0] + biases (slim adds a zero-initialised bias
    unless a normalizer_fn is given), then activation_fn."""
    assert list(kernel_size) == [1, 1] and list(_axes(stride)) in ([1, 1], [1]), "the model only uses 1x1 / stride 1"
    with variable_scope(scope):
        w, b = STORE.get("weights"), STORE.get("biases")
    w = w.reshape(w.shape[-2], w.shape[-1])  # TF stores (1,1,Cin,Cout)
    assert w.shape == (inputs.shape[-1], num_outputs), (STORE.scope, tuple(w.shape), inputs.shape[-1], num_outputs)
    out = torch.matmul(inputs, w) + b
    return activation_fn(out) if activation_fn is not None else out


def _unsupported(name):
    def fn(*a, **k):
        raise NotImplementedError("tf_shim: %s is outside the slice the model's forward pass and loss use" % name)
    return fn


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def build_tensorflow():
    tf = _module(
        "tensorflow", float32=DTYPE, float16=torch.float16, Tensor=T, AUTO_REUSE="AUTO_REUSE",
        variable_scope=variable_scope, get_variable_scope=get_variable_scope, control_dependencies=control_dependencies,
        expand_dims=lambda x, axis=None, dim=None: torch.unsqueeze(x, axis if axis is not None else dim),
        tile=lambda x, multiples: (x if isinstance(x, torch.Tensor) else t(x, torch.int64)).repeat(*[int(m) for m in multiples]),
        shape=lambda x: tuple(x.shape),
        concat=lambda values, axis: torch.cat(list(values), dim=axis),
        stack=lambda values, axis=0: torch.movedim(t(list(values)), 0, axis),
        split=lambda x, n, axis=0: list(torch.chunk(x, n, dim=axis)),
        transpose=lambda x, perm: x.permute(*perm),
        reshape=lambda x, shape: x.reshape(*shape),
        matmul=torch.matmul, cos=torch.cos, sin=torch.sin, atan2=torch.atan2, identity=lambda x: x,
        ones_like=torch.ones_like, zeros_like=torch.zeros_like, maximum=lambda a, b: torch.maximum(t(a), t(b)),
        multiply=torch.mul, squared_difference=lambda a, b: (a - b) ** 2,
        reduce_max=reduce_max, reduce_min=reduce_min, reduce_sum=reduce_sum, reduce_mean=reduce_mean, squeeze=squeeze,
        cond=cond, no_op=lambda: None, Variable=Variable, constant=constant,
        zeros=lambda shape, dtype=None: t(np.zeros(shape)),
        placeholder=_unsupported("tf.placeholder"), gradients=_unsupported("tf.gradients"),
        get_collection=_unsupported("tf.get_collection"), get_variable=_unsupported("tf.get_variable"),
    )
    tf.nn = _module("tensorflow.nn", relu=torch.relu, softplus=torch.nn.functional.softplus, moments=moments,
                    batch_normalization=batch_normalization, l2_normalize=l2_normalize)
    tf.train = _module("tensorflow.train", ExponentialMovingAverage=ExponentialMovingAverage,
                       AdamOptimizer=_unsupported("tf.train.AdamOptimizer"))
    tf.summary = _module("tensorflow.summary", histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)
    slim = _module("tensorflow.contrib.slim", conv2d=conv2d_slim, variance_scaling_initializer=lambda *a, **k: None)
    tf.contrib = _module("tensorflow.contrib", slim=slim)
    return tf, slim


def build_tf_ops():
    """tf_ops.grouping.tf_grouping / tf_ops.sampling.tf_sampling bound to the C oracle (fp32 in, fp32 / int32 out)."""
    from oracle import ops

    def f32(x):
        a = x.detach().numpy().astype(np.float32)
        assert np.array_equal(a.astype(np.float64), x.detach().numpy()), "coordinates must be float32-representable"
        return a

    def query_ball_point(radius, nsample, xyz1, xyz2):
        idx, cnt = ops.query_ball_point(radius, nsample, f32(xyz1), f32(xyz2))
        return t(idx, torch.int32), t(cnt, torch.int32)

    def knn_point(k, xyz1, xyz2):
        val, idx = ops.knn_point(k, f32(xyz1), f32(xyz2))
        return t(val), t(idx, torch.int32)

    grouping = _module("tf_ops.grouping.tf_grouping", query_ball_point=query_ball_point, knn_point=knn_point,
                       group_point=lambda points, idx: t(ops.group_point(f32(points), idx.numpy())))
    sampling = _module("tf_ops.sampling.tf_sampling",
                       farthest_point_sample=lambda npoint, inp: t(ops.farthest_point_sample(npoint, f32(inp)), torch.int32),
                       gather_point=lambda inp, idx: t(ops.gather_point(f32(inp), idx.numpy())))
    pkg = _module("tf_ops")
    pkg.__path__ = []
    pkg.grouping, pkg.sampling = _module("tf_ops.grouping", tf_grouping=grouping), _module("tf_ops.sampling", tf_sampling=sampling)
    pkg.grouping.__path__, pkg.sampling.__path__ = [], []
    return {"tf_ops": pkg, "tf_ops.grouping": pkg.grouping, "tf_ops.sampling": pkg.sampling,
            "tf_ops.grouping.tf_grouping": grouping, "tf_ops.sampling.tf_sampling": sampling}


def install(reference_root="/root/reference"):
    for k in list(sys.modules):
        assert k.split(".")[0] not in ("models", "tensorflow"), "install() must run in a fresh interpreter (%s is loaded)" % k
    tf, slim = build_tensorflow()
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": tf.contrib, "tensorflow.contrib.slim": slim})
    sys.modules.update(build_tf_ops())
    sys.path.insert(0, reference_root)
    return tf

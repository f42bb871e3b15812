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

relu = torch.relu
softplus = F.softplus


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
            new_stats[scope + "/moving_mean"] = (mm - (1 - decay) * (mm - mean)).detach()
            new_stats[scope + "/moving_variance"] = (mv - (1 - decay) * (mv - var)).detach()
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


def conv2d(inputs, num_outputs, kernel_size, stride=[1, 1], padding='SAME', activation=relu, bn=True, bn_decay=None,
           is_training=None, scope=None, reuse=None, params=None, new_stats=None):
    """ 2D convolution with non-linear operation (layers.py:11-46): slim.conv2d WITH bias -> BN -> activation.

    inputs: (B,H,W,C).  Only the 1x1 / stride-1 kernels the model uses are implemented."""
    if list(kernel_size) != [1, 1] or list(stride) != [1, 1]:
        raise ValueError("conv2d: only kernel_size=[1,1], stride=[1,1] is supported (all the model uses)")
    if params is None or scope is None:
        raise ValueError("conv2d needs params= (flat dict keyed by TF scope names) and scope=")
    w, b = params[scope + "/conv2d/weights"], params[scope + "/conv2d/biases"]
    if w.shape[-1] != num_outputs or w.shape[-2] != inputs.shape[-1]:
        raise ValueError("conv2d: weight shape %s does not match (%d -> %d)"
                         % (tuple(w.shape), inputs.shape[-1], num_outputs))
    net = torch.matmul(inputs, w.reshape(w.shape[-2], w.shape[-1])) + b
    if bn:
        net = batch_norm_for_conv2d(net, bool(is_training), bn_decay, scope + "/bn", params, new_stats)
    if activation is not None:
        net = activation(net)
    return net


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

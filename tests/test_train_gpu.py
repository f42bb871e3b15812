"""Parity of the training-step CUDA ops (csrc/train.cu) with their plain statements: the attention-weighted triplet
loss of Feat3dNet.get_loss (reference models/feat3dnet.py:315-357) forward + backward, and TF-1 Adam (:359-375)."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import net as onet

pytestmark = pytest.mark.gpu


def pkg(name):
    return importlib.import_module("3dfeatnet_b200." + name)


def plain_loss(fa, fp, fn, att, margin):
    """The reference's statement, with torch ops (amin splits the gradient equally among ties like tf.reduce_min)."""
    layers = pkg("models.layers")
    bp = layers.pairwise_dist(fa, fp).amin(dim=2)
    bn = layers.pairwise_dist(fa, fn).amin(dim=2)
    if att is None:
        sp, sn = bp.mean(1), bn.mean(1)
    else:
        w = att / att.sum(dim=1)[:, None]
        sp, sn = (w * bp).sum(1), (w * bn).sum(1)
    return torch.clamp(sp - sn + margin, min=0.).mean()


@pytest.mark.parametrize("B,M,F,use_att,margin", [(6, 512, 32, True, 0.2), (3, 100, 16, False, 0.2), (2, 33, 128, True, 5.0),
                                                 (4, 64, 32, True, -10.0), (1, 1, 32, True, 0.2)])
def test_triplet_loss_forward_backward(cuda, B, M, F, use_att, margin):
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(B * 1000 + M)
    fa, fp, fn = (torch.nn.functional.normalize(torch.randn(B, M, F, generator=g), dim=-1).to(cuda).requires_grad_(True)
                  for _ in range(3))
    att = (torch.rand(B, M, generator=g) + 0.05).to(cuda).requires_grad_(True) if use_att else None
    ins = [fa, fp, fn] + ([att] if use_att else [])
    loss = layers.triplet_loss(fa, fp, fn, att, margin)
    grads = torch.autograd.grad(loss * 3.0, ins)
    ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    ref = plain_loss(ins64[0], ins64[1], ins64[2], ins64[3] if use_att else None, margin)
    rgrads = torch.autograd.grad(ref * 3.0, ins64)
    assert abs(loss.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    for name, a, b in zip(("dfa", "dfp", "dfn", "datt"), grads, rgrads):
        scale = b.abs().max().item() + 1e-12
        assert (a.double() - b).abs().max().item() <= 2e-5 * scale + 1e-9, name


def test_triplet_loss_ties_split_gradient(cuda):
    """Duplicate positives make exact ties in the row minima; the gradient is shared equally (tf.reduce_min)."""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(5)
    B, M, F = 2, 48, 32
    fa = torch.randn(B, M, F, generator=g)
    fp = torch.randn(B, M, F, generator=g)
    fp[:, 1::2] = fp[:, 0::2]                     # every positive twice -> every minimum is attained twice
    fn = torch.randn(B, M, F, generator=g)
    fn[:, 3] = fn[:, 7]
    att = torch.rand(B, M, generator=g) + 0.1
    ins = [t.to(cuda).requires_grad_(True) for t in (fa, fp, fn, att)]
    loss = layers.triplet_loss(*ins, 50.0)
    grads = torch.autograd.grad(loss, ins)
    ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    rgrads = torch.autograd.grad(plain_loss(*ins64, 50.0), ins64)
    for a, b in zip(grads, rgrads):
        assert (a.double() - b).abs().max().item() <= 2e-5 * (b.abs().max().item() + 1e-12)
    assert grads[1][:, 0::2].abs().sum() > 0 and torch.equal(grads[1][:, 0::2], grads[1][:, 1::2])


def test_triplet_loss_is_deterministic(cuda):
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(9)
    fa, fp, fn = (torch.randn(6, 512, 32, generator=g).to(cuda).requires_grad_(True) for _ in range(3))
    att = torch.rand(6, 512, generator=g).to(cuda).requires_grad_(True)
    runs = [torch.autograd.grad(layers.triplet_loss(fa, fp, fn, att, 0.2), [fa, fp, fn, att]) for _ in range(3)]
    for r in runs[1:]:
        assert all(torch.equal(x, y) for x, y in zip(r, runs[0]))


def test_adam_matches_tf_style_oracle(cuda):
    """Three updates of all 107 619 variables against oracle.net.adam_step (fp64)."""
    f3 = pkg("models.feat3dnet")
    params = onet.init_params(seed=2, randomize_bn=True)
    net = f3.Feat3dNet({'num_clusters': 32}, weights=params, device=cuda).train_mode()
    var = net.trainable_variables()
    names = list(var)
    oP = {k: torch.as_tensor(params[k]).double() for k in names}
    state = {}
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        grads = {k: torch.randn(var[k].shape, generator=g) * (10.0 ** (step - 2)) for k in names}
        loss = sum((var[k] * grads[k].to(cuda)).sum() for k in names)      # d loss / d var = grads
        flat = net.get_train_op(loss, lr=1e-3)
        assert torch.allclose(flat.cpu(), torch.cat([grads[k].reshape(-1) for k in names]), rtol=1e-6, atol=1e-7)
        onet.adam_step(oP, {k: v.double() for k, v in grads.items()}, state, lr=1e-3)
        for k in names:
            assert torch.allclose(var[k].detach().cpu().double(), oP[k], rtol=1e-5, atol=2e-6), (step, k)


def test_triplet_loss_rejects_bad_arguments(cuda):
    _lib = pkg("_lib")
    L = _lib.lib()
    assert L.f3d_triplet_loss(0, 4, 4, 0.2, None, None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert L.f3d_adam_step(0, None, 0, 1e-3, 0.9, 0.999, 1e-8, 1, 1.0, None, None) == -1


def plain_conv_bn(x, w, b, gamma, beta, relu_mask):
    """relu_mask: None (no ReLU) or the 0/1 mask of the device result -- the gradient of a ReLU is discontinuous at 0, so
    an element whose pre-activation is within rounding of 0 must be routed the same way in both computations."""
    z = x @ w + b
    mean, var = z.mean(0), z.var(0, unbiased=False)
    y = (z - mean) * torch.rsqrt(var + 1e-3) * gamma + beta
    return (y * relu_mask if relu_mask is not None else y), mean, var


@pytest.mark.parametrize("rows,cin,cout,use_relu", [(8192 + 37, 3, 64, True), (5000, 64, 128, True), (4096, 128, 256, True),
                                                    (777, 256, 128, True), (1000, 128, 64, True), (3000, 3, 32, True),
                                                    (3000, 32, 64, True), (3000, 128, 128, False), (999, 128, 32, False),
                                                    (1, 64, 16, True), (130, 16, 16, True), (70000, 128, 256, True),
                                                    (20011, 64, 64, True)])
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_conv_bn_train_forward_backward(cuda, rows, cin, cout, use_relu, precision, monkeypatch):
    """The training-mode layer (reference layers.py:11-46,225-272) against its fp64 torch statement + autograd, for the
    fp32 FFMA contractions and the tcgen05 ones (bf16 hi/lo split, three MMAs: ~1e-5 relative)."""
    layers = pkg("models.layers")
    monkeypatch.setattr(layers, "TRAIN_PRECISION", precision)
    g = torch.Generator().manual_seed(rows + cin * 7 + cout)
    x = torch.randn(rows, cin, generator=g) * 0.7 + 0.1
    w = torch.randn(cin, cout, generator=g) * (2.0 / cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    gamma = 1.0 + 0.2 * torch.randn(cout, generator=g)
    beta = 0.1 * torch.randn(cout, generator=g)
    gy = torch.randn(rows, cout, generator=g)
    ins = [t.to(cuda).requires_grad_(True) for t in (x, w, b, gamma, beta)]
    y, mean, var = layers.conv_bn_train(*ins, use_relu)
    grads = torch.autograd.grad((y * gy.to(cuda)).sum(), ins)
    ins64 = [t.double().requires_grad_(True) for t in (x, w, b, gamma, beta)]
    ry, rmean, rvar = plain_conv_bn(*ins64, (y.detach() > 0).cpu().double() if use_relu else None)
    rgrads = torch.autograd.grad((ry * gy.double()).sum(), ins64)
    tc = precision == "bf16x3"      # bf16 hi/lo split: every product carries ~2^-17 relative error
    assert torch.allclose(mean.cpu().double(), rmean, rtol=1e-4 if tc else 1e-5, atol=2e-5 if tc else 1e-6)
    assert torch.allclose(var.cpu().double(), rvar, rtol=2e-4 if tc else 1e-4, atol=1e-6 if tc else 1e-7)
    if rows > 1:
        assert (y.detach().cpu().double() - ry).abs().max().item() < 2e-5 * max(1.0, ry.abs().max().item())
    for name, a, r in zip(("dx", "dW", "db", "dgamma", "dbeta"), grads, rgrads):
        if name == "db":      # the bias in front of a batch-norm has a mathematically zero gradient: only rounding noise
            assert a.abs().max().item() < 1e-3 * (rgrads[1].abs().max().item() + 1e-6)
            continue
        scale = r.abs().max().item() + 1e-9
        assert (a.cpu().double() - r).abs().max().item() < 2e-4 * scale + 1e-6, name


def test_conv_bn_train_variance_with_a_large_mean(cuda, monkeypatch):
    """var = E[z^2] - mean^2 with |mean| = 20 std (a BN input far from centred): the per-CTA partial sums are combined and
    subtracted in fp64 (bn_stats_finalize_kernel), so what is left is the fp32 rounding of each partial (a few hundred rows).
    Batch statistics of the reference's tf.nn.moments (layers.py:250) in fp64 as the statement.  (At |mean| = 200 std the
    fp32 partials themselves lose the variance -- 6e-4 relative measured -- which only shifted per-thread sums would cure.)"""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(4)
    rows, cin, cout = 60000, 16, 16
    x = torch.randn(rows, cin, generator=g) * 0.05
    w = torch.eye(cin)
    b = torch.full((cout,), 1.0)
    gamma, beta = torch.ones(cout), torch.zeros(cout)
    z = x.double() @ w.double() + b.double()
    for precision in ("fp32", "bf16x3"):
        monkeypatch.setattr(layers, "TRAIN_PRECISION", precision)
        y, mean, var = layers.conv_bn_train(*(t.to(cuda) for t in (x, w, b, gamma, beta)), False)
        assert torch.allclose(mean.cpu().double(), z.mean(0), rtol=2e-6)
        rel = ((var.cpu().double() - z.var(0, unbiased=False)).abs() / z.var(0, unbiased=False)).max().item()
        assert rel < 2e-4, "%s: batch variance relative error %.2e" % (precision, rel)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_conv_bn_train_is_deterministic_and_skips_dx(cuda, precision, monkeypatch):
    layers = pkg("models.layers")
    monkeypatch.setattr(layers, "TRAIN_PRECISION", precision)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(20000, 64, generator=g).to(cuda)
    w, b, ga, be = (t.to(cuda).requires_grad_(True) for t in (torch.randn(64, 128, generator=g) * 0.2, torch.zeros(128),
                                                              torch.ones(128), torch.zeros(128)))
    gy = torch.randn(20000, 128, generator=g).to(cuda)
    runs = []
    for _ in range(2):
        y, _, _ = layers.conv_bn_train(x, w, b, ga, be, True)
        runs.append(torch.autograd.grad((y * gy).sum(), [w, b, ga, be]))
    assert all(torch.equal(a, c) for a, c in zip(*runs))


@pytest.mark.parametrize("shape", [(3, 40, 64, 256), (2, 7, 64, 64), (1, 5, 1, 128), (2, 3, 50, 4)])
def test_max_pool_samples_matches_reduce_max_and_shares_ties(cuda, shape):
    """tf.reduce_max over the sample axis (reference feat3dnet.py:138,147,182): value and tie-sharing gradient vs torch.amax."""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    x = torch.relu(x - 0.8)                       # many exact zeros => ties wherever a whole column is zero
    x[:, :, 1::2] = x[:, :, 0::2][:, :, :x[:, :, 1::2].shape[2]]  # and duplicated samples (ball-query padding)
    xc = x.to(cuda).requires_grad_(True)
    go = torch.randn(shape[0], shape[1], 1, shape[3], generator=g)
    out = layers.max_pool_samples(xc)
    (dx,) = torch.autograd.grad((out * go.to(cuda)).sum(), [xc])
    xr = x.clone().requires_grad_(True)
    ref = xr.amax(dim=2, keepdim=True)
    (rdx,) = torch.autograd.grad((ref * go).sum(), [xr])
    assert torch.equal(out.cpu(), ref.detach())
    assert torch.allclose(dx.cpu(), rdx, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("use_relu", [True, False])
def test_conv_bn_pool_fused_gradient_matches_unfused(cuda, precision, use_relu, monkeypatch):
    """conv2d(pool_samples=True) in training mode: pooled output and every gradient equal the unfused composition
    (layer, then max_pool_samples), which the other tests pin against the fp64 statement."""
    layers = pkg("models.layers")
    monkeypatch.setattr(layers, "TRAIN_PRECISION", precision)
    g = torch.Generator().manual_seed(11)
    B, M, S, cin, cout = 3, 50, 64, 64, 128
    x = torch.relu(torch.randn(B, M, S, cin, generator=g))
    x[:, :, 1::2] = x[:, :, 0::2]          # duplicated samples => tied maxima everywhere
    params = {"l/conv2d/weights": torch.randn(cin, cout, generator=g) * 0.2, "l/conv2d/biases": torch.zeros(cout),
              "l/bn/gamma": torch.rand(cout, generator=g) + 0.5, "l/bn/beta": torch.randn(cout, generator=g) * 0.1,
              "l/bn/moving_mean": torch.zeros(cout), "l/bn/moving_variance": torch.ones(cout)}
    go = torch.randn(B, M, 1, cout, generator=g).to(cuda)
    res = []
    for fused in (True, False):
        P = {k: v.to(cuda).requires_grad_(k.split("/")[-1] in ("weights", "biases", "gamma", "beta")) for k, v in params.items()}
        xc = x.to(cuda).requires_grad_(True)
        act = layers.relu if use_relu else None
        if fused:
            out = layers.conv2d(xc, cout, [1, 1], scope="l", is_training=True, activation=act, params=P, pool_samples=True)
        else:
            out = layers.max_pool_samples(layers.conv2d(xc, cout, [1, 1], scope="l", is_training=True, activation=act, params=P))
        leaves = [xc] + [P[k] for k in ("l/conv2d/weights", "l/bn/gamma", "l/bn/beta")]
        res.append((out.detach(), torch.autograd.grad((out * go).sum(), leaves)))
    assert torch.equal(res[0][0], res[1][0])
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * (b.abs().max().item() + 1e-9))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_conv_mid_without_concat_matches_the_concatenated_statement(cuda, precision, monkeypatch):
    """pointnet_sa_module's conv_mid (reference feat3dnet.py:60-75): concat([h, tile(max_S h)]) -> conv+BN -> max_S.
    conv2d(concat_pooled=, pool_samples=True) never builds the concat; value and all gradients must equal the literal
    graph (torch cat + matmul + batch statistics + amax) in fp64."""
    layers = pkg("models.layers")
    monkeypatch.setattr(layers, "TRAIN_PRECISION", precision)
    g = torch.Generator().manual_seed(21)
    B, M, S, c1, cout = 2, 40, 64, 64, 128
    h = torch.relu(torch.randn(B, M, S, c1, generator=g))
    W = torch.randn(2 * c1, cout, generator=g) * 0.15
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    go = torch.randn(B, M, 1, cout, generator=g)
    P = {"l/conv2d/weights": W.to(cuda).requires_grad_(True), "l/conv2d/biases": torch.zeros(cout, device=cuda, requires_grad=True),
         "l/bn/gamma": gamma.to(cuda).requires_grad_(True), "l/bn/beta": beta.to(cuda).requires_grad_(True),
         "l/bn/moving_mean": torch.zeros(cout, device=cuda), "l/bn/moving_variance": torch.ones(cout, device=cuda)}
    hc = h.to(cuda).requires_grad_(True)
    pooled = layers.max_pool_samples(hc)
    out = layers.conv2d(hc, cout, [1, 1], scope="l", is_training=True, activation=None, params=P, pool_samples=True, concat_pooled=pooled)
    grads = torch.autograd.grad((out * go.to(cuda)).sum(), [hc, P["l/conv2d/weights"], P["l/bn/gamma"], P["l/bn/beta"]])
    # literal graph in fp64
    h64, W64, ga64, be64 = (t.double().requires_grad_(True) for t in (h, W, gamma, beta))
    q = h64.amax(dim=2, keepdim=True)
    z = torch.cat((h64, q.expand(-1, -1, S, -1)), dim=3) @ W64
    mean, var = z.mean(dim=(0, 1, 2)), z.var(dim=(0, 1, 2), unbiased=False)
    ref = ((z - mean) * torch.rsqrt(var + 1e-3) * ga64 + be64).amax(dim=2, keepdim=True)
    rgrads = torch.autograd.grad((ref * go.double()).sum(), [h64, W64, ga64, be64])
    assert (out.detach().cpu().double() - ref).abs().max().item() < 5e-5 * max(1.0, ref.abs().max().item())
    for name, a, r in zip(("dh", "dW", "dgamma", "dbeta"), grads, rgrads):
        assert (a.cpu().double() - r).abs().max().item() < 3e-4 * (r.abs().max().item() + 1e-9) + 1e-6, name


@pytest.mark.parametrize("cin,cout,use_relu,mid", [(128, 256, True, False), (64, 128, False, True), (64, 128, True, False), (128, 128, True, True)])
def test_pool_only_layer_dz_formed_inside_the_contractions(cuda, cin, cout, use_relu, mid):
    """Pool-only layers (detector conv2, descriptor conv_mid) on the tensor-core path: wgrad and dgrad form dz = BN-backward gradient
    from z + the pooled tensors in their operand converters (csrc/dz_source.cuh) instead of reading the (rows, cout) tensor
    bn_bwd_apply_kernel would write.  Same arithmetic, so dW / dx / dgamma / dbeta carry the SAME BITS as the three-kernel path; db and the
    gradient of the per-group term (sums of dz in a different order) agree to rounding."""
    layers, lib_mod = pkg("models.layers"), pkg("_lib")
    g = torch.Generator().manual_seed(5)
    B, M, S = 2, 77, 64
    x = torch.relu(torch.randn(B, M, S, cin, generator=g))
    x[:, :, 1::2] = x[:, :, 0::2]          # duplicated samples => tied maxima everywhere
    params = {"l/conv2d/weights": torch.randn(cin * (2 if mid else 1), cout, generator=g) * 0.2, "l/conv2d/biases": torch.randn(cout, generator=g) * 0.1,
              "l/bn/gamma": torch.rand(cout, generator=g) + 0.5, "l/bn/beta": torch.randn(cout, generator=g) * 0.1,
              "l/bn/moving_mean": torch.zeros(cout), "l/bn/moving_variance": torch.ones(cout)}
    go = torch.randn(B, M, 1, cout, generator=g).to(cuda)
    res = []
    L = lib_mod.lib()
    for fuse in (1, 0):
        prev = L.f3d_debug_set_fuse_dz(fuse)
        try:
            P = {k: v.to(cuda).requires_grad_(k.split("/")[-1] in ("weights", "biases", "gamma", "beta")) for k, v in params.items()}
            xc = x.to(cuda).requires_grad_(True)
            pooled_in = layers.max_pool_samples(xc) if mid else None
            L.f3d_reset_launch_count()
            out = layers.conv2d(xc, cout, [1, 1], scope="l", is_training=True, activation=layers.relu if use_relu else None, params=P,
                                pool_samples=True, concat_pooled=pooled_in)
            leaves = [xc] + [P[k] for k in ("l/conv2d/weights", "l/bn/gamma", "l/bn/beta", "l/conv2d/biases")]
            res.append((out.detach(), torch.autograd.grad((out * go).sum(), leaves), L.f3d_launch_count()))
        finally:
            L.f3d_debug_set_fuse_dz(prev)
    assert torch.equal(res[0][0], res[1][0])
    assert res[0][2] < res[1][2], "the fused path must launch fewer kernels (no bn_bwd_apply, no group_sum)"
    names = ("dx", "dW", "dgamma", "dbeta", "db")
    for name, a, b in zip(names, res[0][1], res[1][1]):
        if name == "db":
            # db = column sums of dz, which the BN backward makes ZERO in exact arithmetic: both paths return rounding noise of the
            # order eps * sum|dz| (measured 5e-6), summed in a different order
            assert a.abs().max().item() < 5e-5 and b.abs().max().item() < 5e-5, name
        elif mid and name in ("dx", "dW"):
            # with the per-group term, dx and the lower half of dW receive d(per-group term) = group sums of dz (different order), and those
            # pass through the 2-way split contractions of layers.linear_rows: rounding-level input differences come out at the split's 1e-5
            assert torch.allclose(a, b, rtol=1e-4, atol=2e-5 * (b.abs().max().item() + 1e-9)), name
        else:
            assert torch.equal(a, b), name


@pytest.mark.parametrize("rows,k,nout", [(9216, 64, 128), (1000, 64, 256), (77, 8, 16), (4096, 128, 128)])
def test_linear_rows_on_the_tensor_cores(cuda, rows, k, nout):
    """layers.linear_rows (f3d_linear_forward / _backward: the per-cluster term of conv_mid, formerly a cuBLAS call) against fp64:
    forward to fp32 rounding (3-way split), gradients to the 2-way split's 1e-5."""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(rows + k)
    x = torch.randn(rows, k, generator=g).to(cuda).requires_grad_(True)
    w = (torch.randn(k, nout, generator=g) * 0.2).to(cuda).requires_grad_(True)
    go = torch.randn(rows, nout, generator=g).to(cuda)
    out = layers.linear_rows(x, w)
    assert out.grad_fn is not None and type(out.grad_fn).__name__.startswith("_LinearTC")
    dx, dw = torch.autograd.grad((out * go).sum(), (x, w))
    xd, wd = x.detach().double(), w.detach().double()
    ref = xd @ wd
    assert (out.detach().double() - ref).abs().max().item() < 5e-6 * ref.abs().max().item()  # fp32 accumulation over k terms
    rdx, rdw = go.double() @ wd.t(), xd.t() @ go.double()
    assert (dx.double() - rdx).abs().max().item() < 3e-5 * rdx.abs().max().item()
    assert (dw.double() - rdw).abs().max().item() < 3e-5 * rdw.abs().max().item()


@pytest.mark.parametrize("cin,cout,use_relu,mid,neg_gamma", [(128, 256, True, False, False), (64, 128, False, True, True), (64, 128, True, False, True),
                                                             (128, 128, False, False, False)])
def test_pool_statistics_taken_in_the_contraction_epilogue(cuda, cin, cout, use_relu, mid, neg_gamma):
    """Pool-only layers on the tensor-core path: the forward contraction's epilogue records, per half group and channel, the largest z (the
    smallest where gamma < 0), its multiplicity and the runner-up; pool_from_extremes_kernel turns them into the pooled maximum and the tie
    counts once the BN statistics exist (re-counting the rare groups where two distinct z round to one activation) -- against the pass over z
    of bn_apply_pool_kernel: same pooled output, same gradients, bit for bit, on inputs full of exact ties and with channels that the ReLU
    switches off."""
    layers, lib_mod = pkg("models.layers"), pkg("_lib")
    g = torch.Generator().manual_seed(17)
    B, M, S = 2, 90, 64
    x = torch.relu(torch.randn(B, M, S, cin, generator=g))
    x[:, :, 1::2] = x[:, :, 0::2]          # duplicated samples => tied maxima everywhere
    x[:, 10:20] = x[:, 10:20, :1]          # whole groups of identical rows
    gamma = torch.rand(cout, generator=g) + 0.5
    if neg_gamma:
        gamma[::3] = -gamma[::3]
        gamma[5] = 0.0
    beta = torch.randn(cout, generator=g) * 0.5
    beta[::4] -= 3.0                        # channels whose activation is (almost) always clamped by the ReLU
    params = {"l/conv2d/weights": torch.randn(cin * (2 if mid else 1), cout, generator=g) * 0.2, "l/conv2d/biases": torch.randn(cout, generator=g) * 0.1,
              "l/bn/gamma": gamma, "l/bn/beta": beta, "l/bn/moving_mean": torch.zeros(cout), "l/bn/moving_variance": torch.ones(cout)}
    go = torch.randn(B, M, 1, cout, generator=g).to(cuda)
    res = []
    L = lib_mod.lib()
    for epi in (1, 0):
        prev = L.f3d_debug_set_epilogue_pool(epi)
        try:
            P = {k: v.to(cuda).requires_grad_(k.split("/")[-1] in ("weights", "biases", "gamma", "beta")) for k, v in params.items()}
            xc = x.to(cuda).requires_grad_(True)
            pooled_in = layers.max_pool_samples(xc) if mid else None
            L.f3d_reset_launch_count()
            out = layers.conv2d(xc, cout, [1, 1], scope="l", is_training=True, activation=layers.relu if use_relu else None, params=P,
                                pool_samples=True, concat_pooled=pooled_in)
            leaves = [xc] + [P[k] for k in ("l/conv2d/weights", "l/bn/gamma", "l/bn/beta", "l/conv2d/biases")]
            res.append((out.detach(), torch.autograd.grad((out * go).sum(), leaves)))
        finally:
            L.f3d_debug_set_epilogue_pool(prev)
    assert torch.equal(res[0][0], res[1][0])
    for name, a, b in zip(("dx", "dW", "dgamma", "dbeta", "db"), res[0][1], res[1][1]):
        assert torch.equal(a, b), name


@pytest.mark.parametrize("B,N,M,fdim", [(2, 2048, 77, 32), (3, 4096, 128, 128)])
def test_chained_layers_carry_the_bits_of_the_materialised_path(cuda, B, N, M, fdim):
    """The per-point MLP chains (detector conv0 -> conv1 -> conv2, descriptor conv0 -> conv1 -> conv_mid) with their intermediate
    activations left unmaterialised (layers.DeferredActivation: the consumers form relu(z * scale + shift) in their operand converters,
    the layer that feeds the pool and conv_mid sums its two gradients on the fly) against the same step with every activation written
    by bn_apply and read back: same arithmetic, so the loss, the gradient of all trainable variables and the BN moments carry the SAME
    BITS -- while the chained step launches fewer kernels (no bn_apply of the four chained layers, no max-pool forward / backward,
    no gradient add)."""
    f3, layers, synth, lib_mod = pkg("models.feat3dnet"), pkg("models.layers"), pkg("synth"), pkg("_lib")
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(cuda) for s in (31, 32, 33))
    params = onet.init_params(seed=4, randomize_bn=True, feature_dim=fdim)
    L = lib_mod.lib()

    def run(chain):
        layers.CHAIN_ACTIVATIONS = chain
        try:
            net = f3.Feat3dNet({'num_clusters': M, 'feature_dim': fdim}, weights=params, device=cuda).train_mode()
            torch.cuda.synchronize()
            L.f3d_reset_launch_count()
            xyz, feats, att, ep = net.get_train_model(a, p, n, True)
            loss, ep = net.get_loss(xyz, feats, att, ep)
            flat = net.get_train_op(loss, lr=1e-5, end_points=ep)
            torch.cuda.synchronize()
            launches = L.f3d_launch_count()
            stats = {k: v.stat.clone() for k, v in ep['bn_updates'].items()}
            return loss.detach().item(), flat.detach().clone(), stats, launches
        finally:
            layers.CHAIN_ACTIVATIONS = True

    l1, g1, s1, n1 = run(True)
    l0, g0, s0, n0 = run(False)
    assert l1 == l0
    assert torch.equal(g1, g0)
    assert s1.keys() == s0.keys() and all(torch.equal(s1[k], s0[k]) for k in s1)
    assert n1 < n0, (n1, n0)


def test_c4_training_step_fused_vs_torch_layers_and_deterministic(cuda):
    """BASELINE configs[3] size (6 triplets x 4096 points, 512 clusters x 64 samples): the CUDA training layers against the
    op-by-op torch fp32 statement of the same graph (loss, gradient direction), and bit-reproducibility of the CUDA path."""
    f3, layers, synth = pkg("models.feat3dnet"), pkg("models.layers"), pkg("synth")
    B, N, M = 6, 4096, 512
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(cuda) for s in (11, 12, 13))
    params = onet.init_params(seed=1, randomize_bn=True)
    torch.backends.cuda.matmul.allow_tf32 = False

    def run(fused, precision="bf16x3"):
        layers.FUSED_TRAINING, layers.TRAIN_PRECISION = fused, precision
        try:
            net = f3.Feat3dNet({'num_clusters': M, 'fused_loss': fused}, weights=params, device=cuda).train_mode()
            xyz, feats, att, ep = net.get_train_model(a, p, n, True)
            loss, ep = net.get_loss(xyz, feats, att, ep)
            flat = net.get_train_op(loss, lr=1e-5, end_points=ep)
            return loss.detach().item(), flat.detach().clone()
        finally:
            layers.FUSED_TRAINING, layers.TRAIN_PRECISION = True, "bf16x3"

    l_ref, g_ref = run(False)
    l_tc, g_tc = run(True, "bf16x3")
    l_tc2, g_tc2 = run(True, "bf16x3")
    l_f32, g_f32 = run(True, "fp32")
    assert l_tc == l_tc2 and torch.equal(g_tc, g_tc2)                      # no atomics anywhere
    for l, g in ((l_tc, g_tc), (l_f32, g_f32)):
        assert abs(l - l_ref) < 2e-4 * max(1.0, abs(l_ref))
        assert torch.isfinite(g).all() and g.numel() == 107619
        # ReLU / max-pool routing flips on rounding-level differences, so the comparison is a direction + scale one
        assert torch.nn.functional.cosine_similarity(g.double(), g_ref.double(), dim=0).item() > 0.999
        assert abs(g.norm().item() / g_ref.norm().item() - 1.0) < 2e-2


def test_captured_train_step_equals_eager_steps(cuda):
    """Feat3dNet.capture_train_step: N replays of the CUDA graph == N eager steps, bit for bit (weights, Adam moments, BN
    shadows) -- the kernels are deterministic and the Adam step count is read on the device."""
    f3, synth = pkg("models.feat3dnet"), pkg("synth")
    B, N, M = 2, 2048, 64
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(cuda) for s in (21, 22, 23))
    params = onet.init_params(seed=8, randomize_bn=True)

    def eager(steps):
        net = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
        for _ in range(steps):
            xyz, feats, att, ep = net.get_train_model(a, p, n, True)
            loss, ep = net.get_loss(xyz, feats, att, ep)
            net.get_train_op(loss, lr=1e-3, end_points=ep)
        return net, loss.detach().clone()

    net_e, loss_e = eager(5)
    net_g = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    replay = net_g.capture_train_step(a, p, n, lr=1e-3, warmup=2)      # 2 eager + 1 captured-run... capture itself does not execute
    for _ in range(3):
        loss_g = replay()
    torch.cuda.synchronize()
    assert int(net_g._adam["t_dev"].item()) == 5 == int(net_e._adam["t_dev"].item())
    assert torch.equal(loss_g, loss_e)
    for k in net_e.weights:
        assert torch.equal(net_e.weights[k], net_g.weights[k]), k
    assert torch.equal(net_e._adam["m"], net_g._adam["m"]) and torch.equal(net_e._adam["v"], net_g._adam["v"])


def test_pipelined_train_step_equals_the_serial_step_on_a_sequence_of_batches(cuda):
    """capture_train_step(pipelined=True): farthest point sampling + ball query of the NEXT batch run on a side stream beside the backward
    pass of the current step.  replay(b_{k+1}) runs step k on the batch staged by the previous call and stages b_{k+1}.  On a sequence of
    DIFFERENT batches the losses, the weights, the Adam moments and the BN shadows equal those of the serial graph fed the same sequence,
    bit for bit."""
    f3, synth = pkg("models.feat3dnet"), pkg("synth")
    B, N, M = 2, 2048, 64
    batches = [[torch.as_tensor(synth.make_batch(B, N, seed0=50 + 10 * k + s)).to(cuda) for s in range(3)] for k in range(4)]
    params = onet.init_params(seed=12, randomize_bn=True)

    net_s = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    rs = net_s.capture_train_step(*batches[0], lr=1e-3, warmup=1)      # the warm-up step runs on batch 0
    losses_s = [rs(*batches[k]).clone() for k in range(4)]

    net_p = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    rp = net_p.capture_train_step(*batches[0], lr=1e-3, warmup=1, pipelined=True)
    assert rp.pipelined
    losses_p = []
    for k in range(4):
        nxt = batches[k + 1] if k + 1 < 4 else (None, None, None)
        losses_p.append(rp(*nxt).clone())                               # runs batch k, stages batch k + 1
    torch.cuda.synchronize()
    for k in range(4):
        assert torch.equal(losses_p[k], losses_s[k]), k
    assert int(net_p._adam["t_dev"].item()) == int(net_s._adam["t_dev"].item()) == 5
    for k in net_s.weights:
        assert torch.equal(net_s.weights[k], net_p.weights[k]), k
    assert torch.equal(net_s._adam["m"], net_p._adam["m"]) and torch.equal(net_s._adam["v"], net_p._adam["v"])


def test_eval_forward_between_graph_replays_sees_the_updated_weights(cuda):
    """replay -> eval -> replay -> eval: the folded eval-mode weight copy is invalidated by every replay (the graph updates the
    weights and the BN shadows without going through get_train_op's Python), so each eval forward equals the one of the eager
    sequence at the same step."""
    f3, synth = pkg("models.feat3dnet"), pkg("synth")
    B, N, M = 2, 2048, 64
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(cuda) for s in (31, 32, 33))
    params = onet.init_params(seed=9, randomize_bn=True)
    probe = torch.as_tensor(synth.make_batch(1, 2048, seed0=40)).to(cuda)

    def eval_out(net):
        kp, feat, att, ep = net.get_inference_model(probe, False)
        return feat.clone(), att.clone(), ep["orientation"].clone()

    net_e = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    outs_e = []
    for step in range(4):
        xyz, feats, att, ep = net_e.get_train_model(a, p, n, True)
        loss, ep = net_e.get_loss(xyz, feats, att, ep)
        net_e.get_train_op(loss, lr=1e-2, end_points=ep)
        if step >= 2:
            outs_e.append(eval_out(net_e))
    net_g = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    replay = net_g.capture_train_step(a, p, n, lr=1e-2, warmup=2)
    outs_g = []
    for _ in range(2):
        replay()
        outs_g.append(eval_out(net_g))
    assert not torch.equal(outs_g[0][0], outs_g[1][0]), "the weights moved between the two eval forwards"
    for e, g in zip(outs_e, outs_g):
        for te, tg in zip(e, g):
            assert torch.equal(te, tg)


def test_train_loop_follows_the_reference_schedule(cuda, tmp_path):
    """train.py:93-184 in miniature on an on-disk dataset: epochs end at the first short batch, checkpoints every n steps,
    validation at step 1 and every n steps, and the saved checkpoint restores into a fresh model"""
    trainer, dg, f3, ck = pkg("trainer"), pkg("data.datagenerator"), pkg("models.feat3dnet"), pkg("checkpoint")
    rng = np.random.default_rng(0)
    n_clouds, lines = 7, []
    (tmp_path / "train").mkdir()
    for i in range(n_clouds):
        cloud = np.concatenate([rng.uniform(-12, 12, (1500, 3)) * [1, 1, 0.2], np.zeros((1500, 3))], axis=1).astype(np.float32)
        cloud.tofile(str(tmp_path / "train" / ("%d.bin" % i)))
        lines.append("%d.bin | %d | %d" % (i, (i + 1) % n_clouds, (i + 2) % n_clouds))
    (tmp_path / "train" / "train.txt").write_text("\n".join(lines) + "\n")
    (tmp_path / "clusters").mkdir()
    gts = []
    for i in range(6):
        a = np.concatenate([rng.normal(0, 1.0, (600, 3)), np.zeros((600, 3))], axis=1).astype(np.float32)
        b = a.copy() if i % 2 else np.concatenate([rng.normal(0, 1.0, (600, 3)) * [2, 0.5, 1], np.zeros((600, 3))], axis=1).astype(np.float32)
        a.tofile(str(tmp_path / "clusters" / ("%d_0.bin" % i)))
        b.tofile(str(tmp_path / "clusters" / ("%d_1.bin" % i)))
        gts.append((i, i % 2))
    gen = dg.DataGenerator(str(tmp_path / "train" / "train.txt"), num_cols=6, seed=1)
    net = f3.Feat3dNet({'num_clusters': 64}, device=cuda, seed=0, precision="fp32").train_mode()
    hist = trainer.train(net, gen, num_epochs=2, batch_size=2, num_points=1024, lr=1e-3, checkpoint_dir=str(tmp_path / "ckpt"),
                         checkpoint_every_n_steps=3, val_folder=str(tmp_path / "clusters"), val_groundtruths=gts, validate_every_n_steps=4)
    assert hist["steps"] == 6  # 7 clouds, batch 2: 3 full batches per epoch, the 1-cloud remainder ends the epoch
    assert len(hist["losses"]) == 6 and all(np.isfinite(hist["losses"]))
    assert [s for s, _ in hist["fp_rates"]] == [1, 4] and all(0.0 <= r <= 1.0 for _, r in hist["fp_rates"])
    assert [os.path.basename(p) for p in hist["checkpoints"]] == ["checkpoint.ckpt-3.npz", "checkpoint.ckpt-6.npz"]
    fresh = f3.Feat3dNet({'num_clusters': 64}, device=cuda, seed=9, precision="fp32")
    ck.initialize_model(fresh, hist["checkpoints"][-1])
    for k, v in net.weights.items():
        assert torch.equal(fresh.weights[k].detach().float().cpu(), v.detach().float().cpu()), k


@pytest.mark.parametrize("rows,k", [(9216, 64), (777, 32), (50, 128)])
def test_detector_heads_op_matches_the_torch_statement(cuda, rows, k):
    """feat3dnet.py:142-149 as one CUDA op: forward values and every gradient against the op-by-op statement in fp64; two runs
    of the backward give identical bits (fixed-order reductions)"""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(rows + k)
    h = torch.relu(torch.randn(rows, k, generator=g, dtype=torch.float64))
    wa, ba = torch.randn(k, generator=g, dtype=torch.float64) * 0.3, torch.randn(1, generator=g, dtype=torch.float64)
    wo, bo = torch.randn(k, 2, generator=g, dtype=torch.float64) * 0.3, torch.randn(2, generator=g, dtype=torch.float64) * 0.1
    ga, go = torch.randn(rows, generator=g, dtype=torch.float64), torch.randn(rows, generator=g, dtype=torch.float64)
    ref_in = [t.clone().requires_grad_(True) for t in (h, wa, ba, wo, bo)]
    a = torch.nn.functional.softplus(ref_in[0] @ ref_in[1] + ref_in[2])
    xy = ref_in[0] @ ref_in[3] + ref_in[4]
    xy = xy * torch.rsqrt(torch.clamp((xy * xy).sum(1, keepdim=True), min=1e-8))
    o = torch.atan2(xy[:, 1], xy[:, 0])
    ref_g = torch.autograd.grad((a * ga).sum() + (o * go).sum(), ref_in)
    outs = []
    for _ in range(2):
        dev_in = [t.float().to(cuda).requires_grad_(True) for t in (h, wa, ba, wo, bo)]
        att, ori = layers._DetectorHeads.apply(*dev_in)
        grads = torch.autograd.grad((att * ga.float().to(cuda)).sum() + (ori * go.float().to(cuda)).sum(), dev_in)
        outs.append((att, ori, grads))
    att, ori, grads = outs[0]
    assert torch.allclose(att.cpu().double(), a.detach(), rtol=1e-5, atol=1e-6)
    d = ori.cpu().double() - o.detach()
    assert torch.atan2(torch.sin(d), torch.cos(d)).abs().max() < 1e-5
    for name, got, want in zip(("dh", "dw_att", "db_att", "dw_ori", "db_ori"), grads, ref_g):
        scale = want.abs().max().item() + 1e-12
        assert (got.cpu().double() - want).abs().max().item() < 2e-5 * scale + 1e-6, name
    for x, y in zip(outs[0][2], outs[1][2]):
        assert torch.equal(x, y)

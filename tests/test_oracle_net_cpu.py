"""CPU tests (-m "not gpu") of oracle/net.py, the statement the network kernels are checked against.  TensorFlow is not
available offline and the reference holds no golden vector for these layers (parity unpinned), so the
oracle is cross-checked against a SECOND, independent statement of the same documented semantics: PyTorch's own library
modules (nn.Conv2d, nn.BatchNorm2d, F.normalize, cdist, optim.Adam), all in float64."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import net as onet


def _p64(seed, randomize_bn=True):
    return onet.to_torch(onet.init_params(seed=seed, randomize_bn=randomize_bn), torch.float64)


def _torch_layer(P, scope, cin, cout, bn):
    conv = torch.nn.Conv2d(cin, cout, 1).double()
    with torch.no_grad():
        conv.weight.copy_(P[scope + "/conv2d/weights"].t().reshape(cout, cin, 1, 1))  # TF HWIO (1,1,Cin,Cout) -> OIHW
        conv.bias.copy_(P[scope + "/conv2d/biases"])
    norm = None
    if bn:
        norm = torch.nn.BatchNorm2d(cout, eps=onet.BN_EPS, momentum=1 - onet.BN_DECAY).double()
        with torch.no_grad():
            norm.weight.copy_(P[scope + "/bn/gamma"]); norm.bias.copy_(P[scope + "/bn/beta"])
            norm.running_mean.copy_(P[scope + "/bn/moving_mean"]); norm.running_var.copy_(P[scope + "/bn/moving_variance"])
    return conv, norm


def test_conv_bn_layers_agree_with_torch_modules():
    """layers.py:11-46,225-272 as restated by oracle.net.conv2d == Conv2d(1x1, bias) -> BatchNorm2d(eps=1e-3, momentum=0.1)
    -> ReLU on NCHW tensors, in eval mode (EMA shadows) and in training mode (batch moments + shadow update)."""
    P = _p64(3)
    g = torch.Generator().manual_seed(0)
    for scope, cin, cout, bn in onet.DET_LAYERS[:5] + onet.desc_layers(32):
        x = torch.randn((2, 5, 7, cin), generator=g, dtype=torch.float64)  # NHWC like the TF graph
        conv, norm = _torch_layer(P, scope, cin, cout, bn)
        for training in (False, True):
            norm.train(training)
            stats = {}
            got = onet.conv2d(x, P, scope, bn, "relu", training, stats)
            want = torch.relu(norm(conv(x.permute(0, 3, 1, 2)))).permute(0, 2, 3, 1)
            assert torch.allclose(got, want, rtol=1e-10, atol=1e-10), (scope, training)
            if training:
                n = x.numel() // cin
                assert torch.allclose(stats[scope + "/bn/moving_mean"], norm.running_mean, rtol=1e-10, atol=1e-12)
                # TF's shadow takes the population variance, torch's running_var the unbiased one: n/(n-1) apart
                mv = P[scope + "/bn/moving_variance"]
                batch_var = (stats[scope + "/bn/moving_variance"] - onet.BN_DECAY * mv) / (1 - onet.BN_DECAY)
                torch_batch_var = (norm.running_var - onet.BN_DECAY * mv) / (1 - onet.BN_DECAY)
                assert torch.allclose(batch_var * n / (n - 1), torch_batch_var, rtol=1e-9, atol=1e-12)


def test_heads_normalisation_and_rotation_agree_with_library_ops():
    """feat3dnet.py:134-149,185 and pointnet_common.py:110-120: softplus attention, l2_normalize (x * rsqrt(max(|x|^2, 1e-8)) ==
    F.normalize(eps=1e-4)), atan2 orientation, and the per-cluster rotation written as the reference's batched matmul."""
    P = _p64(4)
    g = torch.Generator().manual_seed(1)
    h = torch.randn((2, 6, 1, 64), generator=g, dtype=torch.float64)
    att = onet.conv2d(h, P, "detection/attention", False, "softplus")[:, :, 0, 0]
    lin = h[:, :, 0, :] @ P["detection/attention/conv2d/weights"] + P["detection/attention/conv2d/biases"]
    assert torch.allclose(att, torch.log1p(torch.exp(lin[:, :, 0])), rtol=1e-12) and (att > 0).all()
    v = torch.randn((2, 6, 5), generator=g, dtype=torch.float64)
    v[0, 0] = 0.0
    v[0, 1] = 1e-6  # below the clamp: divided by sqrt(1e-8), not by its own norm
    assert torch.allclose(onet.l2_normalize(v, 2, 1e-8), F.normalize(v, dim=2, eps=1e-4), rtol=1e-12, atol=0)
    # rotation: grouped_xyz @ [[c,s,0],[-s,c,0],[0,0,1]] per cluster (pointnet_common.py:112-119)
    xyz = np.random.default_rng(2).uniform(-3, 3, (2, 400, 3)).astype(np.float32)
    kp = xyz[:, :6, :].copy()
    ori = torch.rand((2, 6), generator=g, dtype=torch.float64) * 6.0 - 3.0
    plain = onet.descriptor(xyz, P, kp, None, 2.0, 16, dtype=torch.float64)
    turned = onet.descriptor(xyz, P, kp, ori, 2.0, 16, dtype=torch.float64)
    c, s, o, z = torch.cos(ori), torch.sin(ori), torch.ones_like(ori), torch.zeros_like(ori)
    R = torch.stack([c, s, z, -s, c, z, z, z, o], dim=2).reshape(2, 6, 3, 3)
    assert torch.allclose(turned["rotated_xyz"], torch.matmul(plain["rotated_xyz"], R), rtol=1e-12, atol=1e-14)
    assert torch.allclose(turned["features"].norm(dim=2), torch.ones(2, 6, dtype=torch.float64), rtol=1e-9)


def test_descriptor_graph_agrees_with_a_module_statement():
    """pointnet_sa_module (feat3dnet.py:9-87) built from torch modules on NCHW tensors -- conv0, conv1, max over samples, tile,
    concat, conv_mid_0 (no ReLU), max, conv_post_0 (no ReLU), normalise -- against oracle.net.descriptor (eval and training)."""
    P = _p64(6)
    xyz = np.random.default_rng(5).uniform(-4, 4, (3, 600, 3)).astype(np.float32)
    kp = xyz[:, :10, :].copy()
    for training in (False, True):
        out = onet.descriptor(xyz, P, kp, None, 2.0, 32, is_training=training, dtype=torch.float64)
        x = out["rotated_xyz"].permute(0, 3, 1, 2)  # (B,3,M,S)
        mods = {sc: _torch_layer(P, sc, cin, cout, bn) for sc, cin, cout, bn in onet.desc_layers(32)}
        for conv, norm in mods.values():
            norm.train(training)

        def layer(sc, t, relu):
            conv, norm = mods["description/layer1/" + sc]
            t = norm(conv(t))
            return torch.relu(t) if relu else t

        h = layer("conv1", layer("conv0", x, True), True)
        pooled = F.max_pool2d(h, kernel_size=(1, h.shape[3]))
        h = torch.cat([h, pooled.expand(-1, -1, -1, h.shape[3])], dim=1)
        h = F.max_pool2d(layer("conv_mid_0", h, False), kernel_size=(1, h.shape[3]))
        feat = F.normalize(layer("conv_post_0", h, False)[:, :, :, 0].permute(0, 2, 1), dim=2, eps=1e-4)
        assert torch.allclose(out["features"], feat, rtol=1e-9, atol=1e-11), training


def test_detector_graph_agrees_with_a_module_statement():
    """feature_detection_module (feat3dnet.py:90-151) from torch modules: conv0..2 (+BN+ReLU), max over samples, conv_post_0/1,
    softplus attention head, l2-normalised orientation head -> atan2, against oracle.net.detector (eval and training)."""
    P = _p64(8)
    xyz = np.random.default_rng(9).uniform(-4, 4, (2, 500, 3)).astype(np.float32)
    for training in (False, True):
        out = onet.detector(xyz, P, 12, 2.0, 32, is_training=training, dtype=torch.float64)
        mods = {sc: _torch_layer(P, sc, cin, cout, bn) for sc, cin, cout, bn in onet.DET_LAYERS}
        h = out["grouped_xyz"].permute(0, 3, 1, 2)
        for i, sc in enumerate(["conv0", "conv1", "conv2", "conv_post_0", "conv_post_1"]):
            conv, norm = mods["detection/" + sc]
            norm.train(training)
            h = torch.relu(norm(conv(h)))
            if i == 2:
                h = F.max_pool2d(h, kernel_size=(1, h.shape[3]))
        att = F.softplus(mods["detection/attention"][0](h))[:, 0, :, 0]
        o = F.normalize(mods["detection/orientation"][0](h)[:, :, :, 0], dim=1, eps=1e-4)
        assert torch.allclose(out["attention"], att, rtol=1e-9, atol=1e-12), training
        assert torch.allclose(out["orientation"], torch.atan2(o[:, 1], o[:, 0]), rtol=1e-9, atol=1e-11), training


def test_triplet_loss_agrees_with_cdist_statement():
    """get_loss (feat3dnet.py:315-357): squared distances via torch.cdist, nearest neighbour per anchor keypoint, attention-
    normalised sums, hinge with the margin, batch mean; also the uniform-weight (Attention=False) form."""
    g = torch.Generator().manual_seed(2)
    fa, fp, fn = (F.normalize(torch.randn((4, 30, 32), generator=g, dtype=torch.float64), dim=2) for _ in range(3))
    att = torch.rand((4, 30), generator=g, dtype=torch.float64) + 0.05
    best_p = (torch.cdist(fa, fp) ** 2).min(dim=2).values
    best_n = (torch.cdist(fa, fn) ** 2).min(dim=2).values
    assert torch.allclose(onet.pairwise_dist(fa, fp), torch.cdist(fa, fp) ** 2, rtol=1e-9, atol=1e-12)
    w = att / att.sum(1, keepdim=True)
    want = F.relu(((best_p - best_n) * w).sum(1) + 0.2).mean()
    assert torch.allclose(onet.triplet_loss(fa, fp, fn, att, 0.2, True), want, rtol=1e-12)
    want_u = F.relu((best_p - best_n).mean(1) + 0.2).mean()
    assert torch.allclose(onet.triplet_loss(fa, fp, fn, att, 0.2, False), want_u, rtol=1e-12)
    # margin large enough that the hinge is active, small enough that it clips some
    assert onet.triplet_loss(fa, fp, fn, att, 0.0, True) <= want


def test_adam_step_agrees_with_torch_adam_at_epsilon_hat():
    """tf.train.AdamOptimizer applies lr_t * m / (sqrt(v) + eps) with lr_t = lr sqrt(1-b2^t)/(1-b1^t) (feat3dnet.py:359-375);
    Kingma & Ba's form, which torch.optim.Adam implements, is the same update with eps_hat = eps / sqrt(1-b2^t) -- the
    "epsilon hat" remark of the TF documentation.  Five steps with that per-step epsilon must coincide."""
    g = torch.Generator().manual_seed(3)
    theta0 = torch.randn(50, generator=g, dtype=torch.float64)
    P = {"w": theta0.clone()}
    state = {}
    q = torch.nn.Parameter(theta0.clone())
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    opt = torch.optim.Adam([q], lr=lr, betas=(b1, b2), eps=eps)
    for t in range(1, 6):
        grad = torch.randn(50, generator=g, dtype=torch.float64) * (10.0 ** (-t))  # shrinking gradients make eps matter
        onet.adam_step(P, {"w": grad}, state, lr, b1, b2, eps)
        opt.param_groups[0]["eps"] = eps / math.sqrt(1 - b2 ** t)
        q.grad = grad.clone()
        opt.step()
        assert torch.allclose(P["w"], q.detach(), rtol=1e-12, atol=1e-15), t
    assert not torch.allclose(P["w"], theta0)


# ------------------------------------------------------------------------- the reference's own graph code
# tests/golden/ref_net.npz: outputs of the reference's models/{layers,pointnet_common,feat3dnet}.py, imported and executed
# UNMODIFIED on an eager float64 stand-in for the TensorFlow primitives they call (tests/golden/tf_shim.py,
# make_golden_net.py).  The wiring of the graph -- layer order, scopes / variable names, BN and activation flags, pooling,
# tile + concat, rotation, heads, loss -- is the reference's; oracle/net.py must reproduce it.
import json
import os
import subprocess
import sys

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _ref_net_cases(path=None):
    g = np.load(path or os.path.join(GOLD, "ref_net.npz"))
    for name in sorted({k.split("/")[0] for k in g.files if k.endswith("/config")}):  # the model cases ("pointnet_common/..." apart)
        cfg = json.loads(str(g[name + "/config"]))
        out = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "/out/")}
        clouds = np.load(os.path.join(GOLD, cfg["fixture"])).astype(np.float32)[None] if cfg.get("fixture") else g[name + "/clouds"]
        yield name, cfg, clouds, (g[name + "/keypoints"] if name + "/keypoints" in g.files else None), out


def _angle_diff(a, b):
    d = torch.as_tensor(a) - torch.as_tensor(b)
    return torch.atan2(torch.sin(d), torch.cos(d)).abs().max().item()


def test_oracle_net_reproduces_the_reference_graph_golden():
    names = []
    for name, cfg, clouds, kp, want in _ref_net_cases():
        names.append(name)
        P = onet.to_torch(onet.init_params(seed=cfg["seed"], feature_dim=cfg["feature_dim"], randomize_bn=cfg.get("randomize_bn", True)), torch.float64)
        stats = {}
        got = onet.inference_model(clouds, P, cfg["num_clusters"], 2.0, cfg["num_samples"], cfg["feature_dim"], cfg["no_regress"],
                                   cfg["training"], kp, stats, torch.float64)
        assert np.array_equal(np.asarray(got["xyz"], np.float64), want["xyz"]), name          # keypoints: FPS / all points / fed
        assert torch.allclose(got["features"], torch.as_tensor(want["features"]), rtol=1e-9, atol=1e-11), name
        assert torch.allclose(got["attention"], torch.as_tensor(want["attention_end_point"]), rtol=1e-9, atol=1e-12), name
        assert _angle_diff(got["orientation"], want["orientation"]) < 1e-9, name
        if "grouped_xyz" in want:  # the descriptor's clusters before and after the rotation by the detector's orientation
            desc = onet.descriptor(clouds, P, got["xyz"], got["orientation"], 2.0, cfg["num_samples"], cfg["feature_dim"],
                                   dtype=torch.float64)
            plain = onet.descriptor(clouds, P, got["xyz"], None, 2.0, cfg["num_samples"], cfg["feature_dim"], dtype=torch.float64)
            assert torch.allclose(desc["rotated_xyz"], torch.as_tensor(want["grouped_xyz"]), rtol=1e-9, atol=1e-12)
            assert torch.allclose(plain["rotated_xyz"], torch.as_tensor(want["grouped_xyz_before"]), rtol=1e-12, atol=1e-14)
        if cfg["training"]:
            fa, fp, fn = torch.chunk(got["features"], 3, dim=0)
            att_a = torch.chunk(got["attention"], 3, dim=0)[0]
            loss = onet.triplet_loss(fa, fp, fn, att_a, cfg["margin"], cfg["attention"])
            assert want["loss"] > 0 and abs(float(loss) - float(want["loss"])) < 1e-10, name
            updates = {k[len("bn_update/"):]: v for k, v in want.items() if k.startswith("bn_update/")}
            assert set(updates) == set(stats) and len(updates) == 18, name
            for k, v in updates.items():
                assert torch.allclose(stats[k], torch.as_tensor(v), rtol=1e-9, atol=1e-12), (name, k)
        else:
            assert stats == {}
    assert names == sorted(["c1_oxford_270_bn", "c1_oxford_270_init", "eval_fps", "eval_noregress_f128", "eval_all_points", "eval_keypoints_fed", "train_triplets",
                            "train_no_attention", "train_stage1"])


def test_oracle_gradients_reproduce_the_reference_graph_golden():
    """d loss / d variable of the reference's own graph + get_loss (autograd through the stand-in's primitives) against the
    gradients oracle.net.train_step differentiates -- the ones the CUDA backward is checked against (test_train_gpu.py)."""
    seen = 0
    for name, cfg, clouds, kp, want in _ref_net_cases():
        if not cfg["training"]:
            continue
        P = onet.to_torch(onet.init_params(seed=cfg["seed"], feature_dim=cfg["feature_dim"], randomize_bn=cfg.get("randomize_bn", True)), torch.float64,
                          requires_grad=True)
        a, p, n = np.split(clouds, 3, axis=0)
        loss, grads, _ = onet.train_step(a, p, n, P, {}, cfg["num_clusters"], 2.0, cfg["num_samples"], cfg["feature_dim"],
                                         cfg["margin"], cfg["attention"], cfg["no_regress"], lr=0.0, dtype=torch.float64)
        assert abs(float(loss) - float(want["loss"])) < 1e-10
        sums = {k[len("gradsum/"):]: v for k, v in want.items() if k.startswith("gradsum/")}
        assert set(sums) == set(grads) and len(sums) == 40, name
        nonzero = 0
        for k, (norm, proj) in sums.items():
            g = grads[k].detach().numpy().ravel()
            r = np.random.default_rng(len(k)).standard_normal(g.size)
            scale = max(norm, 1e-12)
            assert abs(np.sqrt((g * g).sum()) - norm) <= 1e-8 * scale + 1e-14, (name, k)
            assert abs(float(g @ r) - proj) <= 1e-7 * scale * np.sqrt(g.size) + 1e-14, (name, k)
            nonzero += norm > 1e-9
            if "grad/" + k in want:
                full = want["grad/" + k].ravel().astype(np.float64)
                assert np.abs(g - full).max() <= 1e-6 * max(np.abs(full).max(), 1e-12) + 1e-12, (name, k)
        # conv biases in front of a training-mode BN have (numerically) zero gradient, and so has the attention head when the
        # loss does not use it; every other variable must receive one
        if cfg["no_regress"] and not cfg["attention"]:  # train.sh's pretraining stage: the detector is not trained at all
            assert all(v[0] == 0.0 for k, v in sums.items() if k.startswith("detection/")) and nonzero >= 11, (name, nonzero)
        else:
            assert nonzero >= (30 if cfg["attention"] else 27), (name, nonzero)
        if not cfg["attention"]:
            assert sums["detection/attention/conv2d/weights"][0] == 0.0  # Attention=False: the detector gets no gradient
        seen += 1
    assert seen == 3


def test_reference_graph_golden_has_the_interesting_cases():
    cases = {c[0]: c for c in _ref_net_cases()}
    kp_case = cases["eval_keypoints_fed"]
    P = onet.init_params(seed=kp_case[1]["seed"], randomize_bn=True)
    det = onet.detector(kp_case[2], onet.to_torch(P), -1, 2.0, 64, keypoints_np=kp_case[3])
    assert (det["pts_cnt"] == 0).any() and ((det["pts_cnt"] > 0) & (det["pts_cnt"] < 64)).any()  # empty balls (fallback), padded ones
    fps_case = cases["eval_fps"]
    full = onet.detector(fps_case[2], onet.to_torch(onet.init_params(seed=fps_case[1]["seed"], randomize_bn=True)), 32, 2.0, 64)
    assert (full["pts_cnt"] == 64).any() and (full["pts_cnt"] < 64).any()                        # truncated at nsample, and padded
    assert cases["eval_noregress_f128"][4]["features"].shape[-1] == 128       # the 256-channel conv_mid_0 variant
    assert cases["eval_all_points"][4]["xyz"].shape[1] == cases["eval_all_points"][2].shape[1]  # num_clusters = -1
    assert not np.allclose(cases["eval_fps"][4]["grouped_xyz"], cases["eval_fps"][4]["grouped_xyz_before"])


@pytest.mark.skipif(not os.path.exists("/root/reference/models/feat3dnet.py"), reason="the reference tree is only in the build container")
def test_reference_graph_golden_regenerates_from_the_reference(tmp_path):
    """Re-runs the reference's model files (fresh interpreter) and compares with the committed file."""
    out = str(tmp_path / "ref_net.npz")
    r = subprocess.run([sys.executable, os.path.join(GOLD, "make_golden_net.py"), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    a, b = np.load(out), np.load(os.path.join(GOLD, "ref_net.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        if a[k].dtype.kind == "f":
            assert np.allclose(a[k], b[k], rtol=1e-12, atol=1e-14), k
        else:
            assert np.array_equal(a[k], b[k]), k

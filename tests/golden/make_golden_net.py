"""Regenerates tests/golden/ref_net.npz by EXECUTING THE REFERENCE'S OWN MODEL FILES (models/layers.py,
models/pointnet_common.py, models/feat3dnet.py, imported unmodified from /root/reference) on the eager TensorFlow stand-in of
tests/golden/tf_shim.py, in float64 (build container only; needs a fresh interpreter).

Per case: Feat3dNet(param).get_inference_model(cloud, is_training, use_bn=True, compute_det_gradients=False) -> keypoints,
descriptors, attention, orientation (+ the BN shadow updates in training mode), and for the training cases the reference's
get_loss on the anchor / positive / negative thirds and its gradient with respect to every trainable variable.  get_train_model itself cannot run: it leaves compute_det_gradients at
True, and feature_detection_module then indexes end_points['gradients'] of an empty dict (feat3dnet.py:104,122).
"keypoints fed" cases reproduce `feed_dict={end_points['keypoints']: ...}` (inference.py:128-131, train.py:291-298): feeding
that tensor replaces the output of sample_points for everything downstream, so models.feat3dnet.sample_points is bound to a
function returning the fed keypoints.

Variables come from oracle.net.init_params(seed, feature_dim, randomize_bn=True) under their TF names (1x1 kernels stored as
(Cin,Cout)); only inputs, configuration and outputs are written.

    python tests/golden/make_golden_net.py [out.npz]
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

CASES = {
    # name: dict(seed, B, N, param..., mode)
    "eval_fps": dict(seed=11, clouds=2, points=6000, num_clusters=32, num_samples=64, feature_dim=32, no_regress=False,
                     attention=True, training=False, keypoints=None, store_grouped=True),
    "eval_noregress_f128": dict(seed=12, clouds=1, points=2500, num_clusters=40, num_samples=32, feature_dim=128, no_regress=True,
                                attention=True, training=False, keypoints=None),
    "eval_all_points": dict(seed=13, clouds=1, points=600, num_clusters=-1, num_samples=64, feature_dim=32, no_regress=False,
                            attention=True, training=False, keypoints=None),
    "eval_keypoints_fed": dict(seed=14, clouds=2, points=2000, num_clusters=-1, num_samples=64, feature_dim=32, no_regress=False,
                               attention=True, training=False, keypoints=37),
    "train_triplets": dict(seed=15, clouds=6, points=1500, num_clusters=32, num_samples=64, feature_dim=32, no_regress=False,
                           attention=True, training=True, keypoints=None, margin=0.2, store_gradients=True),
    "train_no_attention": dict(seed=16, clouds=3, points=1200, num_clusters=24, num_samples=64, feature_dim=32, no_regress=False,
                               attention=False, training=True, keypoints=None, margin=0.5),
    # BASELINE.json configs[0]: example_data/oxford_270.bin (16384 points), 512 clusters, r = 2.0, nsample = 64, random init -- with
    # non-trivial BN statistics, and in the TF random-init state (EMA shadows 0); the cloud is the committed fixture, the variables
    # are those of tests/test_model_gpu.py::test_c1_oxford_270_fp32, which holds the CUDA forward to the oracle on this very input
    "c1_oxford_270_bn": dict(seed=0, fixture="oxford_270_xyz.npy", clouds=1, points=16384, num_clusters=512, num_samples=64,
                             feature_dim=32, no_regress=False, attention=True, training=False, keypoints=None, randomize_bn=True),
    "c1_oxford_270_init": dict(seed=0, fixture="oxford_270_xyz.npy", clouds=1, points=16384, num_clusters=512, num_samples=64,
                               feature_dim=32, no_regress=False, attention=True, training=False, keypoints=None, randomize_bn=False),
    # train.sh's pretraining stage: --noattention --noregress
    "train_stage1": dict(seed=17, clouds=3, points=1200, num_clusters=24, num_samples=64, feature_dim=32, no_regress=True,
                         attention=False, training=True, keypoints=None, margin=0.2),
}


def case_inputs(cfg):
    """Oxford-shape clouds (6 float32 columns like the .bin files: the model must ignore columns 3..5) and, if asked, keypoints
    that are partly cloud points and partly off-cloud positions (some with empty balls: the fallback rule of the ball query)."""
    if cfg.get("fixture"):
        return np.load(os.path.join(HERE, cfg["fixture"])).astype(np.float32)[None], None
    rng = np.random.default_rng(cfg["seed"])
    base = np.load(os.path.join(HERE, "oxford_270_xyz.npy")).astype(np.float32)
    clouds = []
    for _ in range(cfg["clouds"]):
        pts = base[rng.permutation(base.shape[0])[:cfg["points"]]]
        a = rng.uniform(0, 2 * np.pi)
        rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
        pts = (pts @ rot + rng.normal(0, 0.01, pts.shape)).astype(np.float32)
        clouds.append(np.concatenate([pts, rng.integers(0, 4, pts.shape).astype(np.float32)], axis=1))
    clouds = np.stack(clouds)
    kp = None
    if cfg["keypoints"]:
        k = cfg["keypoints"]
        kp = clouds[:, :k, :3].copy()
        kp[:, k // 2:] += rng.normal(0, 1.0, kp[:, k // 2:].shape).astype(np.float32)
        kp[:, -3:] += np.float32(500.0)  # far from every point: empty balls
    return clouds, kp


def run_reference(cfg, clouds, kp, params):
    import tf_shim
    import models.feat3dnet as ref_f3  # the reference's file (tf_shim.install put /root/reference first on sys.path)

    assert ref_f3.__file__.startswith(tf_shim_root), ref_f3.__file__
    store = tf_shim.STORE
    store.params, store.updates, store.touched, store.scope = dict(params), {}, set(), []
    leaves = {}
    if cfg["training"]:  # trainable variables as autograd leaves: the loss of the reference's graph is differentiated below
        import torch
        for k, v in params.items():
            if not k.endswith(("moving_mean", "moving_variance")):
                leaves[k] = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        store.params.update(leaves)
    model = ref_f3.Feat3dNet(dict(NoRegress=cfg["no_regress"], BaseScale=2.0, Attention=cfg["attention"],
                                  num_clusters=cfg["num_clusters"], num_samples=cfg["num_samples"],
                                  feature_dim=cfg["feature_dim"], margin=cfg.get("margin", 0.2), freeze_scopes=None))
    original = ref_f3.sample_points
    if kp is not None:
        ref_f3.sample_points = lambda xyz, npoint: tf_shim.t(kp)  # feed_dict={end_points['keypoints']: kp}
    try:
        xyz, features, attention, ep = model.get_inference_model(tf_shim.t(clouds), cfg["training"], use_bn=True,
                                                                 compute_det_gradients=False)
    finally:
        ref_f3.sample_points = original
    out = dict(xyz=xyz.detach().numpy(), features=features.detach().numpy(), attention_end_point=ep["attention"].detach().numpy(),
               orientation=ep["orientation"].detach().numpy())
    if cfg.get("store_grouped"):  # the descriptor's clusters before / after the rotation by the detector's orientation
        out.update(grouped_xyz_before=ep["grouped_xyz_before"].numpy(), grouped_xyz=ep["grouped_xyz"].numpy())
    assert (attention is None) == (not cfg["attention"])
    if cfg["training"]:
        import tensorflow as tf
        loss, ep2 = model.get_loss(None, tf.split(features, 3, axis=0),
                                   tf.split(attention, 3, axis=0)[0] if attention is not None else None, {})
        out["loss"] = np.float64(loss.detach().numpy())
        out["sum_positive"], out["sum_negative"] = ep2["sum_positive"].detach().numpy(), ep2["sum_negative"].detach().numpy()
        for k, v in store.updates.items():
            out["bn_update/" + k] = v
        # d loss / d variable through the reference's graph (autograd over the stand-in's primitives; tf.reduce_max and
        # torch.amax both share the gradient equally among tied maxima)
        names = sorted(leaves)
        grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
        for k, g in zip(names, grads):
            g = np.zeros_like(params[k], dtype=np.float64) if g is None else g.numpy()
            # float64: norm and a fixed random projection per variable; the full tensor in float32 for one case (file size)
            proj = np.random.default_rng(len(k)).standard_normal(g.size)
            out["gradsum/" + k] = np.array([np.sqrt((g * g).sum()), float(g.ravel() @ proj)])
            if cfg.get("store_gradients"):
                out["grad/" + k] = g.astype(np.float32)
    unused = set(params) - store.touched
    assert not unused, "variables the reference graph never read: %s" % sorted(unused)
    return out


# ---------------------------------------------------------------------------- models/pointnet_common.py on its own
def composition_inputs(seed=23):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(-3, 3, (2, 500, 3)).astype(np.float32)
    points = rng.normal(size=(2, 500, 5)).astype(np.float32)
    keypoints = xyz[:, :20].copy()
    keypoints[:, 10:] += rng.normal(0, 0.5, (2, 10, 3)).astype(np.float32)
    keypoints[:, -2:] += np.float32(100.0)  # empty balls
    orientations = rng.uniform(-np.pi, np.pi, (2, 20)).astype(np.float32)
    return xyz, points, keypoints, orientations


def composition_cases():
    """(function name, keyword arguments) for every branch of the four functions: features or not, use_xyz, kNN, radius
    normalisation, fed keypoints vs FPS, and the two (opposite) rotation conventions."""
    cases = [("sample_points", dict(npoint=16)), ("sample_points", dict(npoint=0))]
    for with_points in (False, True):
        for use_xyz in (True, False):
            for knn in (False, True):
                for normalize in (True, False):
                    for orient in (False, True):
                        kw = dict(with_points=with_points, use_xyz=use_xyz, knn=knn, normalize_radius=normalize, orient=orient)
                        cases.append(("query_and_group_points", dict(kw)))
                        cases.append(("sample_and_group", dict(kw, fed_keypoints=orient or knn)))
            cases.append(("sample_and_group_all", dict(with_points=with_points, use_xyz=use_xyz)))
    return cases


def digest(a, seed):
    """shape (padded to 5) followed by four fixed random projections of the flattened tensor, float64"""
    a = np.asarray(a, np.float64)
    proj = np.random.default_rng(1000 + seed).standard_normal((4, a.size)) @ a.ravel()
    return np.concatenate([np.array(a.shape + (0,) * (5 - a.ndim), np.float64), proj])


def call_composition(mod, wrap, fn, kw, xyz, points, keypoints, orientations, nsample=16, radius=1.5, npoint=20):
    """Calls `fn` of a pointnet_common module (the reference's on the stand-in, or the product's) and returns its tensors in a
    flat, ordered list.  `wrap` turns a NumPy array into the module's tensor type."""
    X, P, K, O = wrap(xyz), wrap(points), wrap(keypoints), wrap(orientations)
    if fn == "sample_points":
        return [mod.sample_points(X, kw["npoint"])]
    pts = P if kw["with_points"] else None
    if fn == "sample_and_group_all":
        return list(mod.sample_and_group_all(X, pts, kw["use_xyz"]))
    ori = O if kw["orient"] else None
    if fn == "query_and_group_points":
        return list(mod.query_and_group_points(X, pts, K, nsample, radius, knn=kw["knn"], use_xyz=kw["use_xyz"],
                                               normalize_radius=kw["normalize_radius"], orientations=ori))
    new_xyz, new_points, idx, grouped, ep = mod.sample_and_group(
        npoint, radius, nsample, X, pts, knn=kw["knn"], use_xyz=kw["use_xyz"], keypoints=K if kw["fed_keypoints"] else None,
        orientations=ori, normalize_radius=kw["normalize_radius"])
    out = [new_xyz, new_points, idx, grouped, ep["grouped_xyz_before"], ep["grouped_xyz"]]
    return out + ([ep["rotation"]] if kw["orient"] else [])


if __name__ == "__main__":
    import tf_shim
    from oracle import net as onet

    tf_shim_root = os.environ.get("F3D_REFERENCE", "/root/reference")
    tf_shim.install(tf_shim_root)
    store = {}
    for name, cfg in CASES.items():
        clouds, kp = case_inputs(cfg)
        params = onet.init_params(seed=cfg["seed"], feature_dim=cfg["feature_dim"], randomize_bn=cfg.get("randomize_bn", True))
        out = run_reference(cfg, clouds, kp, params)
        store[name + "/config"] = np.array(json.dumps(cfg))
        if not cfg.get("fixture"):  # fixture clouds are files of their own
            store[name + "/clouds"] = clouds
        if kp is not None:
            store[name + "/keypoints"] = kp
        for k, v in out.items():
            store[name + "/out/" + k] = v
        print(name, {k: (v.shape if hasattr(v, "shape") and v.shape else float(v)) for k, v in out.items() if "/" not in k},
              "bn updates:", sum(k.startswith("bn_update") for k in out), "gradients:", sum(k.startswith("grad") for k in out))
    import models.pointnet_common as ref_pc  # the reference's file
    assert ref_pc.__file__.startswith(tf_shim_root), ref_pc.__file__
    inputs = composition_inputs()
    for i, (fn, kw) in enumerate(composition_cases()):
        outs = call_composition(ref_pc, tf_shim.t, fn, kw, *inputs)
        full = fn == "sample_points" or (kw["with_points"] and kw["use_xyz"] and (fn == "sample_and_group_all" or (
            not kw["knn"] and kw["normalize_radius"] and kw["orient"])))
        for j, o in enumerate(outs):
            a = o.detach().numpy()
            store["pointnet_common/%03d/%d/digest" % (i, j)] = digest(a, i * 16 + j)
            if full:  # whole tensors for one call per function; shape + projections for every branch (file size)
                store["pointnet_common/%03d/%d/full" % (i, j)] = a.astype(np.int32) if a.dtype.kind == "i" else a.astype(np.float32)
    print("pointnet_common:", len(composition_cases()), "calls")
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "ref_net.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")

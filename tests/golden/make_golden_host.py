"""Regenerates tests/golden/ref_host.npz from the REFERENCE's own host-side Python (needs /root/reference; run in the build
container, not on the GPU box):

  validate / load_validation_groundtruths   train.py:243-315.  train.py cannot be imported (TensorFlow, argparse at module
      level), so the two FunctionDefs are taken out of the parsed file with `ast` and executed unmodified in a namespace
      holding what they read (np, os, NUM_CLUSTERS, the reference's DataGenerator); the TF session is replaced by an object
      whose run(fetches, feed_dict) returns descriptors from `toy_descriptors` below -- the cluster stacking, keypoint
      offsets, distance, percentile and FP-rate arithmetic are the reference's.
  data/augment.py                           imported as it is (NumPy only); its random draws are replaced by fixed values so
      that the geometric statement of every augmentation (axis, sign, composition order, clipping) is recorded.

  compute_descriptors / nms                 inference.py:67-180,226-261, taken out with `ast`; the network factory and the TF
      session are stand-ins evaluating `toy_network`, so the file loop, --num_points, the MAX_POINTS chunking of the attention
      pass, NMS, --use_keypoints_from and the rows written are the reference's.
  tf_ops/grouping/tf_grouping.py::knn_point  taken out with `ast`, executed in float32 on tests/golden/tf_shim.py.

Only inputs, draws and outputs are stored; no reference source is written anywhere.

    python tests/golden/make_golden_host.py
"""
import ast
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("F3D_REFERENCE", "/root/reference")


def reference_module(relpath, name):
    """Import a NumPy-only reference file by path (data/augment.py, data/datagenerator.py)."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_functions(relpath, names, namespace):
    """The named top-level functions of a reference file that cannot be imported, compiled unmodified into `namespace`."""
    path = os.path.join(REFERENCE, relpath)
    tree = ast.parse(open(path).read(), filename=path)
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert sorted(n.name for n in body) == sorted(names), "missing in %s: %s" % (relpath, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), namespace)
    return [namespace[n] for n in names]


# ------------------------------------------------------------------------------------------- validation
def toy_descriptors(cloud, keypoints, dim=8):
    """A deterministic stand-in for the network: descriptor j = moments of the points within 50 m (in x) of keypoint j.
    cloud (1,N,>=3), keypoints (1,M,3) of any numeric dtype -> (1,M,3), (1,M,dim) float32."""
    cloud = np.asarray(cloud, np.float64)[0, :, :3]
    kp = np.asarray(keypoints, np.float64)[0]
    feats = np.zeros((kp.shape[0], dim), np.float64)
    for j in range(kp.shape[0]):
        local = cloud[np.abs(cloud[:, 0] - kp[j, 0]) < 50.0] - kp[j]
        if local.shape[0]:
            feats[j, :3], feats[j, 3:6] = local.mean(0), local.std(0)
            feats[j, 6], feats[j, 7] = np.abs(local).max(), local.shape[0] / 64.0
    return kp[None].astype(np.float32), feats[None].astype(np.float32)


def write_validation_set(folder, num_pairs=600, seed=31):
    """`<i>_0.bin` / `<i>_1.bin` cluster pairs (6 float32 columns) + groundtruths.txt; matching pairs are noisy copies,
    non-matching ones independent draws of the same spread, so that the FP rate at 95 % recall is neither 0 nor 1."""
    rng = np.random.default_rng(seed)
    lines = ["idx dist overlap match"]
    for i in range(num_pairs):
        match = int(rng.random() < 0.5)
        scale = 1.0 + (i % 5)
        a = rng.normal(0, scale, (int(rng.integers(20, 40)), 6)).astype(np.float32)
        b = (a + rng.normal(0, 0.25 * scale, a.shape) if match else rng.normal(0, scale, a.shape)).astype(np.float32)
        a.tofile(os.path.join(folder, "%d_0.bin" % i))
        b.tofile(os.path.join(folder, "%d_1.bin" % i))
        lines.append("%d %.3f %.3f %d" % (i, rng.random(), rng.random(), match))
    path = os.path.join(folder, "groundtruths.txt")
    open(path, "w").write("\n".join(lines) + "\n")
    return path


class ToySession(object):
    """Stands in for tf.Session in the reference's validate(): run(fetches, feed_dict) -> toy_descriptors(cloud, keypoints)."""

    def __init__(self, end_points):
        self.ep = end_points

    def run(self, fetches, feed_dict):
        assert fetches == [self.ep['output_xyz'], self.ep['output_features']]
        return toy_descriptors(feed_dict[self.ep['input_pointclouds']], feed_dict[self.ep['keypoints']])


def reference_validate(folder, gt_path, proportion=1):
    dg = reference_module("data/datagenerator.py", "ref_datagenerator")
    ns = dict(np=np, os=os, NUM_CLUSTERS=512, DataGenerator=dg.DataGenerator)  # NUM_CLUSTERS: train.py:24
    validate, load_gt = reference_functions("train.py", ["validate", "load_validation_groundtruths"], ns)
    gts = load_gt(gt_path, proportion)
    ep = dict(output_xyz="xyz", output_features="features", input_pointclouds="clouds", keypoints="keypoints")
    return gts, validate(ToySession(ep), ep, "is_training", folder, gts, 6)


# ------------------------------------------------------------------------------------------- augmentations
AUG_DRAWS = dict(angle01=0.3125, jitter_scale=3.0, scale=1.125, small_angles=(0.5, -2.0, 4.0), shift=(0.05, -0.1, 0.025))


def reference_augmentations(cloud):
    """Every reference augmentation applied to `cloud` (N,3) with its random draws replaced by AUG_DRAWS."""
    aug = reference_module("data/augment.py", "ref_augment")
    noise = np.random.default_rng(5).standard_normal(cloud.shape)
    out = {}

    class Draws(object):  # what the reference reads from np.random, returning the fixed draws
        def __init__(self, uniform=None, randn=None):
            self._u, self._n = uniform, randn

        def uniform(self, *a, **k):
            return self._u

        def randn(self, *shape):
            return np.asarray(self._n, np.float64).reshape(shape)

    real = aug.np.random
    try:
        aug.np.random = Draws(uniform=AUG_DRAWS["angle01"])
        out["RotateZ"], out["RotateY"] = aug.RotateZ().apply(cloud.copy()), aug.RotateY().apply(cloud.copy())
        aug.np.random = Draws(randn=noise * AUG_DRAWS["jitter_scale"])
        out["Jitter"] = aug.Jitter().apply(cloud.copy())
        aug.np.random = Draws(uniform=AUG_DRAWS["scale"])
        out["Scale"] = aug.Scale().apply(cloud.copy())
        aug.np.random = Draws(randn=AUG_DRAWS["small_angles"])
        out["RotateSmall"] = aug.RotateSmall().apply(cloud.copy())
        aug.np.random = Draws(uniform=np.asarray(AUG_DRAWS["shift"]))
        out["Shift"] = aug.Shift().apply(cloud.copy())
    finally:
        aug.np.random = real
    return noise, out


# ------------------------------------------------------------------------------------------- initialize_model
def reference_restore_plan(model_vars, checkpoint_vars, ignore_missing_vars, restore_exclude):
    """initialize_model of inference.py:183-217 (same text in train.py:187-232) run unmodified over stand-in TF objects:
    returns ("restore", [names handed to tf.train.Saver]) or ("error", [names the Saver would not find in the checkpoint]).
    Restated TF behaviour: get_collection(GLOBAL_VARIABLES, scope) filters with re.match(scope, name); Saver.restore raises
    NotFoundError when a variable of its var_list is absent from the checkpoint."""
    import logging
    import re
    import types

    class Var(object):
        def __init__(self, name):
            self.op = types.SimpleNamespace(name=name)

        def __repr__(self):
            return self.op.name

    variables = [Var(n) for n in model_vars]
    result = {}

    class Saver(object):
        def __init__(self, var_list):
            self.names = [v.op.name for v in var_list]

        def restore(self, sess, checkpoint):
            absent = [n for n in self.names if n not in checkpoint_vars]
            result["plan"] = ("error", absent) if absent else ("restore", self.names)

    tf = types.SimpleNamespace(
        global_variables_initializer=lambda: "init", GraphKeys=types.SimpleNamespace(GLOBAL_VARIABLES="global_variables"),
        get_collection=lambda key, scope=None: [v for v in variables if scope is None or re.match(scope, v.op.name)],
        train=types.SimpleNamespace(Saver=Saver))
    ns = dict(tf=tf, logger=logging.getLogger("reference.inference"), args=types.SimpleNamespace(checkpoint="ckpt"),
              get_tensors_in_checkpoint_file=lambda ckpt: list(checkpoint_vars), print=lambda *a, **k: None)
    (initialize_model,) = reference_functions("inference.py", ["initialize_model"], ns)
    initialize_model(types.SimpleNamespace(run=lambda op: None), "ckpt", ignore_missing_vars, restore_exclude)
    return result["plan"]


# ------------------------------------------------------------------------------------------- knn_point
def reference_knn_point(k, xyz1, xyz2):
    """knn_point of tf_ops/grouping/tf_grouping.py:63-88 (the file loads the op library at import, so the FunctionDef is taken
    out with `ast`), executed in float32 on the TF stand-in of tf_shim.py; its select_top_k is the C oracle's selection sort,
    which is pinned to the reference kernel.  What is the reference's: the (b,m,n,c) tiling, the squared distance, the slices."""
    import contextlib
    import io

    import torch
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import tf_shim
    from oracle import ops

    tf, _ = tf_shim.build_tensorflow()
    tf.slice = lambda x, begin, size: x[tuple(slice(b, None if n == -1 else b + n) for b, n in zip(begin, size))]

    def select_top_k(kk, dist):
        outi, out = ops.select_top_k(kk, dist.numpy())
        return tf_shim.t(outi, torch.int32), tf_shim.t(out, torch.float32)

    (knn_point,) = reference_functions("tf_ops/grouping/tf_grouping.py", ["knn_point"], dict(tf=tf, select_top_k=select_top_k))
    with contextlib.redirect_stdout(io.StringIO()):  # the function prints its shapes
        val, idx = knn_point(k, tf_shim.t(xyz1, torch.float32), tf_shim.t(xyz2, torch.float32))
    return val.numpy(), idx.numpy()


def knn_inputs(seed=21):
    rng = np.random.default_rng(seed)
    xyz1 = rng.uniform(-5, 5, (2, 400, 3)).astype(np.float32)
    xyz2 = rng.uniform(-5, 5, (2, 48, 3)).astype(np.float32)
    xyz1[0, 100:110] = xyz1[0, 0:10]  # duplicated points: equal distances, the selection sort's (unstable) tie order shows
    xyz2[1, :8] = xyz1[1, :8]         # queries that are cloud points: zero distances
    return xyz1, xyz2


# ------------------------------------------------------------------------------------------- compute_descriptors
def toy_network(cloud, keypoints, dim=8):
    """Stand-in for the network in the file flow: attention_j = local density of the cloud around keypoint j (distinct values,
    permutation invariant), descriptor_j = a few moments.  cloud (1,N,>=3), keypoints (1,M,3) -> xyz (1,M,3) f32,
    features (1,M,dim) f32, attention (1,M) f32."""
    pts = np.asarray(cloud, np.float64)[0, :, :3]
    kp = np.asarray(keypoints, np.float64)[0]
    d2 = ((kp[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    w = np.exp(-d2 / 4.0)
    att = w.sum(1) / pts.shape[0] + 1e-3
    mean = (w[:, :, None] * (pts[None] - kp[:, None])).sum(1) / w.sum(1)[:, None]
    feats = np.concatenate([mean, att[:, None], np.sin(kp), np.full((kp.shape[0], dim - 7), pts.shape[0] / 1000.0)], axis=1)
    return kp[None].astype(np.float32), feats[None].astype(np.float32), att[None].astype(np.float32)


def write_inference_set(folder, seed=41):
    """data/<name>.bin clouds (6 columns) and kp/<name>_kp.bin keypoint files (3 columns)."""
    rng = np.random.default_rng(seed)
    data, kps = os.path.join(folder, "data"), os.path.join(folder, "kp")
    os.makedirs(data), os.makedirs(kps)
    for name, n in (("scan_a", 1500), ("scan_b", 900)):
        cloud = np.concatenate([rng.uniform(-6, 6, (n, 3)) * [1, 1, 0.2], rng.normal(size=(n, 3))], axis=1).astype(np.float32)
        cloud.tofile(os.path.join(data, name + ".bin"))
        (cloud[rng.permutation(n)[:23 if name == "scan_a" else 11], :3] + np.float32(0.25)).tofile(os.path.join(kps, name + "_kp.bin"))
    return data, kps


INFERENCE_CASES = {
    "detect": dict(num_points=-1, use_keypoints_from=False, max_keypoints=64, max_points=700),
    "detect_first_800": dict(num_points=800, use_keypoints_from=False, max_keypoints=1024, max_points=30000),
    "keypoints_from": dict(num_points=-1, use_keypoints_from=True, max_keypoints=1024, max_points=30000),
}


def reference_compute_descriptors(data_dir, kp_dir, output_dir, case):
    """The reference's compute_descriptors() and nms() (inference.py:67-180,226-261; FunctionDefs taken out with `ast`, the
    module itself parses the command line and imports TensorFlow) run unmodified: `args` carries the CLI fields (:26-58),
    the network factory returns a class whose ops are names, and tf.Session().run evaluates `toy_network` on the fed cloud and
    keypoints.  MAX_POINTS (inference.py:22, 30000) is lowered in one case so that the attention pass is chunked."""
    import logging
    import types
    from sklearn.neighbors import NearestNeighbors

    class Net(object):
        def __init__(self, param):
            assert param["num_clusters"] == -1 and param["Attention"] is True  # inference.py:81-83
            self.param = param

        def get_placeholders(self, data_dim):
            return "cloud_pl", None, None

        def get_inference_model(self, cloud_pl, is_training, use_bn=True):
            return "xyz_op", "features_op", "attention_op", {"keypoints": "keypoints_pl"}

    class Session(object):
        def __init__(self, config=None):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

        def run(self, fetches, feed_dict):
            assert feed_dict["is_training_pl"] is False
            xyz, feats, att = toy_network(feed_dict["cloud_pl"], feed_dict["keypoints_pl"])
            return [dict(xyz_op=xyz, features_op=feats, attention_op=att)[f] for f in fetches]

    dg = reference_module("data/datagenerator.py", "ref_datagenerator_inf")
    args = types.SimpleNamespace(
        gpu=0, model="3DFeatNet", data_dim=6, num_points=case["num_points"], base_scale=2.0, num_samples=64,
        use_keypoints_from=kp_dir if case["use_keypoints_from"] else None, feature_dim=32, randomize_points=False, nms_radius=0.5,
        min_response_ratio=1e-2, max_keypoints=case["max_keypoints"], data_dir=data_dir, checkpoint="unused", output_dir=output_dir)
    ns = dict(np=np, os=os, args=args, logger=logging.getLogger("reference.inference"), MAX_POINTS=case["max_points"],
              USE_BN=True, config=None, DataGenerator=dg.DataGenerator, NearestNeighbors=NearestNeighbors,
              get_network=lambda name: Net, initialize_model=lambda sess, ckpt: None, log_arguments=lambda: None,
              tf=types.SimpleNamespace(placeholder=lambda dtype: "is_training_pl", bool=bool, Session=Session))
    compute_descriptors, _ = reference_functions("inference.py", ["compute_descriptors", "nms"], ns)
    compute_descriptors()
    return {f: np.fromfile(os.path.join(output_dir, f), dtype=np.float32) for f in sorted(os.listdir(output_dir))}


if __name__ == "__main__":
    import tempfile

    store = {}
    with tempfile.TemporaryDirectory() as tmp:
        gt_path = write_validation_set(tmp)
        for prop in (1, 0.25):
            gts, fp = reference_validate(tmp, gt_path, prop)
            store["validate/fp_rate_%s" % prop] = np.float64(fp)
            store["validate/groundtruths_%s" % prop] = np.array(gts, np.int64)
            print("validate: proportion", prop, "pairs", len(gts), "fp rate", fp)
    cloud = np.random.default_rng(8).uniform(-20, 20, (64, 3))
    noise, outs = reference_augmentations(cloud)
    store["augment/cloud"], store["augment/noise"] = cloud, noise
    for k, v in outs.items():
        store["augment/" + k] = np.asarray(v, np.float64)
    with tempfile.TemporaryDirectory() as tmp:
        data_dir, kp_dir = write_inference_set(tmp)
        for cname, case in INFERENCE_CASES.items():
            for f, rows in reference_compute_descriptors(data_dir, kp_dir, os.path.join(tmp, "out_" + cname), case).items():
                store["inference/%s/%s" % (cname, f)] = rows
                print("compute_descriptors:", cname, f, rows.size // 11, "keypoints")
    xyz1, xyz2 = knn_inputs()
    val, idx = reference_knn_point(16, xyz1, xyz2)
    store["knn/xyz1"], store["knn/xyz2"], store["knn/val"], store["knn/idx"] = xyz1, xyz2, val, idx
    print("knn_point:", val.shape, idx.dtype)
    np.savez_compressed(os.path.join(HERE, "ref_host.npz"), **store)
    print("wrote ref_host.npz", os.path.getsize(os.path.join(HERE, "ref_host.npz")), "bytes")
    sys.exit(0)

"""CPU tests (-m "not gpu"): the host-side mirrors (validation.py, augment.py / data/augment.py, data/datagenerator.py) against
the REFERENCE's own Python.  tests/golden/ref_host.npz holds outputs of the reference's validate() / load_validation_groundtruths()
(train.py:243-315, taken out of the file with `ast`, the TF session replaced by a stand-in network) and of data/augment.py with
fixed random draws (tests/golden/make_golden_host.py); where /root/reference is present the same code is also run live."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests.conftest import pkg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
HAVE_REFERENCE = os.path.exists("/root/reference/train.py")
needs_reference = pytest.mark.skipif(not HAVE_REFERENCE, reason="the reference tree is only in the build container")


def _mk():
    spec = importlib.util.spec_from_file_location("make_golden_host", os.path.join(GOLD, "make_golden_host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ToyModel(object):
    """get_inference_model(cloud, False, keypoints=...) -> the stand-in descriptors the golden file was made with."""

    def __init__(self, mk):
        self.mk = mk

    def get_inference_model(self, pc, is_training, keypoints=None, fetch_features=True):
        assert is_training is False
        xyz, feats = self.mk.toy_descriptors(pc.numpy(), keypoints.numpy())
        return torch.as_tensor(xyz), torch.as_tensor(feats), None, {}


def test_validate_matches_reference_golden(tmp_path):
    """600 cluster pairs = one full pass of 512 stacked clusters + one of 88: stacking 100 m apart, zero keypoints beyond the
    last cluster, descriptor distance, 95th percentile, strict `<`, and the `proportion` subsampling of the ground truths."""
    mk, val = _mk(), pkg("validation")
    g = np.load(os.path.join(GOLD, "ref_host.npz"))
    gt_path = mk.write_validation_set(str(tmp_path))
    for prop in (1, 0.25):
        gts = val.load_validation_groundtruths(gt_path, prop)
        assert np.array_equal(np.array(gts, np.int64), g["validate/groundtruths_%s" % prop])
        fp = val.validate(ToyModel(mk), str(tmp_path), gts, 6, device="cpu")
        assert 0.0 < fp < 1.0 and fp == float(g["validate/fp_rate_%s" % prop]), prop


@needs_reference
def test_validate_matches_reference_live(tmp_path):
    mk, val = _mk(), pkg("validation")
    gt_path = mk.write_validation_set(str(tmp_path), num_pairs=70, seed=5)
    ref_gts, ref_fp = mk.reference_validate(str(tmp_path), gt_path)
    gts = val.load_validation_groundtruths(gt_path)
    assert gts == ref_gts
    assert val.validate(ToyModel(mk), str(tmp_path), gts, 6, device="cpu") == ref_fp
    assert mk.reference_validate(str(tmp_path), gt_path, 0)[0] == val.load_validation_groundtruths(gt_path, 0)


def _ours_with_draws(mk, cloud, noise):
    """3dfeatnet_b200/augment.py with its generator draws replaced by the draws the golden file was made with."""
    aug = pkg("augment")
    d = mk.AUG_DRAWS
    x = torch.as_tensor(cloud)[None]
    real = aug._rand, aug._randn
    out = {}
    try:
        aug._rand = lambda gen, shape, device: torch.full(shape, d["angle01"], dtype=torch.float64)
        out["RotateZ"], out["RotateY"] = aug.rotate_z(x), aug.rotate_y(x)
        aug._randn = lambda gen, shape, device: torch.as_tensor(noise * d["jitter_scale"]).reshape(shape)
        out["Jitter"] = aug.jitter(x)
        # scale = low + (high - low) u  ->  u that reproduces the recorded factor; same for the shift
        aug._rand = lambda gen, shape, device: torch.full(shape, (d["scale"] - 0.8) / (1.25 - 0.8), dtype=torch.float64)
        out["Scale"] = aug.scale(x)
        aug._randn = lambda gen, shape, device: torch.tensor(d["small_angles"], dtype=torch.float64).reshape(shape)
        out["RotateSmall"] = aug.rotate_small(x)
        aug._rand = lambda gen, shape, device: ((torch.tensor(d["shift"], dtype=torch.float64) / 0.1 + 1) / 2).reshape(shape)
        out["Shift"] = aug.shift(x)
    finally:
        aug._rand, aug._randn = real
    return {k: v[0].numpy() for k, v in out.items()}


def test_augmentations_match_reference_golden():
    """axis and sign of RotateZ / RotateY, Rz Ry Rx order and angle clipping of RotateSmall (angles 0.03, -0.12, clip(0.24) = 0.18),
    jitter clipping at +-0.05, per-cloud scale and shift: the reference's outputs for fixed draws."""
    mk = _mk()
    g = np.load(os.path.join(GOLD, "ref_host.npz"))
    ours = _ours_with_draws(mk, g["augment/cloud"], g["augment/noise"])
    for name in ("RotateZ", "RotateY", "Jitter", "Scale", "RotateSmall", "Shift"):
        assert np.allclose(ours[name], g["augment/" + name], rtol=1e-12, atol=1e-12), name
    j = g["augment/Jitter"] - g["augment/cloud"]
    assert np.isclose(np.abs(j).max(), 0.05) and (np.abs(j) < 0.05 - 1e-9).any()  # the fixture exercises the clip


@needs_reference
def test_augmentations_and_generator_match_reference_live(tmp_path):
    mk = _mk()
    cloud = np.random.default_rng(1).uniform(-5, 5, (40, 3))
    noise, want = mk.reference_augmentations(cloud)
    ours = _ours_with_draws(mk, cloud, noise)
    for name, w in want.items():
        assert np.allclose(ours[name], w, rtol=1e-12, atol=1e-12), name
    # get_augmentations_from_list: same objects in the same order for every subset / upright axis
    ref_aug, da = mk.reference_module("data/augment.py", "ref_augment_live"), pkg("data.augment")
    names = ['Jitter', 'RotateSmall', 'Shift', 'Rotate1D', 'Scale']
    for mask in range(32):
        subset = [n for i, n in enumerate(names) if mask >> i & 1]
        for axis in (0, 1, 2):
            assert ([type(a).__name__ for a in da.get_augmentations_from_list(subset, axis)] ==
                    [type(a).__name__ for a in ref_aug.get_augmentations_from_list(subset, axis)])
    assert da.get_augmentations_from_list(None) == ref_aug.get_augmentations_from_list(None) == []
    # DataGenerator: metadata parsing, file reader, crop rule and epoch bookkeeping against the reference class
    from tests.test_host_cpu import _write_dataset
    meta = _write_dataset(tmp_path, n_clouds=7, pts=250, seed=4)
    ref_dg = mk.reference_module("data/datagenerator.py", "ref_datagenerator_live").DataGenerator(meta, num_cols=6)
    our_dg = pkg("data.datagenerator").DataGenerator(meta, num_cols=6, seed=0)
    assert our_dg.paths_and_labels == ref_dg.paths_and_labels and our_dg.size == ref_dg.size
    assert list(our_dg.indices) == list(ref_dg.indices) and our_dg.dataset_folder == ref_dg.dataset_folder
    for i in range(our_dg.size):
        raw = ref_dg.get_point_cloud(i)
        assert np.array_equal(our_dg.get_point_cloud(i), raw)
        a, b = our_dg.process_point_cloud(raw, num_points=64), ref_dg.process_point_cloud(raw, num_points=64)
        assert a.shape == b.shape == (64, 6)
        inside = {r.tobytes() for r in raw[np.sum(np.square(raw[:, :3]), axis=1) <= 400.0]}
        assert {r.tobytes() for r in a} <= inside and {r.tobytes() for r in b} <= inside  # both sample the same cropped set
        pos, neg = our_dg.get_positive_negative(i)
        assert pos in ref_dg.paths_and_labels[i][1] and neg not in ref_dg.paths_and_labels[i][1] | ref_dg.paths_and_labels[i][2]
    # the reference's next_triplet / get_positive_negative call random.sample(<set>, 1), which Python >= 3.11 rejects, so the
    # triplet draw itself cannot be run here; its rules are the assertions above and test_datagenerator_follows_the_reference_interface
    with pytest.raises(TypeError):
        ref_dg.next_triplet(k=1, num_points=32)
    ref_dg.reset()
    our_dg.shuffle()
    our_dg.reset()
    assert list(our_dg.indices) == list(ref_dg.indices) == list(range(7))


def test_knn_point_matches_reference_golden():
    """the C oracle's knn_point (what the CUDA kNN is checked against) vs the reference's own Python knn_point: same float32
    squared distances, same k smallest in the selection sort's order (ties included)"""
    from oracle import ops
    g = np.load(os.path.join(GOLD, "ref_host.npz"))
    val, idx = ops.knn_point(16, g["knn/xyz1"], g["knn/xyz2"])
    assert np.array_equal(idx, g["knn/idx"]) and np.array_equal(val, g["knn/val"])
    assert (g["knn/val"][1, :8, 0] == 0).all() and (np.diff(g["knn/val"], axis=2) >= 0).all()
    assert (np.diff(g["knn/val"][0], axis=1) == 0).any()  # the fixture holds exact ties


@needs_reference
def test_knn_point_matches_reference_live():
    from oracle import ops
    mk = _mk()
    xyz1, xyz2 = mk.knn_inputs(seed=99)
    val, idx = mk.reference_knn_point(5, xyz1, xyz2)
    oval, oidx = ops.knn_point(5, xyz1, xyz2)
    assert np.array_equal(idx, oidx) and np.array_equal(val, oval)


class ToyFlowModel(object):
    """get_inference_model(cloud, False, keypoints=..., fetch_features=...) -> the stand-in network of the golden file"""

    def __init__(self, mk):
        self.mk = mk

    def get_inference_model(self, pc, is_training, keypoints=None, fetch_features=True):
        assert is_training is False
        xyz, feats, att = self.mk.toy_network(pc.numpy(), keypoints.numpy())
        att = torch.as_tensor(att)
        return torch.as_tensor(xyz), (torch.as_tensor(feats) if fetch_features else None), att, {"attention": att}


def _our_compute_descriptors(mk, monkeypatch, data_dir, kp_dir, out_dir, case):
    from oracle import nms as onms
    inf = pkg("inference")

    def cpu_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints):  # the device NMS refuses CPU tensors
        x, a, num, _ = onms.nms(xyz.numpy(), attention.numpy(), nms_radius, min_response_ratio, max_keypoints)
        return torch.as_tensor(x), torch.as_tensor(a), num

    monkeypatch.setattr(inf, "nms", cpu_nms)
    monkeypatch.setattr(inf, "MAX_POINTS", case["max_points"])
    done = inf.compute_descriptors(ToyFlowModel(mk), data_dir, out_dir, data_dim=6, num_points=case["num_points"],
                                   use_keypoints_from=kp_dir if case["use_keypoints_from"] else None,
                                   max_keypoints=case["max_keypoints"], device="cpu")
    return {f: np.fromfile(os.path.join(out_dir, f), dtype=np.float32) for f in done}


def test_compute_descriptors_matches_reference_golden(tmp_path, monkeypatch):
    """the reference's compute_descriptors() + nms() run unmodified around a stand-in network (make_golden_host.py): chunked
    attention pass over every point, NMS, keypoints kept / padded, --num_points, --use_keypoints_from, the rows written"""
    mk = _mk()
    g = np.load(os.path.join(GOLD, "ref_host.npz"))
    data_dir, kp_dir = mk.write_inference_set(str(tmp_path))
    for cname, case in mk.INFERENCE_CASES.items():
        ours = _our_compute_descriptors(mk, monkeypatch, data_dir, kp_dir, str(tmp_path / ("out_" + cname)), case)
        assert sorted(ours) == ["scan_a.bin", "scan_b.bin"]
        for f, rows in ours.items():
            want = g["inference/%s/%s" % (cname, f)]
            assert rows.size == want.size and rows.size % 11 == 0 and rows.size > 0, (cname, f)
            assert np.array_equal(rows, want), (cname, f)


@needs_reference
def test_compute_descriptors_matches_reference_live(tmp_path, monkeypatch):
    mk = _mk()
    data_dir, kp_dir = mk.write_inference_set(str(tmp_path), seed=77)
    case = dict(num_points=1000, use_keypoints_from=False, max_keypoints=200, max_points=333)
    want = mk.reference_compute_descriptors(data_dir, kp_dir, str(tmp_path / "ref_out"), case)
    ours = _our_compute_descriptors(mk, monkeypatch, data_dir, kp_dir, str(tmp_path / "our_out"), case)
    assert sorted(ours) == sorted(want)
    for f in want:
        assert np.array_equal(ours[f], want[f]), f


@needs_reference
def test_initialize_model_restore_rules_match_reference_live(tmp_path):
    """checkpoint.initialize_model vs the reference's initialize_model (inference.py:183-217) run unmodified over stand-in TF
    objects: which variables are restored, which are excluded (regex prefix match of the scope), and when a missing one is fatal"""
    import types
    from oracle import net as onet
    mk, ck = _mk(), importlib.import_module("3dfeatnet_b200.checkpoint")
    params = onet.init_params(seed=1, randomize_bn=True)
    model_vars = sorted(params)
    partial = {k: v for k, v in params.items() if not k.startswith("detection/conv_post_1") and "orientation" not in k}
    full_path, part_path = str(tmp_path / "full.npz"), str(tmp_path / "part.npz")
    ck.save_npz(params, full_path)
    ck.save_npz(partial, part_path)
    for path, ckpt_vars in ((full_path, set(params)), (part_path, set(partial))):
        for ignore in (False, True):
            for exclude in (None, ["detection"], ["description/layer1/conv_mid", "detection/att"], ["detection/conv_post_1", "detection/orient"]):
                kind, names = mk.reference_restore_plan(model_vars, ckpt_vars, ignore, exclude)
                model = types.SimpleNamespace(weights={k: torch.zeros(v.shape) for k, v in params.items()}, invalidate=lambda: None)
                if kind == "error":
                    with pytest.raises(KeyError):
                        ck.initialize_model(model, path, ignore, exclude)
                else:
                    restored = ck.initialize_model(model, path, ignore, exclude)
                    assert sorted(restored) == sorted(names), (path, ignore, exclude)
                    for k in model_vars:
                        same = torch.equal(model.weights[k], torch.as_tensor(params[k]))
                        assert same == (k in names) or not params[k].any(), k


def _product_pointnet_common_on_cpu(monkeypatch):
    """The product's models/pointnet_common.py with its five CUDA operators replaced by the C oracle, so that the composition
    code around them (what the reference's file states in TF) can run without a GPU."""
    from oracle import ops
    pc = pkg("models.pointnet_common")
    f32 = lambda x: np.ascontiguousarray(x.detach().numpy(), np.float32)
    i32 = lambda x: np.ascontiguousarray(x.detach().numpy(), np.int32)
    ti = lambda a: torch.as_tensor(np.asarray(a, np.int32))

    def query_ball_point(radius, nsample, xyz1, xyz2):
        idx, cnt = ops.query_ball_point(radius, nsample, f32(xyz1), f32(xyz2))
        return ti(idx), ti(cnt)

    def knn_point(k, xyz1, xyz2):
        val, idx = ops.knn_point(k, f32(xyz1), f32(xyz2))
        return torch.as_tensor(val), ti(idx)

    monkeypatch.setattr(pc, "query_ball_point", query_ball_point)
    monkeypatch.setattr(pc, "knn_point", knn_point)
    # gathers return the dtype they were given (float64 callers hold float32-representable coordinates)
    monkeypatch.setattr(pc, "group_point", lambda points, idx: torch.as_tensor(ops.group_point(f32(points), i32(idx))).to(points.dtype))
    monkeypatch.setattr(pc, "farthest_point_sample", lambda npoint, inp: ti(ops.farthest_point_sample(npoint, f32(inp))))
    monkeypatch.setattr(pc, "gather_point", lambda inp, idx: torch.as_tensor(ops.gather_point(f32(inp), i32(idx))).to(inp.dtype))
    return pc


def test_pointnet_common_compositions_match_reference_golden(monkeypatch):
    """tests/golden/ref_net.npz holds the outputs of the reference's OWN models/pointnet_common.py (executed unmodified on the TF
    stand-in) for every branch of sample_points / query_and_group_points / sample_and_group / sample_and_group_all: with and
    without point features, use_xyz, kNN, radius normalisation, fed keypoints vs FPS, and the two opposite rotation conventions
    (:49-54 vs :112-119).  The product's compositions must return the same tensors in the same order."""
    spec = importlib.util.spec_from_file_location("make_golden_net", os.path.join(GOLD, "make_golden_net.py"))
    mkn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mkn)
    pc = _product_pointnet_common_on_cpu(monkeypatch)
    g = np.load(os.path.join(GOLD, "ref_net.npz"))
    inputs = mkn.composition_inputs()
    cases = mkn.composition_cases()
    assert len(cases) == 70
    checked_full = 0
    for i, (fn, kw) in enumerate(cases):
        outs = mkn.call_composition(pc, torch.as_tensor, fn, kw, *inputs)
        keys = [k for k in g.files if k.startswith("pointnet_common/%03d/" % i) and k.endswith("/digest")]
        assert len(outs) == len(keys), (fn, kw)
        for j, o in enumerate(outs):
            a = o.detach().numpy()
            want = g["pointnet_common/%03d/%d/digest" % (i, j)]
            got = mkn.digest(a, i * 16 + j)
            assert tuple(got[:5]) == tuple(want[:5]), (fn, kw, j, "shape")
            scale = np.sqrt((np.asarray(a, np.float64) ** 2).sum()) + 1e-12
            assert np.abs(got[5:] - want[5:]).max() <= 2e-5 * scale, (fn, kw, j)
            fk = "pointnet_common/%03d/%d/full" % (i, j)
            if fk in g.files:
                checked_full += 1
                if g[fk].dtype.kind == "i":
                    assert np.array_equal(a, g[fk]), (fn, kw, j)
                else:
                    assert np.allclose(a, g[fk], rtol=1e-5, atol=1e-5), (fn, kw, j)
    assert checked_full >= 2 + 2 + 7 + 4


def test_product_module_graph_matches_reference_graph_golden(monkeypatch):
    """The product's own differentiable statements -- models.feat3dnet.feature_detection_module / feature_extraction_module /
    Feat3dNet.get_loss over models.layers and models.pointnet_common -- run on the CPU in float64 (the five CUDA operators
    replaced by the C oracle) against the outputs of the reference's graph files executed unmodified (tests/golden/ref_net.npz):
    keypoints exact, attention / orientation / descriptors / BN shadow updates / loss to 1e-9, loss gradients to 1e-8.  The fused CUDA kernels are held
    to the same semantics by test_model_gpu.py and test_train_gpu.py."""
    import json
    from oracle import net as onet
    _product_pointnet_common_on_cpu(monkeypatch)
    f3 = pkg("models.feat3dnet")
    g = np.load(os.path.join(GOLD, "ref_net.npz"))
    names = sorted({k.split("/")[0] for k in g.files if k.endswith("/config")})
    assert len(names) == 9
    for name in names:
        cfg = json.loads(str(g[name + "/config"]))
        want = {k[len(name) + 5:]: g[k] for k in g.files if k.startswith(name + "/out/")}
        P = onet.to_torch(onet.init_params(seed=cfg["seed"], feature_dim=cfg["feature_dim"], randomize_bn=cfg.get("randomize_bn", True)), torch.float64,
                          requires_grad=cfg["training"])
        clouds = np.load(os.path.join(GOLD, cfg["fixture"])).astype(np.float32)[None] if cfg.get("fixture") else g[name + "/clouds"]
        xyz = torch.as_tensor(clouds[:, :, :3]).double().contiguous()
        kp = torch.as_tensor(g[name + "/keypoints"]).double() if name + "/keypoints" in g.files else None
        stats = {}
        new_xyz, idx, att, ori, _ = f3.feature_detection_module(
            xyz, None, cfg["num_clusters"], 2.0, cfg["training"], [64, 128, 256], [128, 64], num_samples=cfg["num_samples"],
            params=P, new_stats=stats, keypoints=kp)
        _, feats, ep = f3.feature_extraction_module(
            xyz, None, cfg["training"], [32, 64], [128 if cfg["feature_dim"] <= 64 else 256], [cfg["feature_dim"]],
            keypoints=new_xyz, orientations=None if cfg["no_regress"] else ori, radius=2.0, num_samples=cfg["num_samples"],
            params=P, new_stats=stats)
        assert np.array_equal(new_xyz.numpy(), want["xyz"]), name
        assert torch.allclose(att, torch.as_tensor(want["attention_end_point"]), rtol=1e-9, atol=1e-12), name
        d = ori - torch.as_tensor(want["orientation"])
        assert torch.atan2(torch.sin(d), torch.cos(d)).abs().max() < 1e-9, name
        assert torch.allclose(feats, torch.as_tensor(want["features"]), rtol=1e-9, atol=1e-11), name
        if "grouped_xyz" in want:
            assert torch.allclose(ep["grouped_xyz"], torch.as_tensor(want["grouped_xyz"]), rtol=1e-9, atol=1e-12)
            assert torch.allclose(ep["grouped_xyz_before"], torch.as_tensor(want["grouped_xyz_before"]), rtol=1e-12, atol=1e-14)
        if cfg["training"]:
            net = f3.Feat3dNet.__new__(f3.Feat3dNet)
            net.param = dict(Attention=cfg["attention"], margin=cfg["margin"])
            loss, _ = net.get_loss(None, torch.chunk(feats, 3, dim=0), torch.chunk(att, 3, dim=0)[0] if cfg["attention"] else None, {})
            assert abs(float(loss.detach()) - float(want["loss"])) < 1e-10, name
            updates = {k[len("bn_update/"):]: v for k, v in want.items() if k.startswith("bn_update/")}
            assert set(updates) == set(stats) and len(stats) == 18, name
            for k, v in updates.items():
                assert torch.allclose(stats[k].value(P[k]), torch.as_tensor(v), rtol=1e-9, atol=1e-12), (name, k)
            # d loss / d variable through the product's layers (autograd) == through the reference's graph
            leaves = sorted(k for k, v in P.items() if v.requires_grad)
            grads = torch.autograd.grad(loss, [P[k] for k in leaves], allow_unused=True)
            for k, gr in zip(leaves, grads):
                norm, proj = want["gradsum/" + k]
                gr = np.zeros(P[k].numel()) if gr is None else gr.numpy().ravel()
                r = np.random.default_rng(len(k)).standard_normal(gr.size)
                scale = max(norm, 1e-12)
                assert abs(np.sqrt((gr * gr).sum()) - norm) <= 1e-8 * scale + 1e-14, (name, k)
                assert abs(float(gr @ r) - proj) <= 1e-7 * scale * np.sqrt(gr.size) + 1e-14, (name, k)
        else:
            assert stats == {}

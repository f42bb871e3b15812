"""CPU tests of the host-side logic: argument validation mirrors the reference's OP_REQUIRES checks, the product
refuses to run without CUDA (no fallback), BN folding, batch sharding and the gloo world_size-2 gradient exchange."""
import importlib
import math
import os
import sys

import numpy as np
import pytest
import torch

from tests.conftest import pkg


def test_validation_errors_match_reference_checks(f3d_lib):
    ts = pkg("tf_ops.sampling.tf_sampling")
    tg = pkg("tf_ops.grouping.tf_grouping")
    x = torch.zeros((1, 8, 3))
    with pytest.raises(ValueError):
        ts.farthest_point_sample(0, x)  # tf_sampling.cpp:99 npoint > 0
    with pytest.raises(ValueError):
        ts.farthest_point_sample(4, torch.zeros((1, 8, 2)))  # :105 shape
    with pytest.raises(ValueError):
        ts.gather_point(torch.zeros((1, 8, 4)), torch.zeros((1, 2), dtype=torch.int32))  # :131
    with pytest.raises(ValueError):
        tg.query_ball_point(0.0, 4, x, x)  # tf_grouping.cpp:90 radius > 0
    with pytest.raises(ValueError):
        tg.query_ball_point(1.0, 0, x, x)  # :93 nsample > 0
    with pytest.raises(ValueError):
        tg.query_ball_point(1.0, 4, x, torch.zeros((2, 8, 3)))  # batch mismatch
    with pytest.raises(ValueError):
        tg.select_top_k(0, torch.zeros((1, 2, 3)))  # :179
    with pytest.raises(ValueError):
        tg.group_point(torch.zeros((1, 8)), torch.zeros((1, 2, 2), dtype=torch.int32))  # :215


def test_no_cpu_fallback(f3d_lib):
    """The product path fails loudly on CPU tensors instead of computing anything on the host."""
    tg = pkg("tf_ops.grouping.tf_grouping")
    lib_mod = pkg("_lib")
    x = torch.zeros((1, 8, 3))
    with pytest.raises(lib_mod.F3DError):
        tg.query_ball_point(1.0, 4, x, x)
    with pytest.raises(lib_mod.F3DError):
        pkg("tf_ops.sampling.tf_sampling").farthest_point_sample(2, x)


def test_product_does_not_import_oracle():
    """Nothing under 3dfeatnet_b200/ may import, call or link oracle/."""
    root = os.path.dirname(pkg().__file__)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_no_warp_collective_inside_a_conditional_expression():
    """A full-mask warp collective in one arm of `c ? a : b`, or behind `&&` / `||`, is skipped by the lanes whose condition differs.
    One such shuffle (`lt = serves ? ... __shfl_sync(kFull, re[2], wsrc) ... : 0` in bq_grid_query_grp_kernel) ended in "illegal
    instruction" on the device with 3072 / 6144-point index windows.  The arms that remain are guarded by warp-uniform values
    only (a ballot result, a value broadcast by a shuffle); anything new has to be hoisted or added here with its reason."""
    import re

    uniform_guards = (  # (file, guard as written): the condition is the same in every lane
        ("grouping.cu", "owners ?"), ("grouping.cu", "first < 0 && owners"),
    )
    root = os.path.join(os.path.dirname(pkg().__file__), "csrc")
    pat = re.compile(r"(\?|&&|\|\|)[^;]*(__shfl\w*_sync|__ballot_sync|__reduce_\w+_sync|__any_sync|__all_sync|__match_\w+_sync)\s*\(")
    found = []
    for f in sorted(os.listdir(root)):
        if not f.endswith((".cu", ".cuh")):
            continue
        for i, line in enumerate(open(os.path.join(root, f)).read().split("\n"), 1):
            code = line.split("//")[0]
            if pat.search(code) and not any(f == g[0] and g[1] in code for g in uniform_guards):
                found.append("%s:%d: %s" % (f, i, code.strip()))
    assert not found, found


def test_dropin_import_paths(f3d_lib):
    """The reference's own import lines resolve once the package directory is on sys.path."""
    pkg().install_dropin()
    try:
        m = importlib.import_module("tf_ops.grouping.tf_grouping")
        assert callable(m.query_ball_point) and callable(m.group_point) and callable(m.knn_point)
        m = importlib.import_module("tf_ops.sampling.tf_sampling")
        assert callable(m.farthest_point_sample) and callable(m.gather_point)
        m = importlib.import_module("models.pointnet_common")
        assert callable(m.sample_and_group) and callable(m.query_and_group_points)
        m = importlib.import_module("models.feat3dnet")
        assert callable(m.feature_detection_module) and callable(m.feature_extraction_module)
        assert importlib.import_module("models.net_factory").get_network("3DFeatNet") is m.Feat3dNet
        m = importlib.import_module("data.augment")  # train.py:12
        assert callable(m.get_augmentations_from_list) and issubclass(m.RotateZ, m.Augmentation)
        assert callable(importlib.import_module("data.datagenerator").DataGenerator)  # train.py:13
    finally:
        sys.path.remove(os.path.dirname(pkg().__file__))
        for k in [k for k in sys.modules if k.split(".")[0] in ("tf_ops", "models", "_lib", "data", "augment")]:
            del sys.modules[k]


def test_fold_params_equals_unfolded_layer(f3d_lib):
    """W' = W*s, b' = (b-mean)*s+beta reproduces conv+BN(eval) of models/layers.py for every layer."""
    from oracle import net as onet

    f3 = pkg("models.feat3dnet")
    layers = pkg("models.layers")
    params = {k: torch.as_tensor(v) for k, v in onet.init_params(seed=5, randomize_bn=True).items()}
    packed = f3.fold_params(params, 32)
    import ctypes

    nb = f3d_lib.f3d_packed_weights_num_blocks()
    offs, sizes = (ctypes.c_int * nb)(), (ctypes.c_int * nb)()
    assert f3d_lib.f3d_packed_weights_offsets(32, offs, sizes) == 0
    assert packed.numel() == f3d_lib.f3d_packed_weights_floats(32)
    g = torch.Generator().manual_seed(0)
    for i, (scope, cin, cout, bn) in enumerate(f3.DET_LAYERS + f3.desc_layers(32)):
        x = torch.randn((2, 3, 4, cin), generator=g)
        want = layers.conv2d(x, cout, [1, 1], padding='VALID', bn=bn, is_training=False, activation=None, scope=scope,
                             params=params)
        w = packed[offs[2 * i]:offs[2 * i] + cin * cout].reshape(cin, cout)
        b = packed[offs[2 * i + 1]:offs[2 * i + 1] + cout]
        got = x @ w + b
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), scope
        assert offs[2 * i] % 4 == 0 and offs[2 * i + 1] % 4 == 0


def test_unfused_layers_match_oracle_net_on_cpu():
    """models/layers.py + the loss follow the same semantics as oracle/net.py (both restate SURVEY.md appendix B)."""
    from oracle import net as onet

    layers = pkg("models.layers")
    P = {k: torch.as_tensor(v) for k, v in onet.init_params(seed=2, randomize_bn=True).items()}
    x = torch.randn((2, 5, 7, 3))
    for training in (False, True):
        a = layers.conv2d(x, 64, [1, 1], padding='VALID', bn=True, is_training=training, scope="detection/conv0", params=P)
        b = onet.conv2d(x, P, "detection/conv0", True, "relu", training)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    fa, fp, fn = torch.randn(3, 2, 9, 4).unbind(0)
    att = torch.rand(2, 9) + 0.1
    net = pkg("models.feat3dnet").Feat3dNet.__new__(pkg("models.feat3dnet").Feat3dNet)
    net.param = dict(Attention=True, margin=0.2)
    loss, _ = net.get_loss(None, (fa, fp, fn), att, {})
    assert torch.allclose(loss, onet.triplet_loss(fa, fp, fn, att, 0.2, True))


def test_chain_and_fused_helpers_have_no_cpu_side_door():
    """The training-path fusions of round 2 are device paths: on CPU tensors conv2d refuses a deferred activation (there is no kernel to
    form it), linear_rows / the local frames take the op-by-op torch statement, and a DeferredActivation can be materialised as the
    plain tensor it stands for."""
    from oracle import net as onet

    layers, pc = pkg("models.layers"), pkg("models.pointnet_common")
    P = {k: torch.as_tensor(v) for k, v in onet.init_params(seed=3, randomize_bn=True).items()}
    x = torch.randn((2, 5, 64, 3))
    with pytest.raises(ValueError, match="deferred activations"):
        layers.conv2d(x, 64, [1, 1], padding='VALID', bn=True, is_training=True, scope="detection/conv0", params=P, defer=True)
    with pytest.raises(ValueError, match="also_pool"):
        layers.conv2d(x, 64, [1, 1], padding='VALID', bn=True, is_training=True, scope="detection/conv0", params=P, also_pool=True)
    z = torch.randn(2, 5, 64, 8)
    coef = torch.cat((torch.rand(8) + 0.5, torch.randn(8)))
    d = layers.DeferredActivation(z, coef, True)
    assert d.shape == z.shape and d.dim() == 4 and not d.is_cuda
    assert torch.equal(d.materialize(), torch.relu(z * coef[:8] + coef[8:]))
    a, w = torch.randn(40, 64), torch.randn(64, 128)
    assert torch.equal(layers.linear_rows(a, w), a @ w)
    xyz, kp = torch.randn(2, 50, 3), torch.randn(2, 4, 3)
    idx = torch.randint(0, 50, (2, 4, 8), dtype=torch.int32)
    assert not pc._fused_frames_ok(xyz, kp, idx)


def test_unused_layer_helpers_follow_the_reference_semantics():
    """fully_connected / dropout / batch_norm_for_conv3d (layers.py:107-171,213-223): part of `models.layers`, never called
    by the model; checked against the TF statements written out in fp64."""
    layers = pkg("models.layers")
    g = torch.Generator().manual_seed(11)
    x = torch.randn((6, 5), generator=g)
    P = {"fc/weights": torch.randn((5, 4), generator=g), "fc/biases": torch.randn(4, generator=g),
         "fc/bn/gamma": torch.rand(4, generator=g) + 0.5, "fc/bn/beta": torch.randn(4, generator=g),
         "fc/bn/moving_mean": torch.randn(4, generator=g), "fc/bn/moving_variance": torch.rand(4, generator=g) + 0.1}
    lin = x.double() @ P["fc/weights"].double() + P["fc/biases"].double()
    assert torch.allclose(layers.fully_connected(x, 4, "fc", params=P).double(), lin.clamp_min(0), atol=1e-5)
    for training in (False, True):
        mean = lin.mean(0) if training else P["fc/bn/moving_mean"].double()
        var = lin.var(0, unbiased=False) if training else P["fc/bn/moving_variance"].double()
        want = (lin - mean) / torch.sqrt(var + 1e-3) * P["fc/bn/gamma"].double() + P["fc/bn/beta"].double()
        stats = {}
        got = layers.fully_connected(x, 4, "fc", activation_fn=None, bn=True, is_training=training, params=P, new_stats=stats)
        assert torch.allclose(got.double(), want, atol=1e-5)
        assert (set(stats) == {"fc/bn/moving_mean", "fc/bn/moving_variance"}) == training
    with pytest.raises(ValueError):
        layers.fully_connected(x, 3, "fc", params=P)
    with pytest.raises(ValueError):
        layers.fully_connected(x.reshape(2, 3, 5), 4, "fc", params=P)
    # dropout: identity in eval; in training kept entries are x / keep_prob and the mask follows noise_shape
    y = torch.ones((4, 64, 8))
    assert layers.dropout(y, False, "dp") is y
    torch.manual_seed(3)
    d = layers.dropout(y, True, "dp", keep_prob=0.25, noise_shape=[4, 1, 8])
    assert set(d.unique().tolist()) <= {0.0, 4.0}
    assert torch.equal(d, d[:, :1].expand_as(d)) and 0 < (d > 0).float().mean() < 1
    assert torch.equal(layers.dropout(y, True, "dp", keep_prob=1.0), y)
    with pytest.raises(ValueError):
        layers.dropout(y, True, "dp", keep_prob=0.0)
    # conv3d BN: moments over every axis but the channel one
    v = torch.randn((2, 3, 4, 5, 4), generator=g)
    Pb = {"bn/gamma": P["fc/bn/gamma"], "bn/beta": P["fc/bn/beta"]}
    got = layers.batch_norm_for_conv3d(v, True, None, "bn", Pb)
    flat = v.double().reshape(-1, 4)
    want = (flat - flat.mean(0)) / torch.sqrt(flat.var(0, unbiased=False) + 1e-3) * Pb["bn/gamma"].double() + Pb["bn/beta"].double()
    assert torch.allclose(got.double().reshape(-1, 4), want, atol=1e-5)


def test_shard_range_partitions():
    d = pkg("dist")
    for total in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [d.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    d = importlib.import_module("3dfeatnet_b200.dist")
    d.init("gloo")
    lo, hi = d.shard_range(10, rank, world)
    flat = torch.arange(107619, dtype=torch.float32) * (rank + 1)  # the model's 107 619 trainable floats
    d.allreduce_mean_(flat)
    mx = d.max_over_ranks(float(rank) + 0.5, torch.device("cpu"))
    # the training step's form: SUM all-reduce of the flat gradient, 1/world applied later inside the Adam kernel (grad_scale)
    summed = d.allreduce_sum_(torch.full((107619,), float(rank + 1)))
    d.barrier()
    q.put((rank, lo, hi, float(flat[1000]), mx, float(summed[5]) * (1.0 / world), d.env_world()))
    d.shutdown()


def test_gloo_world_size_2_gradient_exchange():
    """N>1 path on CPU: batch sharding + the single flat-buffer all-reduce (mean) of the training step."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]
    assert all(abs(r[3] - 1000 * 1.5) < 1e-3 for r in res)  # mean of 1x and 2x
    assert all(r[4] == 1.5 for r in res)
    assert all(r[5] == 1.5 for r in res)                     # sum (1 + 2) x grad_scale 1/2 == the mean
    assert [r[6] for r in res] == [(0, 0, 2), (1, 1, 2)]     # RANK / LOCAL_RANK / WORLD_SIZE as torchrun sets them


def test_bin_formats_roundtrip(tmp_path):
    """.bin reader (datagenerator.py:163-182) and the [xyz | descriptor] writer (inference.py:174-177)"""
    inf = pkg("inference")
    cloud = np.random.default_rng(0).random((100, 6)).astype(np.float32)
    p = tmp_path / "cloud.bin"
    cloud.tofile(p)
    assert np.array_equal(inf.load_point_cloud(str(p)), cloud)
    with pytest.raises(ValueError):
        inf.load_point_cloud(str(p), num_cols=7)
    xyz = torch.rand(17, 3)
    feat = torch.rand(17, 32)
    q = tmp_path / "out.bin"
    assert inf.save_keypoints_and_descriptors(str(q), xyz, feat) == (17, 35)
    back = np.fromfile(q, dtype=np.float32).reshape(17, 35)
    assert np.array_equal(back[:, :3], xyz.numpy()) and np.array_equal(back[:, 3:], feat.numpy())


def test_compute_descriptors_directory_flow(tmp_path, monkeypatch):
    """The host logic of compute_descriptors (inference.py:67-180) around a stand-in model: the .bin loop, --num_points,
    --randomize_points, --use_keypoints_from (<name>_kp.bin, 3 columns), data_dim, and the rows written per file.  The device
    work (attention, NMS, descriptors) is covered by test_nms_gpu.py::test_detect_nms_describe_flow."""
    inf = pkg("inference")
    rng = np.random.default_rng(4)
    data, out, kps = tmp_path / "data", tmp_path / "out", tmp_path / "kp"
    data.mkdir(); kps.mkdir()
    clouds = {}
    for name, n in (("b", 90), ("a", 120)):
        clouds[name] = rng.random((n, 6)).astype(np.float32)
        clouds[name].tofile(data / (name + ".bin"))
        rng.random((7 if name == "a" else 5, 3)).astype(np.float32).tofile(kps / (name + "_kp.bin"))
    (data / "notes.txt").write_text("not a cloud")

    class Model:
        calls = []

        def get_inference_model(self, pc, is_training, keypoints=None, fetch_features=True):
            assert is_training is False and pc.dim() == 3 and pc.shape[0] == 1
            Model.calls.append((tuple(pc.shape), tuple(keypoints.shape), fetch_features))
            feats = keypoints.sum(dim=2, keepdim=True).expand(-1, -1, 4) if fetch_features else None
            att = keypoints[:, :, 0]
            return keypoints, feats, att, {"attention": att}

    # described at the keypoints of <name>_kp.bin; the first num_points rows of each cloud are fed
    done = inf.compute_descriptors(Model(), str(data), str(out), num_points=80, use_keypoints_from=str(kps), device="cpu")
    assert done == ["a.bin", "b.bin"] and sorted(os.listdir(out)) == done
    assert Model.calls == [((1, 80, 6), (1, 7, 3), True), ((1, 80, 6), (1, 5, 3), True)]
    rows = np.fromfile(out / "a.bin", dtype=np.float32).reshape(7, 3 + 4)
    kp = np.fromfile(kps / "a_kp.bin", dtype=np.float32).reshape(7, 3)
    assert np.array_equal(rows[:, :3], kp) and np.allclose(rows[:, 3:], kp.sum(1, keepdims=True))
    with pytest.raises(ValueError):
        (kps / "a_kp.bin").write_bytes(b"")
        inf.compute_descriptors(Model(), str(data), str(out), use_keypoints_from=str(kps), device="cpu")

    # detected keypoints: attention at every point in MAX_POINTS chunks, NMS (stand-in: the 3 strongest), describe, keep `num` rows
    def fake_nms(xyz, attention, nms_radius, min_response_ratio, max_keypoints):
        top = attention.argsort(dim=1, descending=True)[:, :max_keypoints]
        sel = torch.gather(xyz, 1, top.unsqueeze(2).expand(-1, -1, 3))
        return sel, torch.gather(attention, 1, top), [3]

    monkeypatch.setattr(inf, "nms", fake_nms)
    monkeypatch.setattr(inf, "MAX_POINTS", 50)
    Model.calls = []
    shape = inf.compute_descriptors_for_file(Model(), str(data / "a.bin"), str(out / "a2.bin"), randomize_points=True, seed=9,
                                             max_keypoints=6, device="cpu")
    assert shape == (3, 7)
    assert [c[1][1] for c in Model.calls] == [50, 50, 20, 6] and [c[2] for c in Model.calls] == [False, False, False, True]
    rows = np.fromfile(out / "a2.bin", dtype=np.float32).reshape(3, 7)
    best = clouds["a"][np.argsort(-clouds["a"][:, 0], kind="stable")[:3], :3]
    assert np.array_equal(rows[:, :3], best)           # the permutation does not change which points win
    # one process per GPU: the sorted file list is cut into contiguous slices, every file is processed exactly once, and the
    # rows a rank writes equal the single-process ones
    for n_files in range(5):
        clouds["a"].tofile(data / ("shard%d.bin" % n_files))
    single = tmp_path / "single"
    all_files = inf.compute_descriptors(Model(), str(data), str(single), max_keypoints=6, device="cpu", randomize_points=True)
    for world in (2, 3, 8):
        sharded = tmp_path / ("world%d" % world)
        parts = [inf.compute_descriptors(Model(), str(data), str(sharded), max_keypoints=6, device="cpu", randomize_points=True,
                                         rank=r, world=world) for r in range(world)]
        assert [f for p in parts for f in p] == all_files and max(map(len, parts)) - min(map(len, parts)) <= 1
        for f in all_files:
            assert (sharded / f).read_bytes() == (single / f).read_bytes(), (world, f)
    # data_dim: a 3-column cloud file
    clouds["a"][:, :3].copy().tofile(data / "c3.bin")
    assert inf.compute_descriptors_for_file(Model(), str(data / "c3.bin"), str(out / "c3.bin"), max_keypoints=6, device="cpu",
                                            data_dim=3) == (3, 7)


def test_checkpoint_round_trips_and_restore_rules(tmp_path):
    """checkpoint.py: npz and TF tensor-bundle round trips, the TF name / layout mapping (1x1 kernels, EMA shadows,
    optimizer slots ignored) and initialize_model's exclude / ignore-missing rules (reference inference.py:183-217)."""
    import types
    from oracle import net as onet
    ck = importlib.import_module("3dfeatnet_b200.checkpoint")
    params = onet.init_params(seed=5, randomize_bn=True)
    tf_names = {}
    for k, v in params.items():                       # what a TF-1 checkpoint of the reference graph would hold
        if k.endswith("/conv2d/weights"):
            tf_names[k] = v.reshape(1, 1, *v.shape)
        elif k.endswith("/bn/moving_mean"):
            s = k[:-len("/bn/moving_mean")]
            tf_names["%s/bn/%s/bn/moments/Squeeze/ExponentialMovingAverage" % (s, s)] = v
        elif k.endswith("/bn/moving_variance"):
            s = k[:-len("/bn/moving_variance")]
            tf_names["%s/bn/%s/bn/moments/Squeeze_1/ExponentialMovingAverage" % (s, s)] = v
        else:
            tf_names[k] = v
    tf_names["detection/conv0/conv2d/weights/Adam"] = np.zeros((1, 1, 3, 64), np.float32)
    tf_names["beta1_power"] = np.float32(0.9)
    prefix = str(tmp_path / "model.ckpt-1000")
    ck.write_tf_bundle(prefix, tf_names)
    raw = ck.read_tf_bundle(prefix)
    assert set(raw) == set(tf_names) and all(np.array_equal(raw[k], np.asarray(tf_names[k], np.float32)) for k in tf_names)
    loaded = ck.load_checkpoint(prefix)
    assert set(loaded) == set(params) and all(np.array_equal(loaded[k], params[k]) for k in params)
    ck.save_npz(params, str(tmp_path / "w.npz"))
    again = ck.load_checkpoint(str(tmp_path / "w.npz"))
    assert all(np.array_equal(again[k], params[k]) for k in params)

    def fresh():
        init = onet.init_params(seed=9)
        m = types.SimpleNamespace(weights={k: torch.as_tensor(v.copy()) for k, v in init.items()}, invalidate=lambda: None)
        return m, init
    m, init = fresh()
    names = ck.initialize_model(m, prefix)
    assert len(names) == len(params) and all(np.array_equal(m.weights[k].numpy(), params[k]) for k in params)
    # --checkpoint <directory> (inference.py:189-190, train.py:192-193: tf.train.latest_checkpoint): the TF state file wins,
    # otherwise the highest step; load_checkpoint / initialize_model accept the directory itself
    ckdir = tmp_path / "ckdir"
    ckdir.mkdir()
    older = onet.init_params(seed=11, randomize_bn=True)
    ck.save_npz(older, str(ckdir / "checkpoint.ckpt-500.npz"))
    ck.write_tf_bundle(str(ckdir / "model.ckpt-1000"), tf_names)
    assert ck.latest_checkpoint(str(ckdir)) == str(ckdir / "model.ckpt-1000")
    from_dir = ck.load_checkpoint(str(ckdir))
    assert all(np.array_equal(from_dir[k], params[k]) for k in params)
    (ckdir / "checkpoint").write_text('model_checkpoint_path: "checkpoint.ckpt-500"\nall_model_checkpoint_paths: "checkpoint.ckpt-500"\n')
    assert ck.latest_checkpoint(str(ckdir)) == str(ckdir / "checkpoint.ckpt-500")
    m, init = fresh()
    ck.initialize_model(m, str(ckdir))
    assert all(np.array_equal(m.weights[k].numpy(), older[k]) for k in older)
    assert ck.latest_checkpoint(str(tmp_path / "nothing-here")) is None
    with pytest.raises(FileNotFoundError):
        (tmp_path / "empty").mkdir()
        ck.load_checkpoint(str(tmp_path / "empty"))
    m, init = fresh()                                   # stage 2 of train.sh: restore everything but the detector
    ck.initialize_model(m, prefix, restore_exclude=["detection"])
    assert all(np.array_equal(m.weights[k].numpy(), init[k] if k.startswith("detection/") else params[k]) for k in params)
    partial = {k: v for k, v in params.items() if not k.startswith("description/layer1/conv_post_0")}
    ck.save_npz(partial, str(tmp_path / "partial.npz"))
    m, init = fresh()
    with pytest.raises(KeyError):
        ck.initialize_model(m, str(tmp_path / "partial.npz"))
    ck.initialize_model(m, str(tmp_path / "partial.npz"), ignore_missing_vars=True)
    assert np.array_equal(m.weights["description/layer1/conv_post_0/conv2d/weights"].numpy(), init["description/layer1/conv_post_0/conv2d/weights"])
    (tmp_path / "bad.index").write_bytes(b"\x00" * 64)
    with pytest.raises(ValueError):
        ck.read_tf_bundle(str(tmp_path / "bad"))


def test_augmentations_follow_the_reference_semantics():
    """3dfeatnet_b200/augment.py vs the statements of the reference's data/augment.py: ranges, per-cloud draws, rigidity."""
    aug = importlib.import_module("3dfeatnet_b200.augment")
    g = torch.Generator().manual_seed(3)
    xyz = torch.randn(5, 200, 3, generator=g) * 10
    gen = torch.Generator().manual_seed(7)
    j = aug.jitter(xyz, gen=gen) - xyz
    assert j.abs().max() <= 0.05 + 1e-6 and 0.005 < j.std() < 0.02                       # N(0, 0.01) clipped at 0.05
    s = aug.shift(xyz, gen=gen) - xyz
    assert s.abs().max() <= 0.1 + 1e-6 and torch.allclose(s[:, :1].expand_as(s), s, atol=1e-5)          # one offset per cloud
    sc = aug.scale(xyz, gen=gen) / xyz
    assert (sc.min() >= 0.8 - 1e-4) and (sc.max() <= 1.25 + 1e-4) and sc.std(dim=(1, 2)).max() < 1e-4
    for fn in (aug.rotate_z, aug.rotate_small):
        r = fn(xyz, gen=gen)
        assert torch.allclose(r.norm(dim=2), xyz.norm(dim=2), rtol=1e-5, atol=1e-4)      # rigid
        mode = "donot_use_mm_for_euclid_dist"
        assert torch.allclose(torch.cdist(r, r, compute_mode=mode), torch.cdist(xyz, xyz, compute_mode=mode), rtol=1e-4, atol=1e-3)
    rz = aug.rotate_z(xyz, gen=gen)
    assert torch.allclose(rz[:, :, 2], xyz[:, :, 2])                                      # about the upright axis
    small = aug.rotate_small(xyz, gen=gen)
    cosang = torch.nn.functional.cosine_similarity(small.reshape(5, -1), xyz.reshape(5, -1), dim=1)
    assert (cosang > math.cos(3 * 0.18)).all()                                            # angles clipped at 0.18 rad per axis
    a = aug.apply_augmentations(xyz, gen=torch.Generator().manual_seed(1))
    b = aug.apply_augmentations(xyz, gen=torch.Generator().manual_seed(1))
    assert torch.equal(a, b) and not torch.equal(a, xyz)
    assert torch.equal(aug.apply_augmentations(xyz, names=()), xyz)
    ry = aug.rotate_y(xyz, gen=gen)
    assert torch.allclose(ry[:, :, 1], xyz[:, :, 1]) and torch.allclose(ry.norm(dim=2), xyz.norm(dim=2), rtol=1e-5, atol=1e-4)
    up1 = aug.apply_augmentations(xyz, names=("Rotate1D",), gen=gen, upright_axis=1)
    assert torch.allclose(up1[:, :, 1], xyz[:, :, 1]) and not torch.allclose(up1[:, :, 2], xyz[:, :, 2])
    assert torch.equal(aug.apply_augmentations(xyz, names=("Rotate1D",), gen=gen, upright_axis=0), xyz)


def test_augmentation_objects_keep_the_reference_interface():
    """data/augment.py:4-137: get_augmentations_from_list order and upright axis, `.apply` on the (N,3) NumPy clouds that
    DataGenerator.next_triplet passes, constructor parameters, batched torch input with one draw per cloud."""
    da = pkg("data.augment")
    assert da.get_augmentations_from_list(None) == []
    objs = da.get_augmentations_from_list(['Shift', 'RotateSmall', 'Scale', 'Jitter', 'Rotate1D'])
    assert [type(o) for o in objs] == [da.RotateZ, da.Jitter, da.Scale, da.RotateSmall, da.Shift]
    assert type(da.get_augmentations_from_list(['Rotate1D'], upright_axis=1)[0]) is da.RotateY
    assert da.get_augmentations_from_list(['Rotate1D'], upright_axis=0) == []
    assert all(isinstance(o, da.Augmentation) for o in objs)
    with pytest.raises(NotImplementedError):
        da.Augmentation().apply(np.zeros((4, 3), np.float32))
    rng = np.random.default_rng(0)
    cloud = (rng.normal(size=(300, 3)) * 8).astype(np.float32)
    torch.manual_seed(5)
    for o in objs + [da.RotateY()]:
        out = o.apply(cloud.copy())
        assert isinstance(out, np.ndarray) and out.shape == cloud.shape and out.dtype == np.float32
        assert not np.array_equal(out, cloud)
    j = da.Jitter(sigma=0.5, clip=0.2).apply(cloud) - cloud
    assert np.abs(j).max() <= 0.2 + 1e-6 and (np.abs(j) > 0.19).mean() > 0.3
    s = da.Shift(shift_range=3.0).apply(cloud) - cloud
    assert np.abs(s).max() <= 3.0 + 1e-5 and np.allclose(s, s[:1], atol=1e-5)
    sc = da.Scale(scale_low=2.0, scale_high=2.0).apply(cloud, keypoints=None)
    assert np.allclose(sc, 2 * cloud, rtol=1e-6)
    assert np.allclose(da.RotateSmall(angle_sigma=0.0).apply(cloud), cloud, atol=1e-5)
    assert np.allclose(da.RotateZ().apply(cloud)[:, 2], cloud[:, 2]) and np.allclose(da.RotateY().apply(cloud)[:, 1], cloud[:, 1])
    # batched tensors: one draw per cloud, reproducible through gen=
    batch = torch.as_tensor(np.stack([cloud, cloud]))
    out = da.Shift(gen=torch.Generator().manual_seed(2)).apply(batch)
    assert out.shape == batch.shape and not torch.allclose(out[0], out[1])
    assert torch.equal(out, da.Shift(gen=torch.Generator().manual_seed(2)).apply(batch))
    with pytest.raises(ValueError):
        da.Jitter().apply(np.zeros((5, 6), np.float32))


def _write_dataset(tmp_path, n_clouds=6, pts=300, seed=0):
    rng = np.random.default_rng(seed)
    lines = []
    for i in range(n_clouds):
        cloud = np.concatenate([rng.uniform(-25, 25, (pts, 3)), rng.normal(size=(pts, 3))], axis=1).astype(np.float32)
        cloud.tofile(str(tmp_path / ("c%d.bin" % i)))
        pos = [(i + 1) % n_clouds]
        nonneg = [(i + 2) % n_clouds]
        lines.append("c%d.bin | %s | %s" % (i, " ".join(map(str, pos)), " ".join(map(str, nonneg))))
    meta = tmp_path / "train.txt"
    meta.write_text("\n".join(lines) + "\n")
    return str(meta)


def test_datagenerator_follows_the_reference_interface(tmp_path):
    """data/datagenerator.py:9-182: metadata parsing, epoch bookkeeping, positive / negative rules, crop + resample"""
    dg_mod = pkg("data.datagenerator")
    meta = _write_dataset(tmp_path)
    gen = dg_mod.DataGenerator(meta, num_cols=6, seed=3)
    assert gen.size == 6 and gen.paths_and_labels[2] == ("c2.bin", {3}, {4})
    gen.shuffle()
    order = list(gen.indices)
    assert sorted(order) == list(range(6))
    seen = 0
    while True:
        a, p, n = gen.next_triplet(k=4, num_points=128)
        if a is None:
            break
        assert a.shape == p.shape == n.shape and a.shape[1:] == (128, 6) and a.dtype == np.float32
        assert (np.square(a[:, :, :3]).sum(2) <= 400.0 + 1e-3).all()  # 20 m crop (datagenerator.py:149-150)
        seen += a.shape[0]
    assert seen == 6  # k=4 then the 2 remaining, then (None, None, None)
    gen.reset()
    assert list(gen.indices) == list(range(6))
    for anchor in range(6):
        pos, neg = gen.get_positive_negative(anchor)
        assert pos == (anchor + 1) % 6 and neg not in {(anchor + 1) % 6, (anchor + 2) % 6}
    # fewer points than requested: every original row is kept and the rest are duplicates (datagenerator.py:153-160)
    small = np.concatenate([np.random.default_rng(1).uniform(-5, 5, (50, 3)), np.zeros((50, 3))], axis=1).astype(np.float32)
    out = gen.process_point_cloud(small, num_points=80)
    assert out.shape == (80, 6) and np.array_equal(out[:50], small)
    assert all(any(np.array_equal(r, s) for s in small) for r in out[50:])
    # augmentation objects with the reference's .apply(xyz) interface are applied to all three clouds
    class Shift1:
        def apply(self, xyz):
            return xyz + 1.0
    gen.reset()
    a0, _, _ = gen.next_triplet(k=1, num_points=64)
    gen.reset()
    gen.rng = np.random.default_rng(3)
    gen2 = dg_mod.DataGenerator(meta, num_cols=6, seed=11)
    b0, _, _ = gen2.next_triplet(k=1, num_points=64)
    gen3 = dg_mod.DataGenerator(meta, num_cols=6, seed=11)
    b1, _, _ = gen3.next_triplet(k=1, num_points=64, augmentation=[Shift1()])
    assert np.allclose(b1[:, :, :3], b0[:, :, :3] + 1.0) and np.array_equal(b1[:, :, 3:], b0[:, :, 3:])
    # ... and so are the reference-style objects of data/augment.py (train.py:45,146)
    gen4 = dg_mod.DataGenerator(meta, num_cols=6, seed=11)
    objs = pkg("data.augment").get_augmentations_from_list(['Jitter', 'Shift'])
    b2, _, _ = gen4.next_triplet(k=1, num_points=64, augmentation=objs)
    d = b2[:, :, :3] - b0[:, :, :3]
    assert b2.dtype == np.float32 and 0 < np.abs(d).max() <= 0.15 + 1e-5 and np.array_equal(b2[:, :, 3:], b0[:, :, 3:])
    with pytest.raises(ValueError):
        (tmp_path / "bad.txt").write_text("c0.bin | 1\n")
        dg_mod.DataGenerator(str(tmp_path / "bad.txt"))


def test_validation_metric_and_cluster_stacking(tmp_path):
    """train.py:240-315: ground-truth file, 100 m stacking of the validation clusters, FP rate at 95 % recall"""
    val = pkg("validation")
    gt = tmp_path / "groundtruths.txt"
    gt.write_text("idx1 idx2 t match\n" + "\n".join("%d %d 0.5 %d" % (i, i + 7, i % 2) for i in range(10)) + "\n")
    g = val.load_validation_groundtruths(str(gt))
    assert g == [(i, i % 2) for i in range(10)]
    assert val.load_validation_groundtruths(str(gt), proportion=0.5) == [(i, i % 2) for i in range(0, 10, 2)]
    rng = np.random.default_rng(0)
    pos, neg = rng.uniform(0, 1, 200), rng.uniform(0.5, 2, 300)
    thr = np.percentile(pos, 95)
    assert val.fp_rate_at_95_recall(pos, neg) == np.count_nonzero(neg < thr) / 300
    assert val.fp_rate_at_95_recall([0.1, 0.2], [0.5, 0.6]) == 0.0 and val.fp_rate_at_95_recall([1.0], [0.5]) == 1.0
    clouds = [rng.normal(size=(20 + j, 6)).astype(np.float32) for j in range(3)]
    pc, offsets = val.stack_clusters(clouds)
    assert pc.shape == (1, 63, 6) and offsets.shape == (1, val.NUM_CLUSTERS, 3)
    assert np.allclose(pc[0, 20:41, 0], clouds[1][:, 0] + 100.0) and np.array_equal(pc[0, 20:41, 1:], clouds[1][:, 1:])
    assert offsets[0, :4, 0].tolist() == [0.0, 100.0, 200.0, 0.0] and not offsets[0, :, 1:].any()
    assert val.validate(None, str(tmp_path), []) == 1  # nothing to validate (train.py:262-263)


def test_bench_reference_arm_prints_the_contracted_line_and_product_arm_needs_cuda():
    """bench.py --impl reference (the CPU statement of the path on the host cores) prints ONE JSON line with the contract's
    keys; the product arm has no CPU fallback: without a GPU it fails instead of timing anything."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    small = ["--steps", "1", "--warmup", "0", "--batch", "2", "--points", "2048", "--clusters", "64"]
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference"] + small, capture_output=True,
                       text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "keypoints+descriptors/sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["config"]["clouds_per_step"] == 2 and "2 clouds" in d["cpu_baseline"]["sample"] and "workload" in d["config"]
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py")] + small, capture_output=True, text=True, timeout=600,
                           cwd=root)
        assert r.returncode != 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def _tensor_bundle_protos():
    """BundleHeaderProto / BundleEntryProto / TensorShapeProto (tensorflow/core/protobuf/tensor_bundle.proto,
    framework/tensor_shape.proto: field numbers as published) built at run time with the protobuf library."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="f3d_test_tensor_bundle.proto", package="f3dtest", syntax="proto3")

    def msg(name, fields, parent=None):
        m = (parent.nested_type if parent is not None else fd.message_type).add(name=name)
        for fname, number, ftype, label, type_name in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = type_name
        return m

    shape = msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".f3dtest.TensorShapeProto.Dim"),
                                     ("unknown_rank", 3, F.TYPE_BOOL, F.LABEL_OPTIONAL, None)])
    msg("Dim", [("size", 1, F.TYPE_INT64, F.LABEL_OPTIONAL, None), ("name", 2, F.TYPE_STRING, F.LABEL_OPTIONAL, None)], parent=shape)
    msg("VersionDef", [("producer", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None), ("min_consumer", 2, F.TYPE_INT32, F.LABEL_OPTIONAL, None)])
    msg("BundleHeaderProto", [("num_shards", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None), ("endianness", 2, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                              ("version", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".f3dtest.VersionDef")])
    msg("BundleEntryProto", [("dtype", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                             ("shape", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".f3dtest.TensorShapeProto"),
                             ("shard_id", 3, F.TYPE_INT32, F.LABEL_OPTIONAL, None), ("offset", 4, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                             ("size", 5, F.TYPE_INT64, F.LABEL_OPTIONAL, None), ("crc32c", 6, F.TYPE_FIXED32, F.LABEL_OPTIONAL, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("f3dtest." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


def test_tf_bundle_reader_on_an_independently_written_index(tmp_path):
    """checkpoint.read_tf_bundle against an index it did not write: entries serialised by the protobuf LIBRARY, table blocks with
    leveldb prefix compression (restart interval 3) spread over several data blocks, two shards -- what a real TF Saver emits,
    where the package's own writer only ever uses shared = 0, one block, one shard.  CRC-32C against the RFC 3720 vectors."""
    import struct
    ck = importlib.import_module("3dfeatnet_b200.checkpoint")
    # CRC-32C (Castagnoli) known answers, and TF's mask ((crc >> 15 | crc << 17) + 0xa282ead8)
    unmask = lambda m: (lambda r: ((r >> 17) | (r << 15)) & 0xFFFFFFFF)((m - 0xA282EAD8) & 0xFFFFFFFF)
    for buf, want in ((b"123456789", 0xE3069283), (b"\x00" * 32, 0x8A9136AA), (b"\xff" * 32, 0x62A8AB43), (bytes(range(32)), 0x46DD794E)):
        assert unmask(ck._masked_crc32c(buf)) == want
    Header, Entry = _tensor_bundle_protos()
    rng = np.random.default_rng(12)
    names = ["detection/conv%d/conv2d/%s" % (i, leaf) for i in range(4) for leaf in ("biases", "weights")] + ["global_step"]
    arrays = {n: rng.normal(size=(1, 1, 3 + i, 4) if n.endswith("weights") else (4,)).astype(np.float32) for i, n in enumerate(names)}
    arrays["global_step"] = np.array(1234, np.int64)
    shard_bytes = [bytearray(), bytearray()]
    entries = [(b"", Header(num_shards=2, version=dict(producer=1)).SerializeToString())]
    for i, n in enumerate(sorted(arrays)):
        a, shard = arrays[n], i % 2
        e = Entry(dtype=9 if a.dtype == np.int64 else 1, shard_id=shard, offset=len(shard_bytes[shard]), size=a.nbytes,
                  crc32c=ck._masked_crc32c(a.tobytes()))
        for s in a.shape:
            e.shape.dim.add(size=s)
        got = ck._parse_entry(e.SerializeToString())
        assert (got["dtype"], got["shape"], got["shard_id"], got["offset"], got["size"]) == (e.dtype, list(a.shape), shard, e.offset, a.nbytes)
        shard_bytes[shard] += a.tobytes()
        entries.append((n.encode(), e.SerializeToString()))
    for s in range(2):
        (tmp_path / ("model.ckpt-7.data-%05d-of-00002" % s)).write_bytes(bytes(shard_bytes[s]))

    def varint(v):
        out = bytearray()
        while v >= 0x80:
            out.append((v & 0x7F) | 0x80)
            v >>= 7
        out.append(v)
        return bytes(out)

    def block(items, interval=3):  # leveldb block: shared-prefix compressed entries, restart array, restart count
        body, restarts, prev = bytearray(), [], b""
        for i, (k, v) in enumerate(items):
            shared = 0
            if i % interval == 0:
                restarts.append(len(body))
            else:
                while shared < min(len(k), len(prev)) and k[shared] == prev[shared]:
                    shared += 1
            body += varint(shared) + varint(len(k) - shared) + varint(len(v)) + k[shared:] + v
            prev = k
        return bytes(body) + b"".join(struct.pack("<I", r) for r in (restarts or [0])) + struct.pack("<I", max(len(restarts), 1))

    out = bytearray()

    def emit(blk):
        handle = varint(len(out)) + varint(len(blk))
        out.extend(blk + b"\x00" + struct.pack("<I", ck._masked_crc32c(blk + b"\x00")))
        return handle

    chunks = [entries[:4], entries[4:8], entries[8:]]
    index_items = [(c[-1][0] + b"\x00", emit(block(c))) for c in chunks]
    meta = emit(block([]))
    index = emit(block(index_items, interval=1))
    footer = meta + index
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xdb4775248b80fb57))
    (tmp_path / "model.ckpt-7.index").write_bytes(bytes(out))
    back = ck.read_tf_bundle(str(tmp_path / "model.ckpt-7"))
    assert sorted(back) == sorted(arrays)
    for n, a in arrays.items():
        assert back[n].dtype == a.dtype and back[n].shape == a.shape and np.array_equal(back[n], a), n
    model = ck.load_checkpoint(str(tmp_path / "model.ckpt-7"))  # TF layout -> model layout: (1,1,Cin,Cout) kernels become (Cin,Cout)
    assert model["detection/conv2/conv2d/weights"].shape == arrays["detection/conv2/conv2d/weights"].shape[2:]
    assert "global_step" not in model


class _StubTrainModel(object):
    """Stands in for Feat3dNet in the trainer's host logic: two 'scopes' of weights, a loss that depends on them, an SGD step."""
    made = []

    def __init__(self, param):
        self.param, self.device = dict(param), torch.device("cpu")
        g = torch.Generator().manual_seed(len(_StubTrainModel.made))
        self.weights = {"detection/conv0/conv2d/weights": torch.randn(3, 4, generator=g),
                        "description/layer1/conv0/conv2d/weights": torch.randn(3, 4, generator=g)}
        self.seen, self.invalidated = [], 0
        _StubTrainModel.made.append(self)

    def invalidate(self):
        self.invalidated += 1

    def get_train_model(self, a, p, n, is_training):
        assert is_training is True and a.shape == p.shape == n.shape and a.shape[2] == 3
        self.seen.append((a.clone(), p.clone(), n.clone()))
        return None, (a, p, n), None, {}

    def get_loss(self, xyz, features, att, ep):
        w = self.weights["description/layer1/conv0/conv2d/weights"]
        return (features[0].mean() - features[1].mean()) ** 2 + (w ** 2).mean(), ep

    def get_train_op(self, loss, lr=1e-5, end_points=None, grad_hook=None, grad_scale=1.0):
        self.weights["description/layer1/conv0/conv2d/weights"] *= (1 - lr)


def test_trainer_schedule_and_two_stage_recipe_on_a_stand_in_model(tmp_path):
    """trainer.train (train.py:93-184): epochs, short last batch ends the epoch, checkpoint / validation schedule, augmentation
    names or objects; trainer.train_two_stage (train.sh): descriptor-only pretraining, then the full model restored from the
    pretrain checkpoint except the `detection` scope, Rotate1D added.  The real model runs the same loop in test_train_gpu.py."""
    tr, dg_mod, ck = pkg("trainer"), pkg("data.datagenerator"), importlib.import_module("3dfeatnet_b200.checkpoint")
    meta = _write_dataset(tmp_path, n_clouds=7, pts=300, seed=2)
    data = dg_mod.DataGenerator(meta, num_cols=6, seed=1)
    _StubTrainModel.made = []
    model = _StubTrainModel({})
    steps_seen = []
    h = tr.train(model, data, num_epochs=2, batch_size=3, num_points=64, augmentation=("Shift",), lr=0.1,
                 checkpoint_dir=str(tmp_path / "ck"), checkpoint_every_n_steps=3, on_step=lambda s, hist: steps_seen.append(s))
    assert h["steps"] == 4 and steps_seen == [1, 2, 3, 4]          # 7 clouds / batch 3 -> 2 full batches per epoch, the short one dropped
    assert [os.path.basename(p) for p in h["checkpoints"]] == ["checkpoint.ckpt-3.npz"] and len(h["losses"]) == 4
    assert tr.latest_checkpoint(str(tmp_path / "ck")) == h["checkpoints"][0] and tr.latest_checkpoint(str(tmp_path / "none")) is None
    a, p, n = model.seen[0]
    assert a.shape == (3, 64, 3) and not torch.equal(a, p)
    # augmentation objects (data/augment.py) are applied on the tensors too; none leaves the generator's clouds untouched
    plain = _StubTrainModel({})
    dg_mod.DataGenerator(meta, num_cols=6, seed=5)
    tr.train(plain, dg_mod.DataGenerator(meta, num_cols=6, seed=5), batch_size=3, num_points=64, augmentation=(), max_steps=1)
    shifted = _StubTrainModel({})
    obj = pkg("data.augment").Shift(shift_range=5.0)
    tr.train(shifted, dg_mod.DataGenerator(meta, num_cols=6, seed=5), batch_size=3, num_points=64, augmentation=[obj], max_steps=1)
    d = shifted.seen[0][0] - plain.seen[0][0]
    assert d.abs().max() > 0.1 and torch.allclose(d, d[:, :1].expand_as(d), atol=1e-4)   # one offset per cloud
    # train.sh
    _StubTrainModel.made = []
    out = tr.train_two_stage(_StubTrainModel, dg_mod.DataGenerator(meta, num_cols=6, seed=3), str(tmp_path / "logs"),
                             param={"feature_dim": 32}, pretrain_epochs=1, num_epochs=2, batch_size=3, num_points=64, lr=0.1)
    m1, m2 = _StubTrainModel.made
    assert m1.param == dict(feature_dim=32, NoRegress=True, Attention=False) and m2.param == dict(feature_dim=32, NoRegress=False, Attention=True)
    assert out["model"] is m2 and out["pretrain"]["steps"] == 2 and out["secondstage"]["steps"] == 4
    assert os.path.basename(out["pretrain"]["checkpoints"][-1]) == "checkpoint.ckpt-2.npz"    # the final pretrain state is saved
    assert out["restored"] == ["description/layer1/conv0/conv2d/weights"] and m2.invalidated == 1
    saved = ck.load_checkpoint(out["pretrain"]["checkpoints"][-1])
    assert torch.allclose(torch.as_tensor(saved["description/layer1/conv0/conv2d/weights"]),
                          m1.weights["description/layer1/conv0/conv2d/weights"])
    # stage 2 started from the pretrained descriptor and its own (fresh) detector, and trained on from there
    assert torch.allclose(m2.weights["description/layer1/conv0/conv2d/weights"],
                          m1.weights["description/layer1/conv0/conv2d/weights"] * 0.9 ** 4)
    assert not torch.equal(m2.weights["detection/conv0/conv2d/weights"], m1.weights["detection/conv0/conv2d/weights"])
    assert os.path.isdir(tmp_path / "logs" / "secondstage" / "ckpt")

"""GPU tests (-m gpu) of the on-device keypoint NMS (inference.py:226-261), prob_sample / cumsum, and the two-pass
detect -> NMS -> describe flow, against the oracle (oracle/nms.py: the reference function on scikit-learn, and its
tree-free definition)."""
import numpy as np
import pytest
import torch

from oracle import nms as onms
from oracle import net as onet
from oracle import ops as oops
from tests.conftest import pkg

pytestmark = pytest.mark.gpu


def run_nms(xyz, att, dev, **kw):
    inf = pkg("inference")
    x, a, num, idx = inf.nms(torch.as_tensor(xyz).to(dev), torch.as_tensor(att).to(dev), return_indices=True, **kw)
    return x.cpu().numpy(), a.cpu().numpy(), num, idx.cpu().numpy()


@pytest.mark.parametrize("kind,n", [("oxford", 16384), ("kitti", 29291), ("oxford", 3000)])
def test_nms_matches_reference_function(cuda, kind, n):
    """the reference's nms() verbatim (sklearn BallTree, 50-NN) on real-shaped clouds with random attention"""
    synth = pkg("synth")
    base = synth.base_cloud(kind)
    xyz = base[None, :n].copy() if n <= base.shape[0] else synth.make_batch(1, n, kind=kind)
    rng = np.random.default_rng(n)
    att = np.log1p(np.exp(rng.standard_normal((1, xyz.shape[1])).astype(np.float32) * 2)).astype(np.float32)
    want = onms.nms(xyz, att)
    got = run_nms(xyz, att, cuda)
    assert got[2] == want[2]
    assert np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_nms_kitti_scale_batch_matches_reference_function(cuda):
    """W4 size (131 072 points, KITTI-shape, >49 in-radius neighbours for some points) as a batch of two clouds with
    different extents: the grid-binned kernel against the reference's nms() on the sklearn BallTree."""
    synth = pkg("synth")
    xyz = synth.make_batch(2, 131072, seed0=77, kind="kitti")
    xyz[1] *= np.float32(0.6)                      # a second, denser cloud: another grid and many truncated neighbourhoods
    rng = np.random.default_rng(1)
    att = np.log1p(np.exp(rng.standard_normal(xyz.shape[:2]).astype(np.float32) * 2)).astype(np.float32)
    want = onms.nms(xyz, att)
    got = run_nms(xyz, att, cuda)
    assert got[2] == want[2]
    assert np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_nms_dense_cloud_uses_50nn_truncation(cuda):
    """more than 49 neighbours inside the radius: only the 49 nearest are consulted (inference.py:236-237)"""
    rng = np.random.default_rng(5)
    xyz = (rng.random((2, 1500, 3)) * np.array([3.0, 3.0, 0.5])).astype(np.float32)  # ~90 points per r=0.5 ball
    att = rng.random((2, 1500)).astype(np.float32) + 0.01
    want = onms.nms(xyz, att, max_keypoints=256)
    brute = onms.nms_bruteforce(xyz, att, max_keypoints=256)
    assert np.array_equal(want[3], brute[3])
    got = run_nms(xyz, att, cuda, max_keypoints=256)
    assert got[2] == want[2] and np.array_equal(got[3], want[3])
    # and it differs from a plain radius test, i.e. the truncation path really ran
    assert want[2] != onms.nms_bruteforce(xyz, att, max_keypoints=256, n_neighbors=1500)[2]


def test_nms_threshold_padding_and_ties(cuda):
    rng = np.random.default_rng(6)
    xyz = (rng.random((1, 400, 3)) * 40).astype(np.float32)  # sparse: every point is its own maximum
    att = np.full((1, 400), 0.5, np.float32)                 # all equal: order is by index, descending
    att[0, 7] = 100.0                                        # threshold = 1.0 -> only point 7 survives
    got = run_nms(xyz, att, cuda, max_keypoints=16)
    want = onms.nms(xyz, att, max_keypoints=16)
    assert got[2] == want[2] == [1] and np.array_equal(got[3], want[3]) and (got[3] == 7).all()
    att[0, 7] = 0.5
    got = run_nms(xyz, att, cuda, max_keypoints=16)
    want = onms.nms(xyz, att, max_keypoints=16)
    assert got[2] == want[2] == [16] and np.array_equal(got[3], want[3])
    assert got[3][0].tolist() == list(range(399, 383, -1))   # (attention, index) tuples sorted in reverse
    inf = pkg("inference")
    with pytest.raises(ValueError):
        inf.nms(torch.zeros((1, 10, 3), device=cuda), torch.zeros((1, 10), device=cuda))  # fewer points than 50 neighbours


@pytest.mark.parametrize("levels,kmax", [(4, 1024), (1, 1024), (4, 100), (64, 3000), (3, 6000)])
def test_nms_top_k_cut_inside_a_tie_group(cuda, levels, kmax):
    """thousands of survivors whose attentions take a few values only: the cut at max_keypoints falls inside a group of equal attentions
    and is decided by the point index (the reference sorts (attention, index) tuples in reverse, inference.py:249-251) -- the radix select
    of nms_topk_kernel runs on into the index bits; kmax > survivors: everything kept, padded with the best"""
    rng = np.random.default_rng(levels * 7 + kmax)
    xyz = (rng.random((2, 5000, 3)) * 400).astype(np.float32)  # sparse: every point is its own maximum
    att = (0.5 + rng.integers(0, levels, (2, 5000)) / 8.0).astype(np.float32)
    want = onms.nms(xyz, att, max_keypoints=kmax)
    got = run_nms(xyz, att, cuda, max_keypoints=kmax)
    assert got[2] == want[2]
    assert np.array_equal(got[3], want[3])
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])


def test_nms_matches_reference_golden(cuda):
    """tests/golden/ref_nms.npz: outputs of the reference's own nms() (taken out of inference.py with `ast` and executed
    unmodified, tests/golden/make_golden_nms.py) -- 1024-keypoint truncation, the 50-NN rule on dense clouds, padding."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nms.npz"))
    for name in sorted({k.split("/")[0] for k in g.files}):
        radius, ratio, kmax = g[name + "/cli"]
        got = run_nms(g[name + "/xyz"], g[name + "/attention"], cuda, nms_radius=float(radius), min_response_ratio=float(ratio),
                      max_keypoints=int(kmax))
        assert got[2] == g[name + "/num_keypoints"].tolist(), name
        assert np.array_equal(got[0], g[name + "/xyz_nms"]) and np.array_equal(got[1], g[name + "/attention_nms"]), name


def test_cumsum_and_prob_sample_bit_exact(cuda):
    ts = pkg("tf_ops.sampling.tf_sampling")
    lib_mod = pkg("_lib")
    rng = np.random.default_rng(12)
    for b, n, m in [(3, 20000, 500), (2, 8192, 64), (1, 8195, 100), (2, 5, 16), (1, 1, 4), (2, 40001, 1000)]:
        w = rng.random((b, n), dtype=np.float32)
        r = rng.random((b, m), dtype=np.float32)
        wd = torch.as_tensor(w).to(cuda)
        out = torch.empty_like(wd)
        lib_mod.check(lib_mod.lib().f3d_cumsum(b, n, lib_mod.ptr(wd), lib_mod.ptr(out), lib_mod.stream()), "cumsum")
        assert np.array_equal(out.cpu().numpy(), oops.cumsum(w)), (b, n)
        got = ts.prob_sample(wd, torch.as_tensor(r).to(cuda)).cpu().numpy()
        assert np.array_equal(got, oops.prob_sample(w, r)), (b, n, m)
    from oracle import ref as oref
    if oref.available("libref_sampling.so"):
        w = rng.random((4, 30000), dtype=np.float32)
        r = rng.random((4, 2000), dtype=np.float32)
        ridx, rcum = oref.gpu_prob_sample(torch.as_tensor(w).to(cuda), torch.as_tensor(r).to(cuda))
        assert np.array_equal(rcum.cpu().numpy(), oops.cumsum(w))  # pins the oracle's summation DAG to the reference kernel
        assert torch.equal(ts.prob_sample(torch.as_tensor(w).to(cuda), torch.as_tensor(r).to(cuda)), ridx)


def test_detect_nms_describe_flow(cuda):
    """inference.py:115-171 on one cloud: attention at every point (M = N), NMS, descriptors at the survivors"""
    f3 = pkg("models.feat3dnet")
    inf = pkg("inference")
    xyz = pkg("synth").make_batch(1, 4096, seed0=77)
    params = onet.init_params(seed=8, randomize_bn=True)
    net = f3.Feat3dNet({'num_clusters': -1}, weights=params, device=cuda, precision="fp32")
    pc = torch.as_tensor(xyz).to(cuda)
    kp, feat, att, num = inf.detect_and_describe(net, pc, max_keypoints=128)
    det = onet.detector(xyz, onet.to_torch(params, torch.float64), -1, 2.0, 64, dtype=torch.float64)
    a_ref = det["attention"].float().numpy()
    want = onms.nms(xyz, a_ref, max_keypoints=128)
    assert num == want[2]
    # attention differs in the last bits between fp32 GPU and fp64 oracle: compare the selected SETS when no near-ties exist
    got_idx = inf.nms(pc, torch.as_tensor(a_ref).to(cuda), max_keypoints=128, return_indices=True)[3].cpu().numpy()
    assert np.array_equal(got_idx, want[3])
    desc = onet.descriptor(xyz, onet.to_torch(params, torch.float64), kp.cpu().numpy(),
                           onet.detector(xyz, onet.to_torch(params, torch.float64), 0, 2.0, 64, keypoints_np=kp.cpu().numpy(),
                                         dtype=torch.float64)["orientation"], 2.0, 64, dtype=torch.float64)
    assert (feat.cpu().double() - desc["features"]).abs().max().item() < 1e-4

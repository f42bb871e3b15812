"""Matching + RANSAC registration: the numpy oracle against known transforms (CPU), and the device
path against the oracle on the same sample triples (-m gpu).  Parity with MATLAB itself is unpinned (see oracle/registration.py)."""
import numpy as np
import pytest
import torch

from oracle import registration as oreg
from tests.conftest import pkg


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    return oreg.quat2rot(q)


def _pair(rng, n=400, outlier_frac=0.4, noise=0.02):
    """pts1 ~ R pts2 + t for the inliers; the rest are random correspondences"""
    R, t = _rot(rng), rng.uniform(-5, 5, 3)
    pts2 = rng.uniform(-20, 20, (n, 3))
    pts1 = pts2 @ R.T + t + rng.normal(0, noise, (n, 3))
    out = rng.random(n) < outlier_frac
    pts1[out] = rng.uniform(-20, 20, (int(out.sum()), 3))
    return pts1.astype(np.float32), pts2.astype(np.float32), R, t, ~out


def test_oracle_rigid_transform_recovers_a_known_motion():
    rng = np.random.default_rng(0)
    for n in (3, 4, 50):
        R, t = _rot(rng), rng.uniform(-3, 3, 3)
        y = rng.uniform(-10, 10, (3, n))
        x = R @ y + t[:, None]
        T, eps = oreg.estimate_rigid_transform(x, y)
        assert np.allclose(T[:3, :3], R, atol=1e-9) and np.allclose(T[:3, 3], t, atol=1e-8) and eps < 1e-8
        assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9


def test_oracle_ransac_follows_the_sequential_rules():
    rng = np.random.default_rng(1)
    pts1, pts2, R, t, inl = _pair(rng)
    triples = pkg("registration").draw_triples(len(pts1), 2000, seed=5)
    assert triples.min() >= 0 and triples.max() < len(pts1)
    assert (triples[:, 0] != triples[:, 1]).all() and (triples[:, 0] != triples[:, 2]).all() and (triples[:, 1] != triples[:, 2]).all()
    Rt, idx, trialcount, best = oreg.ransacfit_rt(pts1, pts2, 1.0, triples)
    assert 10 <= trialcount <= 2000 and 0 <= best < trialcount  # "N = max(N, 10)" (ransac.m:181)
    assert np.allclose(Rt[:, :3], R, atol=5e-3) and np.allclose(Rt[:, 3], t, atol=5e-2)
    assert set(np.flatnonzero(inl)) <= set(idx) or len(set(np.flatnonzero(inl)) - set(idx)) <= 2
    # three correspondences: all of them are inliers, no trial (ransacfitRt.m:49-54); fewer: empty model
    Rt3, idx3, tc3, _ = oreg.ransacfit_rt(pts1[inl][:3], pts2[inl][:3], 1.0, triples)
    assert tc3 == 0 and list(idx3) == [0, 1, 2] and Rt3.shape == (3, 4)
    assert oreg.ransacfit_rt(pts1[:2], pts2[:2], 1.0, triples)[0] is None


def test_registration_refuses_cpu_tensors(f3d_lib):
    reg = pkg("registration")
    with pytest.raises(Exception):
        reg.match_descriptors(torch.zeros(4, 8), torch.zeros(4, 8))
    with pytest.raises(Exception):
        reg.ransacfitRt(torch.zeros(5, 3), torch.zeros(5, 3), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("n1,n2,d", [(1024, 1024, 32), (37, 501, 32), (300, 70, 128), (5, 1, 16)])
def test_match_descriptors_vs_oracle(cuda, n1, n2, d):
    reg = pkg("registration")
    rng = np.random.default_rng(n1 + n2)
    d1 = rng.normal(size=(n1, d)).astype(np.float32)
    d2 = rng.normal(size=(n2, d)).astype(np.float32)
    if n2 > 10:  # exact ties: duplicated candidate rows -> the lowest index wins ('smallest' keeps the first)
        d2[n2 // 2] = d2[3]
        d2[n2 - 1] = d2[3]
        d1[0] = d2[3]
    m, dist = reg.match_descriptors(torch.as_tensor(d1).to(cuda), torch.as_tensor(d2).to(cuda), return_dist=True)
    m, dist = m.cpu().numpy(), dist.cpu().numpy()
    ref_m, ref_d = oreg.match_descriptors(d1, d2)
    all_d = ((d1[:, None, :].astype(np.float64) - d2[None].astype(np.float64)) ** 2).sum(2)
    chosen = all_d[np.arange(n1), m]
    assert np.all(chosen <= ref_d * (1 + 1e-5) + 1e-6), "a returned match is not a nearest neighbour"
    margin = np.partition(all_d, 1, axis=1)[:, 1] - ref_d if n2 > 1 else np.ones(n1)
    clear = margin > 1e-4 * (1 + ref_d)
    assert np.array_equal(m[clear], ref_m[clear])
    assert np.allclose(dist, ref_d, rtol=1e-5, atol=1e-5)
    if n2 > 10:
        assert m[0] == 3


@pytest.mark.gpu
def test_rigid_fit_vs_oracle(cuda):
    reg = pkg("registration")
    rng = np.random.default_rng(3)
    for n in (3, 7, 1000):
        pts1, pts2, R, t, _ = _pair(rng, n=n, outlier_frac=0.0, noise=0.05)
        Rt = reg.estimateRt(torch.as_tensor(pts1).to(cuda), torch.as_tensor(pts2).to(cuda)).cpu().numpy()
        ref = oreg.estimate_rigid_transform(pts1.T, pts2.T)[0][:3]
        assert np.allclose(Rt, ref, atol=1e-8), "n=%d: max diff %.3e" % (n, np.abs(Rt - ref).max())
        mask = rng.random(n) < 0.6
        if mask.sum() >= 3:
            Rtm = reg.estimateRt(torch.as_tensor(pts1).to(cuda), torch.as_tensor(pts2).to(cuda), torch.as_tensor(mask).to(cuda)).cpu().numpy()
            assert np.allclose(Rtm, oreg.estimate_rigid_transform(pts1[mask].T, pts2[mask].T)[0][:3], atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("n,outliers,seed", [(1024, 0.5, 0), (400, 0.8, 1), (64, 0.3, 2), (3, 0.0, 3)])
def test_ransac_vs_oracle_on_the_same_triples(cuda, n, outliers, seed):
    """same samples -> same trial count, same chosen trial's inlier set, same refit (fp64 on both sides)"""
    reg = pkg("registration")
    rng = np.random.default_rng(seed)
    pts1, pts2, R, t, _ = _pair(rng, n=n, outlier_frac=outliers)
    triples = reg.draw_triples(max(n, 3), 10001, seed=seed + 10) if n > 3 else np.zeros((0, 3), np.int32)
    Rt, inl, trialcount = reg.ransacfitRt(torch.as_tensor(pts1).to(cuda), torch.as_tensor(pts2).to(cuda), 1.0, triples=triples)
    ref_Rt, ref_inl, ref_trials, _ = oreg.ransacfit_rt(pts1, pts2, 1.0, triples)
    assert trialcount == ref_trials
    assert np.array_equal(inl.cpu().numpy(), ref_inl)
    assert np.allclose(Rt.cpu().numpy(), ref_Rt, atol=1e-8)
    if n > 3:
        assert np.allclose(ref_Rt[:, :3], R, atol=2e-2)


@pytest.mark.gpu
def test_register_two_clouds_end_to_end(cuda):
    """computeAndVisualizeMatches.m:42-52: descriptors -> matches -> RANSAC recovers the motion between two keypoint sets"""
    reg = pkg("registration")
    rng = np.random.default_rng(9)
    n, d = 1024, 32
    R, t = _rot(rng), rng.uniform(-5, 5, 3)
    xyz2 = rng.uniform(-30, 30, (n, 3)).astype(np.float32)
    desc2 = rng.normal(size=(n, d)).astype(np.float32)
    perm = rng.permutation(n)
    xyz1 = (xyz2[perm] @ R.T + t).astype(np.float32)
    desc1 = desc2[perm] + rng.normal(0, 0.05, (n, d)).astype(np.float32)
    bad = rng.random(n) < 0.5  # half of the keypoints of cloud 1 have unrelated descriptors -> wrong matches
    desc1[bad] = rng.normal(size=(int(bad.sum()), d)).astype(np.float32)
    Rt, inl, trials, matches = reg.register(torch.as_tensor(xyz1).to(cuda), torch.as_tensor(desc1).to(cuda),
                                            torch.as_tensor(xyz2).to(cuda), torch.as_tensor(desc2).to(cuda), t=1.0, seed=1)
    Rt = Rt.cpu().numpy()
    assert np.allclose(Rt[:, :3], R, atol=1e-3) and np.allclose(Rt[:, 3], t, atol=2e-2)
    good = np.flatnonzero(~bad)
    assert np.array_equal(matches.cpu().numpy()[good, 1], perm[good])
    assert len(set(good) - set(inl.cpu().numpy().tolist())) == 0 and trials >= 10


def test_oracle_rigid_fit_agrees_with_scipy():
    """The registration oracle restates MATLAB code that cannot run here (parity unpinned); as a second, independent statement
    of the same mathematics SciPy's Rotation is used: quat2rot (w-first quaternion) == Rotation.from_quat (x,y,z,w), and the
    quaternion least-squares fit of estimateRigidTransform.m == Rotation.align_vectors (Kabsch) on noisy correspondences."""
    from scipy.spatial.transform import Rotation
    from oracle import registration as oreg
    rng = np.random.default_rng(3)
    for _ in range(5):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        assert np.allclose(oreg.quat2rot(q), Rotation.from_quat([q[1], q[2], q[3], q[0]]).as_matrix(), atol=1e-12)
    for n, noise in ((3, 0.0), (10, 0.01), (200, 0.05)):
        y = rng.normal(size=(3, n)) * 5
        R = Rotation.from_rotvec(rng.normal(size=3)).as_matrix()
        t = rng.normal(size=3) * 3
        x = R @ y + t[:, None] + rng.normal(size=(3, n)) * noise
        T, eps = oreg.estimate_rigid_transform(x, y)
        xc, yc = x.mean(1, keepdims=True), y.mean(1, keepdims=True)
        kabsch, _ = Rotation.align_vectors((x - xc).T, (y - yc).T)       # rotation taking y-centred onto x-centred, least squares
        assert np.allclose(T[:3, :3], kabsch.as_matrix(), atol=1e-8), (n, noise)
        assert np.allclose(T[:3, 3], (xc - kabsch.as_matrix() @ yc)[:, 0], atol=1e-8)
        assert np.allclose(T[3], [0, 0, 0, 1]) and abs(np.linalg.det(T[:3, :3]) - 1) < 1e-10 and eps >= -1e-12
        if noise == 0.0:
            assert np.allclose(T[:3, :3], R, atol=1e-9) and np.allclose(T[:3, 3], t, atol=1e-8)

"""CPU tests (-m "not gpu"): the oracle against every golden vector / known-answer / property the reference holds
for this path, plus semantics the reference never tested but the CUDA path must match."""
import os
import re

import numpy as np
import pytest

from oracle import ops, ref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


# ---------------------------------------------------------------------------- reference golden vectors
def test_selection_sort_known_answer():
    """tf_ops/grouping/test/selection_sort.cpp:65-93: b=2,n=4,m=2,k=3, dist[i]=10-i.  The committed text is the
    stdout of that program built as-is (tests/golden/make_golden_cpu.py)."""
    lines = open(os.path.join(GOLD, "selection_sort_ref.txt")).read().strip().splitlines()
    ref_idx = np.array(lines[-2].split(), dtype=np.int32)
    ref_val = np.array(lines[-1].split(), dtype=np.float32)
    dist = (10 - np.arange(16, dtype=np.float32)).reshape(2, 2, 4)
    outi, out = ops.select_top_k(3, dist)
    assert np.array_equal(outi.ravel(), ref_idx)
    assert np.array_equal(out.ravel(), ref_val)
    assert np.array_equal(ref_idx, np.tile([3, 2, 1, 0], 4))


def test_selection_sort_is_unstable_like_the_reference():
    """SURVEY.md A-6 (sim-verified): [2,2,1,2,1,3], k=4 -> [2,4,0,3]; a stable sort would give [2,4,0,1]."""
    outi, _ = ops.select_top_k(4, np.array([[[2, 2, 1, 2, 1, 3]]], np.float32))
    assert outi[0, 0, :4].tolist() == [2, 4, 0, 3]


def test_oracle_matches_reference_cpu_loops_golden():
    """Committed outputs of the reference's CPU loops (test/query_ball_point.cpp:19-84)."""
    g = np.load(os.path.join(GOLD, "ref_cpu_grouping.npz"))
    idx, cnt = ops.query_ball_point(float(g["radius"]), int(g["nsample"]), g["xyz1"], g["xyz2"])
    assert (cnt > 0).all()
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(ops.group_point(g["points"], idx), g["grouped"])
    assert np.array_equal(ops.group_point_grad(g["points"], idx, g["grad"]), g["grad_points"])


@pytest.mark.skipif(not ref.available("libref_cpu_grouping.so"), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_reference_cpu_loops_live():
    rng = np.random.default_rng(5)
    xyz1 = rng.random((3, 700, 3), dtype=np.float32)
    xyz2 = rng.random((3, 90, 3), dtype=np.float32)
    idx, cnt = ops.query_ball_point(0.15, 16, xyz1, xyz2)
    ridx = ref.cpu_query_ball_point(0.15, 16, xyz1, xyz2, fill=-7)
    ne = cnt > 0
    assert ne.any() and (~ne).any() or ne.all()
    assert np.array_equal(idx[ne], ridx[ne])  # the CPU program has no empty-ball branch
    pts = rng.random((3, 700, 5), dtype=np.float32)
    safe = np.where(idx < 0, 0, idx)
    assert np.array_equal(ops.group_point(pts, safe), ref.cpu_group_point(pts, safe))
    g = rng.standard_normal((3, 90, 16, 5)).astype(np.float32)
    assert np.array_equal(ops.group_point_grad(pts, safe, g), ref.cpu_group_point_grad(pts, safe, g))


def test_query_ball_point2_property_like_reference_test():
    """tf_grouping_op_test.py:32-65 re-expressed: pts_cnt == #(dist < radii) and set(idx row) == in-ball set."""
    from scipy.spatial.distance import cdist

    rng = np.random.RandomState(0)
    xyz1 = rng.random_sample((1, 128, 3)).astype("float32")
    xyz2 = rng.random_sample((1, 8, 3)).astype("float32")
    radii = rng.uniform(low=0.2, high=0.4, size=(1, 8)).astype("float32")
    idx, pts_cnt = ops.query_ball_point2(radii, 32, xyz1, xyz2)
    assert idx.max() < 128 and pts_cnt.max() <= 32
    Y = cdist(xyz1[0].astype(np.float64), xyz2[0].astype(np.float64))
    within = Y < radii[0][None, :]
    assert np.array_equal(pts_cnt[0], within.sum(0))
    for j in range(8):
        assert set(idx[0, j]) == set(np.nonzero(within[:, j])[0])


def test_group_point_grad_numeric_like_reference_test():
    """tf_grouping_op_test.py:10-27 re-expressed: analytic GroupPointGrad vs numeric Jacobian, err < 1e-4.
    group_point is linear in `points`, so J^T g is compared with a central difference of <group_point(p), g>."""
    rng = np.random.RandomState(1)
    points = rng.random_sample((1, 128, 16)).astype("float32")
    xyz1 = rng.random_sample((1, 128, 3)).astype("float32")
    xyz2 = rng.random_sample((1, 8, 3)).astype("float32")
    idx, _ = ops.query_ball_point(0.3, 32, xyz1, xyz2)
    idx = np.where(idx < 0, 0, idx)
    g = rng.standard_normal((1, 8, 32, 16)).astype("float32")
    analytic = ops.group_point_grad(points, idx, g).astype(np.float64)
    eps = 1e-2
    num = np.zeros_like(analytic)
    for i in rng.choice(128 * 16, 64, replace=False):
        d = np.zeros(128 * 16, np.float32)
        d[i] = eps
        d = d.reshape(1, 128, 16)
        fp = (ops.group_point(points + d, idx).astype(np.float64) * g).sum()
        fm = (ops.group_point(points - d, idx).astype(np.float64) * g).sum()
        num.reshape(-1)[i] = (fp - fm) / (2 * eps)
        assert abs(num.reshape(-1)[i] - analytic.reshape(-1)[i]) < 1e-4 * max(1.0, abs(analytic.reshape(-1)[i]))


# ---------------------------------------------------------------------------- semantics the CUDA path must match
def test_fps_first_index_and_tie_rule():
    """first index 0; ties -> lowest (k mod 512, k).  Duplicates at k=5 and k=513: 513 wins
    (513 mod 512 = 1 < 5); at k=5 and k=517 (same residue): 5 wins."""
    rng = np.random.default_rng(3)
    base = rng.random((1, 1024, 3), dtype=np.float32) * 0.1
    far = np.array([50.0, 50.0, 50.0], np.float32)
    a = base.copy()
    a[0, 5] = far
    a[0, 513] = far
    assert ops.farthest_point_sample(4, a)[0, :2].tolist() == [0, 513]
    b = base.copy()
    b[0, 5] = far
    b[0, 517] = far
    assert ops.farthest_point_sample(4, b)[0, :2].tolist() == [0, 5]


def test_fps_matches_plain_argmax_without_ties():
    rng = np.random.default_rng(4)
    p = rng.standard_normal((2, 3000, 3)).astype(np.float32)
    got = ops.farthest_point_sample(64, p)
    for i in range(2):
        td = np.full(3000, 1e38, np.float32)
        old, exp = 0, [0]
        for _ in range(63):
            d = p[i] - p[i, old]
            dd = np.float32(d[:, 1] * d[:, 1])
            dd = (d[:, 0].astype(np.float64) * d[:, 0] + dd).astype(np.float32)  # fma(dx,dx,dy*dy)
            dd = (d[:, 2].astype(np.float64) * d[:, 2] + dd).astype(np.float32)  # fma(dz,dz,.)
            td = np.minimum(td, dd)
            old = int(np.argmax(td))
            exp.append(old)
        assert got[i].tolist() == exp
    assert len(set(got[0].tolist())) == 64


def test_ball_query_pad_and_count():
    xyz1 = np.zeros((1, 10, 3), np.float32)
    xyz1[0, :, 0] = np.arange(10)
    xyz2 = np.array([[[3.0, 0, 0]]], np.float32)
    idx, cnt = ops.query_ball_point(1.5, 8, xyz1, xyz2)
    assert cnt[0, 0] == 3 and idx[0, 0].tolist() == [2, 3, 4, 2, 2, 2, 2, 2]  # ascending order, padded with first hit
    idx, cnt = ops.query_ball_point(100.0, 4, xyz1, xyz2)
    assert cnt[0, 0] == 4 and idx[0, 0].tolist() == [0, 1, 2, 3]  # first nsample in index order, strict '<' radius
    idx, cnt = ops.query_ball_point(1.0, 4, xyz1, xyz2)
    assert cnt[0, 0] == 1 and idx[0, 0].tolist() == [3, 3, 3, 3]  # d == radius is NOT a hit


def test_ball_query_empty_ball_carry_quirk():
    """nearest_d / nearest_k carry across the centres j = t, t+256, ... of one reference thread."""
    rng = np.random.default_rng(9)
    n, m = 300, 600
    xyz1 = rng.random((1, n, 3), dtype=np.float32)
    xyz2 = rng.random((1, m, 3), dtype=np.float32)
    xyz2[0, 256:] += 500.0  # centres 256.. are empty; their fallback is whatever thread j%256 carried
    xyz2[0, 7] = xyz1[0, 123]  # centre 7 coincides with point 123: pins nearest_d = 1e-20 for thread 7
    idx, cnt = ops.query_ball_point(0.05, 8, xyz1, xyz2)
    assert (cnt[0, 256:] == 0).all()
    assert (idx[0, 263] == 123).all() and (idx[0, 519] == 123).all()  # stale carry, not the true nearest
    # thread 44 (300 % 256): first strict minimum over centre 44's examined prefix followed by centre 300's full scan
    def dist(j, ks):
        d = xyz2[0, j] - xyz1[0, ks]
        s = np.float32(d[:, 1] * d[:, 1])
        s = (d[:, 0].astype(np.float64) * d[:, 0] + s).astype(np.float32)
        s = (d[:, 2].astype(np.float64) * d[:, 2] + s).astype(np.float32)
        return np.maximum(np.sqrt(s), np.float32(1e-20))
    kexit = int(idx[0, 44, 7]) if cnt[0, 44] == 8 else n - 1
    seq_d = np.concatenate([dist(44, np.arange(kexit + 1)), dist(300, np.arange(n))])
    seq_k = np.concatenate([np.arange(kexit + 1), np.arange(n)])
    assert idx[0, 300, 0] == seq_k[int(np.argmin(seq_d))]
    # a centre with no earlier hit-free history and nothing nearer: gets its true nearest
    xyz2b = xyz2[:, :200].copy()
    xyz2b[0, 100] += 500.0
    idxb, cntb = ops.query_ball_point(0.05, 8, xyz1, xyz2b)
    db = np.sqrt(((xyz1[0] - xyz2b[0, 100]) ** 2).sum(1))
    assert cntb[0, 100] == 0 and idxb[0, 100, 0] == int(np.argmin(db))


def test_knn_point_matches_bruteforce():
    rng = np.random.default_rng(11)
    xyz1 = rng.random((2, 200, 3), dtype=np.float32)
    xyz2 = rng.random((2, 17, 3), dtype=np.float32)
    val, idx = ops.knn_point(5, xyz1, xyz2)
    d = ops.knn_dist(xyz1, xyz2)
    order = np.argsort(d, axis=2, kind="stable")[:, :, :5]
    assert np.array_equal(np.sort(idx, axis=2), np.sort(order, axis=2))
    assert np.allclose(val, np.take_along_axis(d, order, 2))


def test_cumsum_and_prob_sample():
    rng = np.random.default_rng(12)
    w = rng.random((3, 20000), dtype=np.float32)
    c = ops.cumsum(w)
    assert np.allclose(c, np.cumsum(w.astype(np.float64), 1), rtol=2e-6)
    r = rng.random((3, 500), dtype=np.float32)
    out = ops.prob_sample(w, r)
    q = r * c[:, -1:]
    for i in range(3):
        exp = np.searchsorted(c[i], q[i], side="left")
        assert np.array_equal(out[i], np.minimum(exp, 19999))


# ---------------------------------------------------------------------------- reference GPU kernels' committed outputs
@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_gpu_ops.npz")), reason="GPU golden not generated yet")
def test_oracle_matches_reference_cuda_kernels_golden():
    """tests/golden/ref_gpu_ops.npz holds outputs of the REFERENCE CUDA kernels (built unmodified for sm_100a) run on a
    B200 by tests/golden/make_golden_gpu.py on seeded inputs regenerated here.  This pins the CPU oracle -- including the
    FMA association, the FPS tie rule and the ball-query carry -- to the reference itself."""
    import importlib

    gen = importlib.import_module("tests.golden.make_golden_gpu")
    g = np.load(os.path.join(GOLD, "ref_gpu_ops.npz"))
    for name, case in gen.cases().items():
        xyz2 = case.get("xyz2")
        if name + "/fps" in g:
            fps = ops.farthest_point_sample(case["m"], case["xyz1"])
            assert np.array_equal(fps, g[name + "/fps"]), name
            xyz2 = ops.gather_point(case["xyz1"], fps)
        if name + "/bq_idx" in g:
            idx, cnt = ops.query_ball_point(case["radius"], case["nsample"], case["xyz1"], xyz2)
            assert np.array_equal(cnt, g[name + "/bq_cnt"]), name
            assert np.array_equal(idx, g[name + "/bq_idx"]), name
        if name + "/topk_idx" in g:
            outi, out = ops.select_top_k(case["k"], case["dist"])
            assert np.array_equal(outi[:, :, :case["k"]], g[name + "/topk_idx"]), name


# ---------------------------------------------------------------------------- C ABI surface
def test_c_abi_exports_every_declared_symbol(f3d_lib):
    hdr = open(os.path.join(os.path.dirname(HERE), "include", "feat3dnet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(f3d_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    import importlib

    sigs = importlib.import_module("3dfeatnet_b200._lib").SIGNATURES
    for n in names:
        assert hasattr(f3d_lib, n), "library does not export %s" % n
        assert n in sigs, "host binding misses %s" % n
    assert sorted(sigs) == names
    assert f3d_lib.f3d_version() >= 100
    assert f3d_lib.f3d_packed_weights_floats(32) > 100000

"""GPU parity tests (-m gpu) of the sampling / grouping operators, called through the C ABI (ctypes) via the
reference-shaped Python wrappers.  Checkers: the CPU oracle (oracle/ops_oracle.c) and, when oracle/_ref is present,
the reference's own CUDA kernels built unmodified for sm_100a.  Indices are compared BIT-EXACTLY."""
import numpy as np
import pytest
import torch

from oracle import ops as oops
from oracle import ref as oref
from tests.conftest import pkg

pytestmark = pytest.mark.gpu

HAVE_REF = oref.available("libref_grouping.so") and oref.available("libref_sampling.so")


def T(a, dev):
    return torch.as_tensor(np.ascontiguousarray(a)).to(dev)


def clouds(kind, b, n, seed):
    synth = pkg("synth")
    if kind == "uniform":
        return synth.uniform_cloud(b, n, seed)
    if kind == "oxford":
        return synth.make_batch(b, n, seed0=seed)
    if kind == "kitti":
        return synth.make_batch(b, n, seed0=seed, kind="kitti")
    rng = np.random.default_rng(seed)
    if kind == "dups":  # 10 % exact duplicates (datagenerator.py:148-157 pads clouds this way)
        x = synth.uniform_cloud(b, n, seed)
        k = max(1, n // 10)
        for i in range(b):
            x[i, rng.choice(n, k, replace=False)] = x[i, rng.choice(n, k, replace=False)]
        return x
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------------ FPS / gather
@pytest.mark.parametrize("kind,b,n,m", [
    ("uniform", 2, 4096, 512), ("oxford", 2, 16384, 512), ("dups", 3, 5000, 300), ("uniform", 1, 777, 64),
    ("dups", 2, 8192, 256), ("uniform", 33, 1024, 32), ("uniform", 1, 40000, 128), ("uniform", 2, 100, 100),
    ("dups", 2, 20000, 300), ("dups", 1, 100000, 200), ("oxford", 3, 131072, 96), ("uniform", 1, 16385, 40),
    ("uniform", 1, 140000, 20),
])
def test_fps_bit_exact_vs_oracle(cuda, kind, b, n, m):
    ts = pkg("tf_ops.sampling.tf_sampling")
    x = clouds(kind, b, n, 100 + n)
    got = ts.farthest_point_sample(m, T(x, cuda)).cpu().numpy()
    want = oops.farthest_point_sample(m, x)
    assert got.dtype == np.int32 and np.array_equal(got, want)
    kp = ts.gather_point(T(x, cuda), T(want, cuda)).cpu().numpy()
    assert np.array_equal(kp, oops.gather_point(x, want))


@pytest.mark.parametrize("case,n,m", [("more_samples_than_points", 50, 80), ("single_point", 1, 4), ("same_xy", 3000, 200),
                                      ("outlier", 16384, 128), ("boundary", 2048, 64), ("boundary", 2049, 64), ("boundary", 4096, 64),
                                      ("boundary", 8193, 64), ("boundary", 16384, 64), ("cluster_ragged", 49999, 96),
                                      ("two_values", 6000, 300)])
def test_fps_group_kernel_edge_cases(cuda, case, n, m):
    """shapes that stress the binned / group-culled kernel: degenerate grid extents, empty groups, padded tails, the
    2048 / 4096 / 8192 / 16384 template boundaries, a ragged last chunk of a CTA cluster, massive exact ties"""
    ts = pkg("tf_ops.sampling.tf_sampling")
    rng = np.random.default_rng(n * 7 + m)
    x = rng.uniform(-30, 30, (2, n, 3)).astype(np.float32)
    if case == "same_xy":       # one xy cell, only z differs
        x[:, :, 0] = 1.5
        x[:, :, 1] = -2.25
    elif case == "outlier":     # one far point stretches the grid: almost everything lands in one cell
        x[:, 17] = (1.0e6, -1.0e6, 3.0)
    elif case == "two_values":  # every point is one of two locations: every round is a many-way tie
        x[:, ::2] = x[:, :1]
        x[:, 1::2] = x[:, 1:2]
    got = ts.farthest_point_sample(m, T(x, cuda)).cpu().numpy()
    assert np.array_equal(got, oops.farthest_point_sample(m, x))


@pytest.mark.parametrize("b,n,m", [(3, 5000, 300), (2, 16384, 512), (2, 40000, 128), (1, 140000, 20)])
def test_fps_with_fused_gather_equals_the_two_ops(cuda, b, n, m):
    """f3d_farthest_point_sample_gather (sample_points in one launch): same indices, new_xyz == gather_point(inp, idx) bit for bit"""
    ts = pkg("tf_ops.sampling.tf_sampling")
    lib_mod = pkg("_lib")
    x = T(clouds("dups", b, n, 5 + n), cuda)
    idx_ref = ts.farthest_point_sample(m, x)
    kp_ref = ts.gather_point(x, idx_ref)
    idx = torch.full((b, m), -1, dtype=torch.int32, device=cuda)
    kp = torch.full((b, m, 3), float("nan"), device=cuda)
    temp = torch.empty((b, n), device=cuda) if n > 131072 else None
    lib_mod.check(lib_mod.lib().f3d_farthest_point_sample_gather(b, n, m, lib_mod.ptr(x), lib_mod.ptr(temp), lib_mod.ptr(idx), lib_mod.ptr(kp),
                                                                 lib_mod.stream()), "fps+gather")
    assert torch.equal(idx, idx_ref) and torch.equal(kp, kp_ref)


def test_fps_all_points_identical(cuda):
    """every distance ties at 0: the reference tie rule picks k=0 each round"""
    ts = pkg("tf_ops.sampling.tf_sampling")
    x = np.ones((1, 2000, 3), np.float32)
    got = ts.farthest_point_sample(5, T(x, cuda)).cpu().numpy()
    assert np.array_equal(got, oops.farthest_point_sample(5, x))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,b,n,m", [("oxford", 4, 16384, 512), ("dups", 2, 4096, 512), ("uniform", 2, 131072, 64)])
def test_fps_bit_exact_vs_reference_cuda(cuda, kind, b, n, m):
    ts = pkg("tf_ops.sampling.tf_sampling")
    x = T(clouds(kind, b, n, 7), cuda)
    assert torch.equal(ts.farthest_point_sample(m, x), oref.gpu_farthest_point_sample(m, x))


def test_gather_point_grad_deterministic_and_exact(cuda):
    ts = pkg("tf_ops.sampling.tf_sampling")
    rng = np.random.default_rng(1)
    b, n, m = 3, 500, 4000  # m > n: many repeated indices
    inp = rng.random((b, n, 3), dtype=np.float32)
    idx = rng.integers(0, n, (b, m)).astype(np.int32)
    g = rng.standard_normal((b, m, 3)).astype(np.float32)
    want = oops.gather_point_grad(inp, idx, g)
    got = ts.gather_point_grad(n, T(idx, cuda), T(g, cuda))
    assert np.array_equal(got.cpu().numpy(), want)  # same ascending-j summation order as the oracle: bit-exact
    assert torch.equal(got, ts.gather_point_grad(n, T(idx, cuda), T(g, cuda)))
    x = T(inp, cuda).requires_grad_(True)
    ts.gather_point(x, T(idx, cuda)).backward(T(g, cuda))
    assert np.array_equal(x.grad.cpu().numpy(), want)


# ------------------------------------------------------------------------------------------------ ball query
def _centres(x, m, mode, seed):
    rng = np.random.default_rng(seed)
    b, n, _ = x.shape
    if mode == "subset":
        idx = np.stack([rng.choice(n, m, replace=(m > n)) for _ in range(b)]).astype(np.int32)
        return oops.gather_point(x, idx)
    if mode == "external":  # not in the cloud; a third of them far away (empty balls -> carried fallback)
        c = x[:, rng.integers(0, n, m)] + rng.normal(0, 0.5, (b, m, 3)).astype(np.float32)
        c[:, m // 3::3] += 1000.0
        c[:, 1] = x[:, 3]
        return np.ascontiguousarray(c, np.float32)
    raise ValueError(mode)


@pytest.mark.parametrize("kind,b,n,m,radius,ns,mode", [
    ("uniform", 2, 4096, 512, 2.0, 64, "subset"), ("oxford", 2, 16384, 512, 2.0, 64, "subset"),
    ("uniform", 2, 8192, 256, 2.0, 64, "external"), ("dups", 2, 4096, 700, 1.5, 32, "external"),
    ("uniform", 1, 1237, 101, 3.0, 64, "subset"), ("uniform", 3, 333, 1030, 4.0, 16, "external"),
    ("oxford", 1, 32768, 1024, 2.0, 64, "subset"), ("uniform", 1, 65536, 2048, 2.0, 64, "subset"),
    ("uniform", 1, 131072, 4096, 2.0, 64, "subset"), ("uniform", 2, 600, 70, 0.05, 8, "external"),
    ("uniform", 1, 50, 9, 100.0, 64, "subset"),
    # >= 4096 centres per cloud: the grid query walks the centres in spatially binned order (same rows, written at the centre's own index)
    ("oxford", 2, 16384, 8192, 2.0, 64, "subset"), ("uniform", 1, 20000, 6000, 2.0, 32, "external"),
    # >= 32768 points: the cloud is binned per index window and a centre stops at nsample hits (dense KITTI-shape neighbourhoods: thousands of
    # candidates per centre; sparse ones take all windows in one pass); ragged n, empty balls, nsample 16 / 128, duplicates
    ("kitti", 1, 131072, 6000, 2.0, 64, "subset"), ("kitti", 2, 50001, 5000, 2.0, 64, "external"),
    ("uniform", 1, 40000, 300, 6.0, 128, "subset"), ("dups", 1, 33000, 4200, 3.0, 16, "subset"),
    ("kitti", 1, 262144, 500, 1.0, 64, "external"),
    # 18 / 21 index windows (n / 32 rounded up to a power of two: 4096 / 8192 points each), 2 clouds of 32 windows of 1024 points
    ("kitti", 1, 70000, 5000, 2.0, 64, "subset"), ("uniform", 1, 170000, 3000, 2.5, 64, "external"), ("kitti", 2, 32768, 4100, 1.5, 48, "subset"),
])
def test_ball_query_bit_exact_vs_oracle(cuda, kind, b, n, m, radius, ns, mode):
    tg = pkg("tf_ops.grouping.tf_grouping")
    x = clouds(kind, b, n, 5 + n)
    c = _centres(x, m, mode, n + m)
    widx, wcnt = oops.query_ball_point(radius, ns, x, c)
    for use_grid in (True, False):  # grid-accelerated kernel and plain scan: both bit-exact
        idx, cnt = tg.query_ball_point(radius, ns, T(x, cuda), T(c, cuda), use_grid=use_grid)
        assert idx.dtype == torch.int32 and cnt.dtype == torch.int32
        assert np.array_equal(cnt.cpu().numpy(), wcnt), use_grid
        assert np.array_equal(idx.cpu().numpy(), widx), use_grid
    if mode == "external":
        assert (wcnt == 0).any()  # the fallback path was exercised


@pytest.mark.parametrize("kind,n,start,m,ns", [("kitti", 131072, 0, 30000, 64), ("kitti", 131072, 30000, 30000, 64),
                                               ("kitti", 131072, 120000, 11072, 64), ("dups", 40000, 4097, 5000, 16),
                                               ("uniform", 33000, 28000, 5000, 32), ("kitti", 262144, 250000, 8000, 64)])
def test_ball_query_centres_that_are_a_slice_of_the_cloud(cuda, kind, n, start, m, ns):
    """inference.py:118-131 scores every point of a scan, MAX_POINTS centres at a time: the centres are points start .. start + m - 1 of the
    cloud, handed over as a view of the same memory -- the windowed kernel then takes them in the order of the cloud's own binning (no
    second binning), rows still at the centre's own index; chunk borders fall inside index windows"""
    tg = pkg("tf_ops.grouping.tf_grouping")
    x = clouds(kind, 1, n, 11 + n)
    radius = 2.0
    widx, wcnt = oops.query_ball_point(radius, ns, x, np.ascontiguousarray(x[:, start:start + m]))
    xd = T(x, cuda)
    view = xd[:, start:start + m, :]
    assert view.is_contiguous() and view.data_ptr() == xd.data_ptr() + 12 * start
    idx, cnt = tg.query_ball_point(radius, ns, xd, view)
    assert np.array_equal(cnt.cpu().numpy(), wcnt) and np.array_equal(idx.cpu().numpy(), widx)
    grid = tg.BallGrid(radius, xd, max_centres=m)   # the file flow's form: one binning, a query per chunk
    idx2, cnt2 = tg.query_ball_point(radius, ns, xd, view, grid=grid)
    assert torch.equal(idx2, idx) and torch.equal(cnt2, cnt)
    idx3, cnt3 = tg.query_ball_point(radius, ns, xd, view.clone(), grid=grid)  # a copy of the centres: the binned-centres path, same rows
    assert torch.equal(idx3, idx) and torch.equal(cnt3, cnt)


def test_ball_query_radius_on_representable_boundaries(cuda):
    """points at exactly d == r (not a hit: strict '<'), one ulp inside, one ulp outside; radii 0.5, 2.0, sqrt-inexact 0.3"""
    tg = pkg("tf_ops.grouping.tf_grouping")
    for r in (0.5, 2.0, 0.3, 1e-3, 37.25):
        r32 = np.float32(r)
        xs = np.array([r32, np.nextafter(r32, np.float32(0)), np.nextafter(r32, np.float32(100)), 0.0,
                       np.nextafter(np.nextafter(r32, np.float32(0)), np.float32(0))], np.float32)
        x = np.zeros((1, 5 * 3, 3), np.float32)
        for a in range(3):
            x[0, a * 5:(a + 1) * 5, a] = xs
        c = np.zeros((1, 1, 3), np.float32)
        widx, wcnt = oops.query_ball_point(float(r32), 16, x, c)
        for use_grid in (True, False):
            idx, cnt = tg.query_ball_point(float(r32), 16, T(x, cuda), T(c, cuda), use_grid=use_grid)
            assert np.array_equal(cnt.cpu().numpy(), wcnt) and np.array_equal(idx.cpu().numpy(), widx), (r, use_grid)
    # random near-boundary configurations: centres at distance r*(1 +- few ulp) in random directions
    rng = np.random.default_rng(0)
    n = 4096
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    scale = 2.0 * (1 + rng.integers(-4, 5, n) * 2.0 ** -24)
    x = (d * scale[:, None]).astype(np.float32)[None]
    c = np.zeros((1, 1, 3), np.float32)
    widx, wcnt = oops.query_ball_point(2.0, 4096, x, c)
    assert 0 < wcnt[0, 0] < n
    for use_grid in (True, False):
        idx, cnt = tg.query_ball_point(2.0, 4096, T(x, cuda), T(c, cuda), use_grid=use_grid)
        assert np.array_equal(cnt.cpu().numpy(), wcnt) and np.array_equal(idx.cpu().numpy(), widx), use_grid


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,b,n,m,mode", [("oxford", 2, 16384, 512, "subset"), ("uniform", 2, 4096, 777, "external"),
                                             ("uniform", 1, 131072, 4096, "subset")])
def test_ball_query_bit_exact_vs_reference_cuda(cuda, kind, b, n, m, mode):
    tg = pkg("tf_ops.grouping.tf_grouping")
    x = clouds(kind, b, n, 9)
    c = _centres(x, m, mode, 10)
    idx, cnt = tg.query_ball_point(2.0, 64, T(x, cuda), T(c, cuda))
    ridx, rcnt = oref.gpu_query_ball_point(2.0, 64, T(x, cuda), T(c, cuda))
    assert torch.equal(cnt, rcnt) and torch.equal(idx, ridx)


def test_query_ball_point2_vs_oracle(cuda):
    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.default_rng(2)
    x = rng.random((2, 3000, 3), dtype=np.float32)
    c = rng.random((2, 130, 3), dtype=np.float32)
    c[:, ::7] += 10.0
    radii = rng.uniform(0.05, 0.4, (2, 130)).astype(np.float32)
    idx, cnt = tg.query_ball_point2(T(radii, cuda), 32, T(x, cuda), T(c, cuda))
    widx, wcnt = oops.query_ball_point2(radii, 32, x, c)
    ne = wcnt > 0
    assert np.array_equal(cnt.cpu().numpy(), wcnt) and (~ne).any()
    assert np.array_equal(idx.cpu().numpy()[ne], widx[ne])  # empty rows are undefined in the reference


# ------------------------------------------------------------------------------------------------ group_point (+grad)
@pytest.mark.parametrize("b,n,c,m,ns", [(2, 4096, 3, 512, 64), (2, 1000, 64, 128, 32), (1, 333, 5, 17, 7), (1, 16384, 3, 2048, 64),
                                        (2, 512, 16, 128, 64),
                                        # per-cloud scatter-add path with 4 / 2 / 1 CTAs per cloud (and the radix path above: 131072 slots)
                                        (3, 16384, 3, 512, 64), (40, 8192, 3, 256, 64), (64, 16384, 3, 512, 64), (1, 5000, 7, 1000, 64)])
def test_group_point_and_grad_vs_oracle(cuda, b, n, c, m, ns):
    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.default_rng(n + c)
    pts = rng.random((b, n, c), dtype=np.float32)
    idx = rng.integers(0, n, (b, m, ns)).astype(np.int32)
    idx[:, :, ns // 2:] = idx[:, :, :1]  # padded rows: repeated indices, multiplicity matters
    out = tg.group_point(T(pts, cuda), T(idx, cuda))
    assert np.array_equal(out.cpu().numpy(), oops.group_point(pts, idx))
    g = rng.standard_normal((b, m, ns, c)).astype(np.float32)
    want = oops.group_point_grad(pts, idx, g)
    got = tg.group_point_grad(n, T(idx, cuda), T(g, cuda))
    assert np.array_equal(got.cpu().numpy(), want)  # ascending (j,k) order, no atomics: bit-exact and reproducible
    assert torch.equal(got, tg.group_point_grad(n, T(idx, cuda), T(g, cuda)))
    p = T(pts, cuda).requires_grad_(True)
    tg.group_point(p, T(idx, cuda)).backward(T(g, cuda))
    assert np.array_equal(p.grad.cpu().numpy(), want)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_group_point_grad_vs_reference_cuda_atomics(cuda):
    """the reference's atomicAdd result is order-dependent: compare within fp32 reassociation tolerance"""
    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.default_rng(3)
    b, n, c, m, ns = 2, 2048, 16, 256, 64
    pts = T(rng.random((b, n, c), dtype=np.float32), cuda)
    idx = T(rng.integers(0, n, (b, m, ns)).astype(np.int32), cuda)
    g = T(rng.standard_normal((b, m, ns, c)).astype(np.float32), cuda)
    assert torch.equal(tg.group_point(pts, idx), oref.gpu_group_point(pts, idx))
    assert torch.allclose(tg.group_point_grad(n, idx, g), oref.gpu_group_point_grad(pts, idx, g), rtol=1e-5, atol=1e-5)


def test_reference_unit_test_group_point_grad(cuda):
    """tf_grouping_op_test.py:10-27 with its shapes: points (1,128,16), xyz1 (1,128,3), xyz2 (1,8,3), r=0.3, nsample=32;
    analytic gradient vs numeric Jacobian, err < 1e-4."""
    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.RandomState(0)
    points = T(rng.random_sample((1, 128, 16)).astype("float32"), cuda).requires_grad_(True)
    xyz1 = T(rng.random_sample((1, 128, 3)).astype("float32"), cuda)
    xyz2 = T(rng.random_sample((1, 8, 3)).astype("float32"), cuda)
    idx, _ = tg.query_ball_point(0.3, 32, xyz1, xyz2)
    w = T(rng.standard_normal((1, 8, 32, 16)).astype("float32"), cuda)
    (tg.group_point(points, idx) * w).sum().backward()
    analytic = points.grad.clone()
    eps = 0.5  # group_point is linear in `points`: a large step only reduces fp32 cancellation noise
    with torch.no_grad():
        for i in rng.choice(128 * 16, 48, replace=False):
            d = torch.zeros(128 * 16, device=cuda)
            d[i] = eps
            d = d.reshape(1, 128, 16)
            num = ((tg.group_point(points + d, idx).double() * w).sum()
                   - (tg.group_point(points - d, idx).double() * w).sum()) / (2 * eps)
            assert abs(num.item() - analytic.reshape(-1)[i].item()) < 1e-4 * max(1.0, abs(num.item()))


def test_reference_unit_test_query_ball_point2(cuda):
    """tf_grouping_op_test.py:32-65 with its shapes and its cdist property check."""
    from scipy.spatial.distance import cdist

    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.RandomState(0)
    xyz1 = rng.random_sample((1, 128, 3)).astype("float32")
    xyz2 = rng.random_sample((1, 8, 3)).astype("float32")
    radii = rng.uniform(low=0.2, high=0.4, size=(1, 8)).astype("float32")
    idx, pts_cnt = tg.query_ball_point2(T(radii, cuda), 32, T(xyz1, cuda), T(xyz2, cuda))
    idx, pts_cnt = idx.cpu().numpy(), pts_cnt.cpu().numpy()
    assert idx.max() < 128 and pts_cnt.max() <= 32
    Y = cdist(xyz1[0].astype(np.float64), xyz2[0].astype(np.float64))
    within = Y < radii[0][None, :]
    assert np.array_equal(pts_cnt[0], within.sum(0))
    for j in range(8):
        assert set(idx[0, j]) == set(np.nonzero(within[:, j])[0])


# ------------------------------------------------------------------------------------------------ top-k / kNN
def test_select_top_k_known_answer_and_ties(cuda):
    tg = pkg("tf_ops.grouping.tf_grouping")
    dist = (10 - np.arange(16, dtype=np.float32)).reshape(2, 2, 4)  # test/selection_sort.cpp:65-93
    outi, out = tg.select_top_k(3, T(dist, cuda))
    assert np.array_equal(outi.cpu().numpy().ravel(), np.tile([3, 2, 1, 0], 4))
    outi, _ = tg.select_top_k(4, T(np.array([[[2, 2, 1, 2, 1, 3]]], np.float32), cuda))
    assert outi[0, 0, :4].tolist() == [2, 4, 0, 3]  # unstable-sort tie order, SURVEY.md A-6
    rng = np.random.default_rng(4)
    d = rng.integers(0, 9, (2, 50, 300)).astype(np.float32)
    outi, out = tg.select_top_k(64, T(d, cuda))
    wi, wo = oops.select_top_k(64, d)
    assert np.array_equal(outi.cpu().numpy(), wi) and np.array_equal(out.cpu().numpy(), wo)  # whole (b,m,n) arrays
    if HAVE_REF:
        ri, ro = oref.gpu_select_top_k(64, T(d, cuda))
        assert torch.equal(outi, ri) and torch.equal(out, ro)


@pytest.mark.parametrize("b,n,m,c,k", [(2, 512, 128, 3, 64), (1, 1000, 33, 3, 16), (2, 300, 20, 5, 300)])
def test_knn_point_vs_oracle(cuda, b, n, m, c, k):
    tg = pkg("tf_ops.grouping.tf_grouping")
    rng = np.random.default_rng(k)
    x1 = rng.random((b, n, c), dtype=np.float32)
    x1[:, n // 2:] = x1[:, :n - n // 2]  # duplicates: exact distance ties
    x2 = rng.random((b, m, c), dtype=np.float32)
    val, idx = tg.knn_point(k, T(x1, cuda), T(x2, cuda))
    wval, widx = oops.knn_point(k, x1, x2)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(val.cpu().numpy(), wval)


# ------------------------------------------------------------------------------------------------ C2 sweep (BASELINE.json configs[1])
@pytest.mark.parametrize("n", [4096, 8192, 16384, 32768, 65536, 131072])
def test_c2_sweep_knn_topk_group_point(cuda, n):
    """configs[1]: uniform clouds U(-30,30)^2 x U(-2,15) of 4k .. 128k points (10 % exact duplicates => distance ties), nsample 64:
    knn_point / select_top_k (tf_grouping.py:63-88, tf_grouping_g.cu:137-177) and group_point + its gradient (:94-132) against the
    oracle bit for bit, and against the reference's own CUDA kernels built as-is (selection sort and group exactly, the atomics of
    the gradient to reassociation tolerance).  Ball query and FPS have their own sweeps above."""
    tg = pkg("tf_ops.grouping.tf_grouping")
    k = 64
    x1 = clouds("dups", 1, n, 900 + n)
    rng = np.random.default_rng(n)
    m_knn = 64                                            # centres for kNN: the (m, n) distance matrix is what scales with n
    x2 = np.ascontiguousarray(x1[:, rng.choice(n, m_knn, replace=False)] + rng.normal(0, 0.3, (1, m_knn, 3)).astype(np.float32))
    val, idx = tg.knn_point(k, T(x1, cuda), T(x2, cuda))
    wval, widx = oops.knn_point(k, x1, x2)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(val.cpu().numpy(), wval)
    # the op alone on a distance matrix with many exact ties (quantised distances): whole (b,m,n) outputs
    d = np.floor(rng.random((1, 16, n), dtype=np.float32) * 200.0).astype(np.float32)
    outi, out = tg.select_top_k(k, T(d, cuda))
    wi, wo = oops.select_top_k(k, d)
    assert np.array_equal(outi.cpu().numpy(), wi) and np.array_equal(out.cpu().numpy(), wo)
    if HAVE_REF:
        ri, ro = oref.gpu_select_top_k(k, T(d, cuda))
        assert torch.equal(outi, ri) and torch.equal(out, ro)
    # group_point (+ grad) on M = n / 32 clusters x 64 samples taken from the kNN-shaped index range, with padded (repeated) rows
    m = n // 32
    gidx = rng.integers(0, n, (1, m, k)).astype(np.int32)
    gidx[:, :, k // 2:] = gidx[:, :, :1]
    for c in (3, 16):
        pts = rng.random((1, n, c), dtype=np.float32)
        got = tg.group_point(T(pts, cuda), T(gidx, cuda))
        assert np.array_equal(got.cpu().numpy(), oops.group_point(pts, gidx))
        g = rng.standard_normal((1, m, k, c)).astype(np.float32)
        grad = tg.group_point_grad(n, T(gidx, cuda), T(g, cuda))
        assert np.array_equal(grad.cpu().numpy(), oops.group_point_grad(pts, gidx, g))
        if HAVE_REF:
            assert torch.equal(got, oref.gpu_group_point(T(pts, cuda), T(gidx, cuda)))
            assert torch.allclose(grad, oref.gpu_group_point_grad(T(pts, cuda), T(gidx, cuda), T(g, cuda)), rtol=1e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ layer-level composition
def test_sample_and_group_matches_oracle_composition(cuda):
    pc = pkg("models.pointnet_common")
    x = clouds("oxford", 2, 4096, 3)
    xyz = T(x, cuda)
    ori = torch.linspace(-3, 3, 2 * 64, device=cuda).reshape(2, 64)
    new_xyz, new_points, idx, grouped, ep = pc.sample_and_group(64, 2.0, 32, xyz, None, orientations=ori,
                                                                normalize_radius=True)
    fps = oops.farthest_point_sample(64, x)
    kp = oops.gather_point(x, fps)
    widx, wcnt = oops.query_ball_point(2.0, 32, x, kp)
    assert np.array_equal(new_xyz.cpu().numpy(), kp) and np.array_equal(idx.cpu().numpy(), widx)
    g = (oops.group_point(x, widx) - kp[:, :, None, :]) / np.float32(2.0)
    assert np.allclose(ep['grouped_xyz_before'].cpu().numpy(), g, rtol=1e-6, atol=1e-6)
    c, s = np.cos(ori.cpu().numpy())[:, :, None], np.sin(ori.cpu().numpy())[:, :, None]
    rot = np.stack([g[..., 0] * c - g[..., 1] * s, g[..., 0] * s + g[..., 1] * c, g[..., 2]], -1)
    assert np.allclose(grouped.cpu().numpy(), rot, rtol=1e-5, atol=1e-5)
    assert torch.equal(new_points, grouped)
    # detector-side helper: identity sampling when npoint <= 0 (pointnet_common.py:24-25)
    assert torch.equal(pc.sample_points(xyz, -1), xyz)


@pytest.mark.parametrize("radius,normalize", [(2.0, True), (1.7, True), (2.0, False)])
def test_fused_local_frames_carry_the_bits_of_the_op_chain(cuda, radius, normalize):
    """csrc/frames.cu (group_point - centre, / radius, rotation in ONE launch) against the op-by-op statement of
    pointnet_common.py:42-54 / :104-119 on torch ops: same rows bit for bit at the model's radius (both conventions of the rotation,
    with and without angles), same 'grouped_xyz_before' / 'rotation' end points, and the gradient with respect to the angles to rounding."""
    pc = pkg("models.pointnet_common")
    x = clouds("oxford", 3, 4096, 7)
    xyz = T(x, cuda)
    g = torch.Generator().manual_seed(3)
    ori = ((torch.rand(3, 96, generator=g) - 0.5) * 7.0).to(cuda)
    kp = pc.sample_points(xyz, 96)
    res = {}
    for fused in (True, False):
        pc.FUSED_FRAMES = fused
        try:
            o = ori.clone().requires_grad_(True)
            c, npts, idx, grouped, ep = pc.sample_and_group(96, radius, 64, xyz, None, keypoints=kp, orientations=o, normalize_radius=normalize)
            w = torch.randn(grouped.shape, generator=torch.Generator().manual_seed(9)).to(cuda)
            (go,) = torch.autograd.grad((grouped * w).sum(), o)
            o2 = ori.clone().requires_grad_(True)
            q, qidx = pc.query_and_group_points(xyz, None, kp, 64, radius, normalize_radius=normalize, orientations=o2)
            (go2,) = torch.autograd.grad((q * w).sum(), o2)
            plain, _ = pc.query_and_group_points(xyz, None, kp, 64, radius, normalize_radius=normalize, orientations=None)
            res[fused] = (grouped.detach(), ep['grouped_xyz_before'].detach(), ep['rotation'].detach(), go, q.detach(), go2, plain.detach(), idx, qidx)
        finally:
            pc.FUSED_FRAMES = True
    a, b = res[True], res[False]
    # the kernel divides by the radius like TensorFlow's RealDiv; torch multiplies by the rounded reciprocal of a Python scalar: the same
    # bits when the radius is a power of two (the model's 2.0), one unit in the last place apart otherwise
    exact = (not normalize) or radius == 2.0
    for i in (0, 1, 2, 4, 6, 7, 8):
        assert torch.equal(a[i], b[i]) if (exact or i in (2, 7, 8)) else torch.allclose(a[i], b[i], rtol=3e-7, atol=1e-7), i
    assert torch.equal(a[6], a[1])  # no angles: the rows before the rotation
    for i in (3, 5):
        assert torch.allclose(a[i], b[i], rtol=1e-4, atol=1e-4 * b[i].abs().max().item()), i
    # a cloud that needs a gradient takes the op-by-op statement (the fused op carries none to xyz)
    xg = xyz.clone().requires_grad_(True)
    q, _ = pc.query_and_group_points(xg, None, kp, 64, radius, normalize_radius=normalize, orientations=ori)
    (gx,) = torch.autograd.grad(q.sum(), xg)
    assert gx.abs().sum().item() > 0


# ------------------------------------------------------------------------------------------------ the reference's smoke mains
def test_reference_grouping_main_shapes(cuda):
    """tf_ops/grouping/tf_grouping.py:90-118 (`__main__`, no assertion there): np.random.seed(100), points (32,512,64),
    xyz1 (32,512,3), xyz2 (32,128,3), kNN with k = nsample = 64 then group_point -- and its radius = 0.1 branch -- with the
    oracle as the check the reference script lacks."""
    tg = pkg("tf_ops.grouping.tf_grouping")
    np.random.seed(100)
    pts = np.random.random((32, 512, 64)).astype('float32')
    tmp1 = np.random.random((32, 512, 3)).astype('float32')
    tmp2 = np.random.random((32, 128, 3)).astype('float32')
    points, xyz1, xyz2 = T(pts, cuda), T(tmp1, cuda), T(tmp2, cuda)
    val, idx = tg.knn_point(64, xyz1, xyz2)
    grouped = tg.group_point(points, idx)
    wval, widx = oops.knn_point(64, tmp1, tmp2)
    assert np.array_equal(idx.cpu().numpy(), widx) and np.array_equal(val.cpu().numpy(), wval)
    assert grouped.shape == (32, 128, 64, 64) and grouped.dtype == torch.float32
    assert np.array_equal(grouped.cpu().numpy(), oops.group_point(pts, widx))
    idx2, cnt2 = tg.query_ball_point(0.1, 64, xyz1, xyz2)          # the script's knn=False branch
    widx2, wcnt2 = oops.query_ball_point(0.1, 64, tmp1, tmp2)
    assert np.array_equal(idx2.cpu().numpy(), widx2) and np.array_equal(cnt2.cpu().numpy(), wcnt2)
    assert np.array_equal(tg.group_point(points, idx2).cpu().numpy(), oops.group_point(pts, widx2))


def test_reference_sampling_main_flow(cuda):
    """tf_ops/sampling/tf_sampling.py:60-89 (`__main__`, no assertion there): 5 random triangles (seed 100), 8192 surface
    samples drawn with prob_sample over the triangle areas + gather_point, then farthest_point_sample(1024) + gather_point.
    The element-wise arithmetic is done once in NumPy; the four operator calls are compared with the oracle bit for bit."""
    ts = pkg("tf_ops.sampling.tf_sampling")
    np.random.seed(100)
    triangles = np.random.rand(1, 5, 3, 3).astype('float32')
    tria, trib, tric = (np.ascontiguousarray(triangles[:, :, i, :]) for i in range(3))
    areas = np.sqrt((np.cross(trib - tria, tric - tria) ** 2).sum(2) + 1e-9).astype(np.float32)
    rng = np.random.default_rng(100)                                  # tf.random_uniform in the script
    randomnumbers, us, vs = (rng.random((1, 8192), dtype=np.float32) for _ in range(3))
    triids = ts.prob_sample(T(areas, cuda), T(randomnumbers, cuda))
    want_ids = oops.prob_sample(areas, randomnumbers)
    assert np.array_equal(triids.cpu().numpy(), want_ids) and set(np.unique(want_ids)) <= set(range(5))
    corners = []
    for tri in (tria, trib, tric):
        got = ts.gather_point(T(tri, cuda), triids).cpu().numpy()
        assert np.array_equal(got, oops.gather_point(tri, want_ids))
        corners.append(got)
    uplusv, uminusv = 1 - np.abs(us + vs - 1), us - vs
    us, vs = (uplusv + uminusv) * 0.5, (uplusv - uminusv) * 0.5
    pt_sample = (corners[0] + (corners[1] - corners[0]) * us[..., None] + (corners[2] - corners[0]) * vs[..., None]).astype(np.float32)
    pt = T(pt_sample, cuda)
    fps = ts.farthest_point_sample(1024, pt)
    want_fps = oops.farthest_point_sample(1024, pt_sample)
    assert np.array_equal(fps.cpu().numpy(), want_fps)
    reduced = ts.gather_point(pt, fps).cpu().numpy()
    assert reduced.shape == (1, 1024, 3) and reduced.dtype == np.float32
    assert np.array_equal(reduced, oops.gather_point(pt_sample, want_fps))

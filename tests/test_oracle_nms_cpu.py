"""CPU tests (-m "not gpu"): oracle/nms.py against the REFERENCE's own nms() (inference.py:226-261).  The committed fixture
tests/golden/ref_nms.npz holds outputs of that function, taken out of the reference file with `ast` and executed unmodified
(tests/golden/make_golden_nms.py); when /root/reference is present (build container) the same function is also run live."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import nms as onms

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _golden_cases():
    g = np.load(os.path.join(GOLD, "ref_nms.npz"))
    for name in sorted({k.split("/")[0] for k in g.files}):
        radius, ratio, kmax = g[name + "/cli"]
        yield name, g[name + "/xyz"], g[name + "/attention"], dict(
            nms_radius=float(radius), min_response_ratio=float(ratio), max_keypoints=int(kmax)), (
            g[name + "/xyz_nms"], g[name + "/attention_nms"], g[name + "/num_keypoints"].tolist())


def test_golden_fixture_covers_truncation_and_padding():
    names = {c[0]: c for c in _golden_cases()}
    assert set(names) == {"oxford6000", "dense", "sparse_padded"}
    assert names["dense"][4][2] == [64, 64] and names["oxford6000"][4][2] == [1024]
    num = names["sparse_padded"][4][2][0]
    want_xyz = names["sparse_padded"][4][0]
    assert 0 < num < 256 and (want_xyz[0, num:] == want_xyz[0, 0]).all()  # padded with the best keypoint


@pytest.mark.parametrize("fn", ["nms", "nms_bruteforce"])
def test_oracle_nms_matches_reference_golden(fn):
    """both oracle statements (on scikit-learn, and tree-free) reproduce the reference function's stored outputs bit for bit"""
    for name, xyz, att, cli, (want_xyz, want_att, want_num) in _golden_cases():
        got_xyz, got_att, got_num, got_idx = getattr(onms, fn)(xyz, att, **cli)
        assert got_num == want_num, name
        assert np.array_equal(got_xyz, want_xyz) and np.array_equal(got_att, want_att), name
        b = np.arange(xyz.shape[0])[:, None]
        assert np.array_equal(xyz[b, got_idx], want_xyz) and np.array_equal(att[b, got_idx], want_att), name


@pytest.mark.skipif(not os.path.exists("/root/reference/inference.py"), reason="the reference tree is only in the build container")
def test_oracle_nms_matches_reference_function_live():
    spec = importlib.util.spec_from_file_location("make_golden_nms", os.path.join(GOLD, "make_golden_nms.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    rng = np.random.default_rng(123)
    xyz = rng.uniform(0, 6, (2, 2500, 3)).astype(np.float32)
    att = np.log1p(np.exp(rng.standard_normal((2, 2500)) * 2)).astype(np.float32)
    for cli in (dict(nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024),
                dict(nms_radius=0.8, min_response_ratio=0.2, max_keypoints=32)):
        want_xyz, want_att, want_num = mk.reference_nms(**cli)(xyz, att.copy())
        got_xyz, got_att, got_num, _ = onms.nms(xyz, att, **cli)
        assert got_num == want_num
        assert np.array_equal(got_xyz, want_xyz) and np.array_equal(got_att, want_att)

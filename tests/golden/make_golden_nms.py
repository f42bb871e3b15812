"""Regenerates tests/golden/ref_nms.npz from the REFERENCE's own `nms` function (needs /root/reference and scikit-learn; run
in the build container, not on the GPU box).

inference.py cannot be imported -- its module level parses the command line and imports TensorFlow -- so the `nms`
FunctionDef (inference.py:226-261) is taken out of the parsed file with `ast` and executed, unmodified, in a namespace that
holds what it reads: `np`, scikit-learn's `NearestNeighbors`, and an `args` object with the CLI fields (:40-47).  No reference
source is written anywhere; only the inputs and the function's outputs are stored.

    python tests/golden/make_golden_nms.py
"""
import ast
import os
import sys
import types

import numpy as np
from sklearn.neighbors import NearestNeighbors

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("F3D_REFERENCE", "/root/reference")


def reference_nms(nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024):
    """The reference's nms(xyz, attention), bound to the given CLI values."""
    path = os.path.join(REFERENCE, "inference.py")
    tree = ast.parse(open(path).read(), filename=path)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "nms"]
    assert len(fn) == 1, "inference.py: expected exactly one top-level nms()"
    ns = dict(np=np, NearestNeighbors=NearestNeighbors,
              args=types.SimpleNamespace(nms_radius=nms_radius, min_response_ratio=min_response_ratio, max_keypoints=max_keypoints))
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    return ns["nms"]


def cases():
    """name -> (xyz (B,N,3) f32, attention (B,N) f32, CLI values).  Attention values are distinct within a cloud and no two
    neighbours of a point are equidistant in these clouds (jittered), so the result does not depend on sklearn's tie order."""
    rng = np.random.default_rng(77)
    out = {}
    ox = np.load(os.path.join(HERE, "oxford_270_xyz.npy")).astype(np.float32)[:6000]
    ox = ox + rng.normal(0, 1e-3, ox.shape).astype(np.float32)
    att = rng.permutation(ox.shape[0]).astype(np.float32) / ox.shape[0] + 0.01
    out["oxford6000"] = (ox[None], att[None], dict(nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024))
    # dense cloud: far more than 50 points inside the radius, so the 50-NN truncation decides; 2 clouds; few keypoints wanted
    dense = rng.uniform(0, 3, (2, 3000, 3)).astype(np.float32)
    datt = np.stack([rng.permutation(3000), rng.permutation(3000)]).astype(np.float32) + 1.0
    out["dense"] = (dense, datt, dict(nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=64))
    # sparse cloud with a high response threshold: fewer survivors than max_keypoints -> padded with the best
    sparse = rng.uniform(-40, 40, (1, 500, 3)).astype(np.float32)
    satt = (rng.permutation(500).astype(np.float32)[None] + 1.0) ** 3
    out["sparse_padded"] = (sparse, satt, dict(nms_radius=2.0, min_response_ratio=0.3, max_keypoints=256))
    return out


if __name__ == "__main__":
    store = {}
    for name, (xyz, att, cli) in cases().items():
        xyz_nms, att_nms, num = reference_nms(**cli)(xyz, att.copy())
        store[name + "/xyz"], store[name + "/attention"] = xyz, att
        store[name + "/cli"] = np.array([cli["nms_radius"], cli["min_response_ratio"], cli["max_keypoints"]], np.float64)
        store[name + "/xyz_nms"], store[name + "/attention_nms"] = xyz_nms, att_nms
        store[name + "/num_keypoints"] = np.array(num, np.int32)
        print(name, "keypoints", num, "of", cli["max_keypoints"])
    np.savez_compressed(os.path.join(HERE, "ref_nms.npz"), **store)
    print("wrote ref_nms.npz", os.path.getsize(os.path.join(HERE, "ref_nms.npz")), "bytes")
    sys.exit(0)

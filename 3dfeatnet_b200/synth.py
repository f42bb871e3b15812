"""Synthetic workloads of the named shapes: seeded, no dataset access.

Oxford-shape cloud: the xyz of example_data/oxford_270.bin (committed as tests/golden/oxford_270_xyz.npy, a data
fixture) under a random permutation, a random rotation about z and N(0,0.01) jitter clipped at 0.05, seed = 1000+b.
For N != 16384 the cloud is tiled with sigma=0.05 jitter or subsampled.  KITTI-shape: same from kitti_00_001554.
"""
import os

import numpy as np

_GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def base_cloud(kind="oxford"):
    name = {"oxford": "oxford_270_xyz.npy", "kitti": "kitti_00_001554_xyz.npy"}[kind]
    return np.load(os.path.join(_GOLD, name)).astype(np.float32)


def shaped_cloud(n, seed, kind="oxford", base=None):
    rng = np.random.default_rng(seed)
    base = base_cloud(kind) if base is None else base
    nb = base.shape[0]
    if n <= nb:
        pts = base[rng.permutation(nb)[:n]].copy()
    else:
        reps = (n + nb - 1) // nb
        pts = np.concatenate([base] + [base + rng.normal(0, 0.05 if kind == "oxford" else 0.1, base.shape).astype(np.float32)
                                       for _ in range(reps - 1)], axis=0)
        pts = pts[rng.permutation(pts.shape[0])[:n]]
    th = rng.uniform(0, 2 * np.pi)
    c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
    pts = pts @ R.T
    pts = pts + np.clip(rng.normal(0, 0.01, pts.shape), -0.05, 0.05).astype(np.float32)
    return np.ascontiguousarray(pts, dtype=np.float32)


def make_batch(batch, n, seed0=1000, kind="oxford"):
    base = base_cloud(kind)
    return np.stack([shaped_cloud(n, seed0 + b, kind, base) for b in range(batch)], axis=0)


def uniform_cloud(batch, n, seed):
    rng = np.random.default_rng(seed)
    lo = np.array([-30, -30, -2], np.float32)
    hi = np.array([30, 30, 15], np.float32)
    return (rng.random((batch, n, 3), dtype=np.float32) * (hi - lo) + lo).astype(np.float32)

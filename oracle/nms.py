"""CPU restatement of the keypoint non-maximum suppression of inference.py:226-261.

TEST INFRASTRUCTURE ONLY.  Pinned to the reference function: tests/golden/ref_nms.npz holds outputs of
the reference's own nms() (its FunctionDef taken out of inference.py with `ast` and executed unmodified,
tests/golden/make_golden_nms.py) and tests/test_oracle_nms_cpu.py holds both statements below to them
bit for bit.  The neighbour search is scikit-learn's NearestNeighbors(n_neighbors=50,
algorithm='ball_tree') (reference pin 0.24.2, requirements.txt:39; this image has 1.9.0 -- the one
remaining gap); the reference has no test of its own for it.  The function body follows the reference
statement by statement; args.* of the reference become keyword arguments with the CLI defaults
(inference.py:40-47: nms_radius 0.5, min_response_ratio 1e-2, max_keypoints 1024).

`nms_bruteforce` states the same rule without a tree (float64 distances, neighbours ordered by
(distance, index)); it is the definition the CUDA kernel implements and equals `nms` whenever no two
neighbours of a point are at exactly the same float64 distance (sklearn's order among exact ties is
unspecified).
"""
import numpy as np


def nms(xyz, attention, nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024, n_neighbors=50):
    from sklearn.neighbors import NearestNeighbors

    num_models = xyz.shape[0]
    num_keypoints = [0] * num_models
    xyz_nms = np.zeros((num_models, max_keypoints, 3), xyz.dtype)
    attention_nms = np.zeros((num_models, max_keypoints), xyz.dtype)
    all_indices = np.zeros((num_models, max_keypoints), np.int32)
    for i in range(num_models):
        nbrs = NearestNeighbors(n_neighbors=n_neighbors, algorithm="ball_tree").fit(xyz[i, :, :])
        distances, indices = nbrs.kneighbors(xyz[i, :, :])
        knn_attention = attention[i, indices]
        outside_ball = distances > nms_radius
        knn_attention[outside_ball] = 0.0
        is_max = np.where(np.argmax(knn_attention, axis=1) == 0)[0]
        # NumPy 1.19 (requirements.txt:29): float32 scalar * Python float -> float64; stated explicitly so NumPy 2 agrees
        attention_thresh = float(np.max(attention[i, :])) * min_response_ratio
        is_max_attention = [(attention[i, m], m) for m in is_max if float(attention[i, m]) > attention_thresh]
        is_max_attention = sorted(is_max_attention, reverse=True)
        max_indices = [m[1] for m in is_max_attention]
        if len(max_indices) >= max_keypoints:
            max_indices = max_indices[:max_keypoints]
            num_keypoints[i] = len(max_indices)
        else:
            num_keypoints[i] = len(max_indices)
            max_indices = np.pad(max_indices, (0, max_keypoints - len(max_indices)), "constant",
                                 constant_values=max_indices[0])
        xyz_nms[i, :, :] = xyz[i, max_indices, :]
        attention_nms[i, :] = attention[i, max_indices]
        all_indices[i, :] = max_indices
    return xyz_nms, attention_nms, num_keypoints, all_indices


def nms_bruteforce(xyz, attention, nms_radius=0.5, min_response_ratio=1e-2, max_keypoints=1024, n_neighbors=50):
    """Tree-free statement of the same rule (small N only: O(N^2) memory)."""
    num_models, n, _ = xyz.shape
    num_keypoints = [0] * num_models
    xyz_nms = np.zeros((num_models, max_keypoints, 3), xyz.dtype)
    attention_nms = np.zeros((num_models, max_keypoints), xyz.dtype)
    all_indices = np.zeros((num_models, max_keypoints), np.int32)
    for i in range(num_models):
        p = xyz[i].astype(np.float64)
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        d = np.sqrt(d2)
        d[np.arange(n), np.arange(n)] = -1.0  # the query point itself sorts first (sklearn returns self at position 0)
        order = np.lexsort((np.broadcast_to(np.arange(n), (n, n)), d), axis=1)[:, :n_neighbors]
        dist = np.take_along_axis(d, order, axis=1)
        dist[:, 0] = 0.0
        a = attention[i, order].copy()
        a[dist > nms_radius] = 0.0
        is_max = np.where(np.argmax(a, axis=1) == 0)[0]
        thresh = float(np.max(attention[i])) * min_response_ratio
        cand = sorted([(attention[i, m], m) for m in is_max if float(attention[i, m]) > thresh], reverse=True)
        max_indices = [m[1] for m in cand]
        num_keypoints[i] = min(len(max_indices), max_keypoints)
        if len(max_indices) >= max_keypoints:
            max_indices = max_indices[:max_keypoints]
        else:
            max_indices = np.pad(max_indices, (0, max_keypoints - len(max_indices)), "constant",
                                 constant_values=max_indices[0])
        xyz_nms[i] = xyz[i, max_indices]
        attention_nms[i] = attention[i, max_indices]
        all_indices[i] = max_indices
    return xyz_nms, attention_nms, num_keypoints, all_indices

"""train.py's validation (train.py:240-315): descriptor distance of matching / non-matching cluster pairs and the false-positive
rate at 95 % recall.  The reference stacks up to NUM_CLUSTERS validation clusters into ONE cloud, 100 m apart along x, and
describes them in a single pass with the cluster centres fed as `keypoints`; the same is done here through
Feat3dNet.get_inference_model(keypoints=...)."""
import importlib
import os

import numpy as np
import torch

_ROOT = __name__.split(".")[0]
_dg = importlib.import_module("3dfeatnet_b200.data.datagenerator" if _ROOT == "3dfeatnet_b200" else "data.datagenerator")

NUM_CLUSTERS = 512  # train.py:24


def load_validation_groundtruths(fname, proportion=1):
    """train.py:240-255: one header line, then rows whose LAST column is 1 (match) / 0 (non-match)."""
    groundtruths = []
    with open(fname) as fid:
        fid.readline()
        for i_gt, line in enumerate(l for l in fid if l.strip()):
            groundtruths.append((i_gt, int(line.split()[-1])))
    if 0 < proportion < 1:
        groundtruths = groundtruths[0::int(1.0 / proportion)]
    return groundtruths


def fp_rate_at_95_recall(positive_dist, negative_dist):
    """train.py:309-314: threshold = 95th percentile of the positive distances; FP rate among the negatives (strict <)."""
    positive_dist = np.asarray(positive_dist, dtype=np.float64)
    negative_dist = np.asarray(negative_dist, dtype=np.float64)
    if positive_dist.size == 0 or negative_dist.size == 0:
        raise ValueError("fp_rate_at_95_recall needs at least one positive and one negative pair")
    d_at_95_recall = np.percentile(positive_dist, 95)
    num_fp = int(np.count_nonzero(negative_dist < d_at_95_recall))
    return num_fp / negative_dist.size


def stack_clusters(clouds, spacing=100.0):
    """train.py:271-289: shift cluster j by j*spacing along x and concatenate; returns (1,N,cols) cloud and the (1,NUM_CLUSTERS,3)
    centres (rows beyond len(clouds) are the origin, as in the reference)."""
    shifted = []
    for j, c in enumerate(clouds):
        c = np.array(c, dtype=np.float32, copy=True)
        c[:, 0] += j * spacing
        shifted.append(c)
    offsets = np.zeros((1, NUM_CLUSTERS, 3), dtype=np.float32)
    offsets[0, :len(clouds), 0] = np.arange(len(clouds), dtype=np.float32) * spacing
    return np.concatenate(shifted, axis=0)[None], offsets


def pair_distances(model, clouds1, clouds2, device="cuda"):
    """Descriptor distance of cluster j of clouds1 vs cluster j of clouds2 (train.py:291-303), <= NUM_CLUSTERS pairs per call."""
    if len(clouds1) != len(clouds2) or not 0 < len(clouds1) <= NUM_CLUSTERS:
        raise ValueError("pair_distances expects 1..%d cluster pairs" % NUM_CLUSTERS)
    feats = []
    for clouds in (clouds1, clouds2):
        pc, offsets = stack_clusters(clouds)
        with torch.no_grad():
            _, features, _, _ = model.get_inference_model(torch.as_tensor(pc).to(device), False, keypoints=torch.as_tensor(offsets).to(device))
        feats.append(features[0, :len(clouds)].float().cpu().numpy())
    return np.sqrt(np.sum(np.square(feats[0] - feats[1]), axis=1))


def validate(model, val_folder, val_groundtruths, data_dim=6, device="cuda"):
    """train.py:260-315.  Cluster pair i lives in `<val_folder>/<i>_0.bin` and `<i>_1.bin`.  Returns the FP rate at 95 % recall
    (1 when there is nothing to validate, like the reference)."""
    if val_groundtruths is None or len(val_groundtruths) == 0:
        return 1
    positive_dist, negative_dist = [], []
    for i_test in range(0, len(val_groundtruths), NUM_CLUSTERS):
        chunk = val_groundtruths[i_test:i_test + NUM_CLUSTERS]
        clouds1 = [_dg.DataGenerator.load_point_cloud(os.path.join(val_folder, "%d_0.bin" % idx), data_dim) for idx, _ in chunk]
        clouds2 = [_dg.DataGenerator.load_point_cloud(os.path.join(val_folder, "%d_1.bin" % idx), data_dim) for idx, _ in chunk]
        d = pair_distances(model, clouds1, clouds2, device)
        positive_dist += [d[i] for i in range(len(d)) if chunk[i][1] == 1]
        negative_dist += [d[i] for i in range(len(d)) if chunk[i][1] == 0]
    return fp_rate_at_95_recall(positive_dist, negative_dist)

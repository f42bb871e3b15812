"""Triplet data generator with the reference's interface (data/datagenerator.py:9-182): metadata file
`fname | positive indices | non-negative indices`, shuffle / reset, next_triplet(k, num_points, augmentation),
20 m crop + random down-sampling (or duplication up) to num_points.

Host-side logic only (file IO and index bookkeeping are CPU work in the reference too); the augmentations it applies are the
batched device functions of 3dfeatnet_b200/augment.py, or any object with the reference's `.apply(xyz)` method.
Randomness comes from one numpy Generator per instance (seedable), not from the global `random` / `np.random` state."""
import logging
import os
from collections import deque

import numpy as np


class DataGenerator(object):
    def __init__(self, filename="", num_cols=6, seed=None):
        """filename: metadata text file; point-cloud paths inside it are relative to its folder (datagenerator.py:20-23)."""
        self.logger = logging.getLogger(self.__class__.__name__)
        self.dataset_folder = os.path.split(filename)[0]
        self.paths_and_labels = []
        self.load_metadata(filename)
        self.num_cols = num_cols
        self.size = len(self.paths_and_labels)
        self.indices = deque(range(self.size))
        self.data = [None] * self.size
        self.rng = np.random.default_rng(seed)

    def load_metadata(self, path):
        """datagenerator.py:31-39: `fname | p0 p1 ... | n0 n1 ...` per line."""
        self.paths_and_labels = []
        with open(path) as f:
            for line in f:
                if not line.strip():
                    continue
                parts = [s.strip() for s in line.split("|")]
                if len(parts) != 3:
                    raise ValueError("%s: expected 'fname | positives | nonnegatives', got %r" % (path, line))
                fname, positives, negatives = parts
                self.paths_and_labels.append((fname, set(int(s) for s in positives.split()), set(int(s) for s in negatives.split())))

    def reset(self):
        """datagenerator.py:41-45"""
        self.indices = deque(range(len(self.data)))

    def shuffle(self):
        """datagenerator.py:47-52: new random order, called at the start of each epoch."""
        self.indices = deque(int(i) for i in self.rng.permutation(len(self.data)))

    def next_triplet(self, k=1, num_points=4096, augmentation=()):
        """datagenerator.py:54-104.  Returns (anchors, positives, negatives), each (<=k, num_points, num_cols) float32, or
        (None, None, None) when the epoch is exhausted."""
        anchors, positives, negatives = [], [], []
        for _ in range(k):
            try:
                i_anchor = self.indices.popleft()
            except IndexError:
                break
            i_positive, i_negative = self.get_positive_negative(i_anchor)
            trio = [self.process_point_cloud(self.get_point_cloud(i), num_points=num_points) for i in (i_anchor, i_positive, i_negative)]
            for a in augmentation:
                for c in trio:
                    c[:, :3] = a.apply(c[:, :3])
            anchors.append(trio[0])
            positives.append(trio[1])
            negatives.append(trio[2])
        if not anchors:
            return None, None, None
        return np.stack(anchors, axis=0), np.stack(positives, axis=0), np.stack(negatives, axis=0)

    def get_point_cloud(self, i):
        """datagenerator.py:106-120"""
        if not 0 <= i < len(self.data):
            raise IndexError("point cloud index %d out of range" % i)
        return DataGenerator.load_point_cloud(os.path.join(self.dataset_folder, self.paths_and_labels[i][0]), num_cols=self.num_cols)

    def get_positive_negative(self, anchor):
        """datagenerator.py:122-142: a random positive; a random cloud that is neither positive nor non-negative."""
        _, positives, nonnegatives = self.paths_and_labels[anchor]
        if not positives:
            raise ValueError("cloud %d has no positives" % anchor)
        if len(positives | nonnegatives) >= self.size:
            raise ValueError("cloud %d has no admissible negative" % anchor)  # the reference would loop forever
        positive = sorted(positives)[int(self.rng.integers(len(positives)))]
        while True:
            negative = int(self.rng.integers(self.size))
            if negative not in positives and negative not in nonnegatives:
                return positive, negative

    def process_point_cloud(self, cloud, num_points=4096):
        """datagenerator.py:144-160: crop to a 20 m radius, then sample num_points rows without replacement (or pad with
        randomly duplicated rows when there are not enough)."""
        cloud = cloud[np.sum(np.square(cloud[:, :3]), axis=1) <= 20 * 20, :]
        if cloud.shape[0] == 0:
            raise ValueError("no point within 20 m of the origin")
        if cloud.shape[0] <= num_points:
            self.logger.warning("Only %i out of %i required points in raw point cloud. Duplicating...", cloud.shape[0], num_points)
            pad = cloud[self.rng.integers(cloud.shape[0], size=num_points - cloud.shape[0]), :]
            return np.concatenate((cloud, pad), axis=0)
        return cloud[self.rng.choice(cloud.shape[0], size=num_points, replace=False), :]

    @staticmethod
    def load_point_cloud(path, num_cols=6):
        """datagenerator.py:162-182: `.bin` = raw float32 rows of num_cols values; anything else = comma-delimited text."""
        if path.endswith("bin"):
            model = np.fromfile(path, dtype=np.float32)
            if model.size % num_cols:
                raise ValueError("%s: %d floats is not a multiple of %d columns" % (path, model.size, num_cols))
            return np.reshape(model, (-1, num_cols))
        return np.loadtxt(path, dtype=np.float32, delimiter=",", ndmin=2)

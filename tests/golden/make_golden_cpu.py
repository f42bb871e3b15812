"""Regenerates the CPU-side golden fixtures from the REFERENCE's own code (needs /root/reference; run in the
build container, not on the GPU box):

  selection_sort_ref.txt   stdout of tf_ops/grouping/test/selection_sort.cpp built as-is (oracle/_ref/selection_sort_ref)
  ref_cpu_grouping.npz     outputs of the CPU loops of tf_ops/grouping/test/query_ball_point.cpp:19-84 (built as-is into
                           oracle/_ref/libref_cpu_grouping.so) on seeded inputs

    python tests/golden/make_golden_cpu.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ops, ref  # noqa: E402

ops.build()
out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "selection_sort_ref")], capture_output=True, text=True).stdout
lines = [l for l in out.splitlines() if not l.startswith("selection sort cpu time")]
open(os.path.join(HERE, "selection_sort_ref.txt"), "w").write("\n".join(lines) + "\n")

rng = np.random.default_rng(20240)
b, n, m, ns, c = 2, 256, 64, 32, 8
xyz1 = rng.random((b, n, 3), dtype=np.float32)
xyz2 = rng.random((b, m, 3), dtype=np.float32)
xyz2[:, ::2] = xyz1[:, :m // 2]  # half of the centres are cloud points: their balls are never empty
points = rng.random((b, n, c), dtype=np.float32)
grad = rng.standard_normal((b, m, ns, c)).astype(np.float32)
radius = np.float32(0.2)
idx = ref.cpu_query_ball_point(float(radius), ns, xyz1, xyz2, fill=-7)
idx_valid = np.where(idx < 0, 0, idx)
np.savez_compressed(os.path.join(HERE, "ref_cpu_grouping.npz"), xyz1=xyz1, xyz2=xyz2, points=points, grad=grad,
                    radius=radius, nsample=ns, idx=idx, grouped=ref.cpu_group_point(points, idx_valid),
                    grad_points=ref.cpu_group_point_grad(points, idx_valid, grad))
print("wrote", os.listdir(HERE))

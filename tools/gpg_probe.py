"""Probe: group_point_grad at the C3 shape (for ncu launch lists)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
tg = importlib.import_module("3dfeatnet_b200.tf_ops.grouping.tf_grouping"); ts = importlib.import_module("3dfeatnet_b200.tf_ops.sampling.tf_sampling")
synth = importlib.import_module("3dfeatnet_b200.synth")
B, N, M, S = 64, 16384, 512, 64
xyz = torch.as_tensor(synth.make_batch(B, N)).cuda()
kp = ts.gather_point(xyz, ts.farthest_point_sample(M, xyz))
idx, _ = tg.query_ball_point(2.0, S, xyz, kp)
g = torch.randn((B, M, S, 3), device="cuda")
for _ in range(3):
    out = tg.group_point_grad(N, idx, g)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))

"""The on-device NMS alone on one KITTI-shape scan (131 072 points, random softplus attention) -- the command the ncu capture of the
nms_* kernels runs (tools/gpu_r02_bn.sh); prints the average time of a call."""
import importlib, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
synth = importlib.import_module("3dfeatnet_b200.synth"); inf = importlib.import_module("3dfeatnet_b200.inference")
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
xyz = torch.as_tensor(synth.make_batch(1, n, seed0=5, kind="kitti")[:, :, :3].copy()).to(dev)
att = torch.as_tensor(np.log1p(np.exp(np.random.default_rng(1).standard_normal((1, n)).astype(np.float32) * 2))).to(dev)
inf.nms(xyz, att); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(calls):
    out = inf.nms(xyz, att)
e.record(); torch.cuda.synchronize()
print("nms of %d points: %.1f us per call, %d keypoints" % (n, s.elapsed_time(e) / calls * 1e3, int(out[2][0])))

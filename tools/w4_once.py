"""One KITTI-shape scan (131 072 points) through the W4 flow (attention at every point, NMS, descriptors), bf16x3, a few calls -- the
command the ncu captures of the flow's kernels run (tools/gpu_r02_bo.sh)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
synth = importlib.import_module("3dfeatnet_b200.synth"); inf = importlib.import_module("3dfeatnet_b200.inference")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
dev = torch.device("cuda:0")
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
pc = torch.as_tensor(synth.make_batch(1, 131072, seed0=5, kind="kitti")).to(dev)
net = f3.Feat3dNet({'num_clusters': 1024}, device=dev, seed=0, precision="bf16x3")
for _ in range(calls):
    out = inf.detect_and_describe(net, pc)
torch.cuda.synchronize()
print("ok", int(out[3][0]))

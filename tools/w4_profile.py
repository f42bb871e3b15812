import importlib, os, sys, torch
sys.path.insert(0, ".")
synth = importlib.import_module("3dfeatnet_b200.synth"); inf = importlib.import_module("3dfeatnet_b200.inference")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
dev = torch.device("cuda:0")
pc = torch.as_tensor(synth.make_batch(1, 131072, seed0=5, kind="kitti")).to(dev)
net = f3.Feat3dNet({'num_clusters': 1024}, device=dev, seed=0, precision="bf16x3")
for _ in range(3): inf.detect_and_describe(net, pc)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    inf.detect_and_describe(net, pc); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=60))

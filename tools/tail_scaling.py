"""Measurement aid: kernel times of the detector / descriptor forward against the number of clusters (b clouds x 512), to separate the
fixed part of the two per-cluster tails (launch, tensor-memory allocation, weight staging) from their per-round part (one round = one
64-cluster tile per CTA).  Prints one JSON line per batch size; CUDA-event brackets of the library's own kernel timer, best of 5."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
lib = importlib.import_module("3dfeatnet_b200._lib"); synth = importlib.import_module("3dfeatnet_b200.synth")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet"); pc = importlib.import_module("3dfeatnet_b200.models.pointnet_common")
dev = torch.device("cuda:0")
N, M = 4096, 512
net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0, precision="bf16x3")
packed = net.packed_weights()
for b in (1, 9, 18, 37, 55, 64, 74, 111, 148):
    xyz = torch.as_tensor(synth.make_batch(b, N, seed0=3)).to(dev)[:, :, :3].contiguous()
    kp = pc.sample_points(xyz, M)
    idx, _ = pc.query_ball_point(2.0, 64, xyz, kp)
    best = {}
    for rep in range(6):
        lib.lib().f3d_debug_kernel_timer(1)
        att, ori = f3.detector_forward_fused(xyz, kp, idx, 2.0, packed, "bf16x3")
        f3.descriptor_forward_fused(xyz, kp, idx, ori, 2.0, packed, 32, "bf16x3")
        torch.cuda.synchronize()
        t = lib.kernel_timings()
        lib.lib().f3d_debug_kernel_timer(0)
        if rep == 0: continue
        for name, ms, _ in t:
            best[name] = min(best.get(name, 1e9), ms)
    print(json.dumps(dict(clouds=b, tiles64=b * M // 64, rounds_148=-(-(b * M // 64) // 148), us={k: round(v * 1e3, 2) for k, v in best.items()})))

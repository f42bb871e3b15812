"""Bring-up tool: FPS time per call over the workload shapes (C3 / C4 / KITTI / duplicates / ragged), L2 flushed."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
ts = importlib.import_module("3dfeatnet_b200.tf_ops.sampling.tf_sampling"); synth = importlib.import_module("3dfeatnet_b200.synth")
variants = sys.argv[1:] or ["16"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, B, N, M in (("C3", 64, 16384, 512), ("C4", 18, 4096, 512), ("8k", 32, 8192, 512), ("C5", 2, 131072, 1024), ("dup", 4, 16384, 512), ("ragged", 3, 10000, 300)):
    xyz = synth.make_batch(B, N, seed0=11)
    if name == "dup":
        xyz[:, 1::3] = xyz[:, 0:-1:3][:, : xyz[:, 1::3].shape[1]]  # a third of the points duplicated
    x = torch.as_tensor(xyz).cuda()
    ref = None
    for v in variants:
        os.environ["F3D_FPS_W"] = v
        out = ts.farthest_point_sample(M, x); torch.cuda.synchronize()
        t = []
        for _ in range(5):
            flush.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = ts.farthest_point_sample(M, x); e.record(); torch.cuda.synchronize()
            t.append(s.elapsed_time(e))
        if ref is None:
            ref = out.clone()
        print("%-6s B=%d N=%d M=%d variant=%-4s %.3f ms  equal_to_default=%s" % (name, B, N, M, v or "old", min(t), bool(torch.equal(out, ref))), flush=True)

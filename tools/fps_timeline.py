"""Bring-up tool: per-round clock64() breakdown of the FPS kernel (CTA 0)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
ts = importlib.import_module("3dfeatnet_b200.tf_ops.sampling.tf_sampling"); synth = importlib.import_module("3dfeatnet_b200.synth")
lib = importlib.import_module("3dfeatnet_b200._lib")
for N in (16384, 4096):
    xyz = torch.as_tensor(synth.make_batch(4, N)).cuda()
    ts.farthest_point_sample(512, xyz); torch.cuda.synchronize()
    buf = torch.zeros(256 * 32 * 4, dtype=torch.int64, device="cuda")
    lib.lib().f3d_debug_set_fps_timeline(lib.ptr(buf))
    ts.farthest_point_sample(512, xyz); torch.cuda.synchronize()
    lib.lib().f3d_debug_set_fps_timeline(None)
    a = buf.cpu().numpy().reshape(256, 32, 4)[32:250]
    start, pre_bar, post_bar, done = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    work = pre_bar - start                       # per warp: box test (+ update when active)
    active = work > np.percentile(work, 50) * 2  # crude: active warps take much longer
    print("N=%d: round period %.0f cycles | slowest-warp work %.0f (median warp %.0f) | barrier exit - slowest arrival %.0f | post-barrier reduce %.0f | active warps/round %.1f"
          % (N, np.diff(start[:, 0]).mean(), work.max(1).mean(), np.median(work), (post_bar.min(1) - pre_bar.max(1)).mean(),
             (done - post_bar).mean(), active.sum(1).mean()))

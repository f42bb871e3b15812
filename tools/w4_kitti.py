"""W4: the inference.py-faithful flow on a KITTI-shape cloud -- attention at EVERY point (M = N, centres in
chunks of 30 000), on-device NMS, descriptors at the <= 1024 surviving keypoints -- vs the north-star form (FPS-selected 1024
clusters).  Prints one JSON line per cloud size.  Reference: inference.py:99-180 (runs this per file on one GPU + sklearn CPU NMS)."""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
synth = importlib.import_module("3dfeatnet_b200.synth"); inf = importlib.import_module("3dfeatnet_b200.inference")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
dev = torch.device("cuda:0")
for n in (29291, 131072):
    pc = torch.as_tensor(synth.make_batch(1, n, seed0=5, kind="kitti")).to(dev)
    for precision in ("bf16x3", "fp32"):
        net = f3.Feat3dNet({'num_clusters': 1024}, device=dev, seed=0, precision=precision)
        for _ in range(2):
            out = inf.detect_and_describe(net, pc)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 5
        s.record()
        for _ in range(K):
            xyz_nms, feat, att, num = inf.detect_and_describe(net, pc)
        e.record(); torch.cuda.synchronize()
        ms_full = s.elapsed_time(e) / K
        s.record()
        for _ in range(K):
            net.get_inference_model(pc, False)
        e.record(); torch.cuda.synchronize()
        ms_ns = s.elapsed_time(e) / K
        att_all = torch.rand(1, n, device=dev)
        inf.nms(pc[:, :, :3].contiguous(), att_all); torch.cuda.synchronize()
        s.record()
        for _ in range(K):
            inf.nms(pc[:, :, :3].contiguous(), att_all)
        e.record(); torch.cuda.synchronize()
        ms_nms = s.elapsed_time(e) / K
        print(json.dumps(dict(workload="W4 KITTI-shape", n_points=n, precision=precision, ms_attention_everywhere_nms_describe=ms_full,
                              points_scored_per_s=n / ms_full * 1e3, num_keypoints=int(num[0]), ms_north_star_fps1024=ms_ns, ms_nms_alone=ms_nms,
                              finite=bool(torch.isfinite(feat).all()))))

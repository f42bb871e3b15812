#!/usr/bin/env python
"""The baselines BASELINE.md section 3 asks for, measured on the GPU box next to the product (measurement tool, not product code):

  UNFUSED-GPU : the detector + descriptor graph as one torch op per TF op (matmul, BN, ReLU, amax; fp32, TF32 off), every intermediate
                in HBM -- `feature_detection_module` / `feature_extraction_module` of models/feat3dnet.py in eval mode, which is what the
                reference's TF graph does on a GPU.  Index ops by the product's kernels, so the difference is the network alone.
  REF-CUDA    : FPS + ball query + group_point by the reference's own kernels built as-is (oracle/_ref), same clouds.
  KD-TREE     : SciPy cKDTree.query_ball_point -> sort -> first 64 -> pad, one cloud, host (timing data point only: float64 predicate).
  product     : the fused step (serial, eager launches) on the same batch.

    python tools/baselines.py [--clouds 64]  ->  one JSON line, also gpurun_out/baselines.json"""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402  (reported baseline only)

f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
ts = importlib.import_module("3dfeatnet_b200.tf_ops.sampling.tf_sampling")
tg = importlib.import_module("3dfeatnet_b200.tf_ops.grouping.tf_grouping")
synth = importlib.import_module("3dfeatnet_b200.synth")


def timeit(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(0)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(e))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clouds", type=int, default=64)
    ap.add_argument("--chunk", type=int, default=16, help="clouds per unfused forward (the (B,512,64,256) activations: 2.1 GB at 64)")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda:0")
    B, N, M, S, r = args.clouds, 16384, 512, 64, 2.0
    xyz = torch.as_tensor(synth.make_batch(B, N, seed0=1000, kind="oxford")).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    model = f3.Feat3dNet({'num_clusters': M}, device=dev, precision="bf16x3", seed=0)
    res = dict(workload="C3: %d Oxford-shape clouds x %d points, %d clusters x %d" % (B, N, M, S), clouds=B)

    # ---- product, fused (serial eager step through the public model call)
    def fused():
        with torch.no_grad():
            return model.get_inference_model(xyz, False)

    res["product_fused_ms"] = timeit(fused, 10, flush)
    kp, feat, att, ep = fused()

    # ---- UNFUSED-GPU: op-by-op torch graph on the same keypoints (index ops = product kernels, timed apart)
    W = model.weights

    def unfused_net():
        outs = []
        with torch.no_grad():
            for lo in range(0, B, args.chunk):
                x = xyz[lo:lo + args.chunk]
                k, idx, a, o, e = f3.feature_detection_module(x, None, M, r, False, [64, 128, 256], [128, 64], num_samples=S, params=W,
                                                               keypoints=kp[lo:lo + args.chunk])
                _, f, _ = f3.feature_extraction_module(x, None, False, [32, 64], [128], [32], keypoints=k, orientations=o, radius=r,
                                                       num_samples=S, params=W, neighbours=(idx, e['pts_cnt']))
                outs.append((a, o, f))
        return outs

    res["unfused_gpu_network_ms"] = timeit(unfused_net, 5, flush)
    outs = unfused_net()
    f_un = torch.cat([o[2] for o in outs])
    res["unfused_vs_fused_max_abs_descriptor_diff"] = float((f_un - feat).abs().max())
    res["index_ops_product_ms"] = timeit(lambda: tg.query_ball_point(r, S, xyz, ts.gather_point(xyz, ts.farthest_point_sample(M, xyz))), 5, flush)
    res["unfused_gpu_total_ms"] = res["unfused_gpu_network_ms"] + res["index_ops_product_ms"]

    # ---- REF-CUDA index ops (the reference's kernels built unmodified for sm_100a)
    if oref.available("libref_grouping.so"):
        def ref_ops():
            i = oref.gpu_farthest_point_sample(M, xyz)
            k = ts.gather_point(xyz, i)
            ridx, _ = oref.gpu_query_ball_point(r, S, xyz, k)
            return oref.gpu_group_point(xyz, ridx)

        res["ref_cuda_fps_ballquery_group_ms"] = timeit(ref_ops, 3, flush)
        res["ref_cuda_ops_plus_unfused_network_ms"] = res["ref_cuda_fps_ballquery_group_ms"] + res["unfused_gpu_network_ms"]

    # ---- KD-tree data point (host, one cloud)
    try:
        from scipy.spatial import cKDTree

        x0, k0 = xyz[0].cpu().numpy().astype(np.float64), kp[0].cpu().numpy().astype(np.float64)
        t0 = time.perf_counter()
        tree = cKDTree(x0)
        hits = tree.query_ball_point(k0, r)
        idx = np.empty((M, S), np.int32)
        for j, h in enumerate(hits):
            h = np.sort(np.asarray(h, np.int32))[:S]
            idx[j, :len(h)] = h
            idx[j, len(h):] = h[0] if len(h) else 0
        res["kdtree_ball_query_one_cloud_ms"] = (time.perf_counter() - t0) * 1e3
        mine = ep['idx'][0].cpu().numpy()
        res["kdtree_rows_equal_to_product"] = float((idx == mine).all(axis=1).mean())
    except ImportError:
        pass
    for k in ("product_fused_ms", "unfused_gpu_total_ms", "ref_cuda_ops_plus_unfused_network_ms"):
        if k in res:
            res[k.replace("_ms", "_keypoints_per_s")] = B * M / res[k] * 1e3
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "baselines.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

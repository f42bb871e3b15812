"""Per-operator timing on the GPU box: our kernels vs the reference's own CUDA kernels built as-is for sm_100a
(oracle/_ref), same inputs, CUDA events, L2 flushed between repetitions.  Writes gpurun_out/op_timing.json.
Measurement tool only (uses oracle/_ref as the reported baseline, never as the product)."""
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref as oref  # noqa: E402

ts = importlib.import_module("3dfeatnet_b200.tf_ops.sampling.tf_sampling")
tg = importlib.import_module("3dfeatnet_b200.tf_ops.grouping.tf_grouping")
synth = importlib.import_module("3dfeatnet_b200.synth")

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
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
    return min(ms)


def main():
    have_ref = oref.available("libref_grouping.so")
    res = {}
    for name, B, N, M in [("C1", 1, 16384, 512), ("C3", 64, 16384, 512), ("C4", 18, 4096, 512), ("C5", 1, 131072, 1024)]:
        kind = "kitti" if name == "C5" else "oxford"
        xyz = torch.as_tensor(synth.make_batch(B, N, seed0=1000, kind=kind)).to(dev)
        S = 64
        r = {}
        fps = ts.farthest_point_sample(M, xyz)
        kp = ts.gather_point(xyz, fps)
        idx, cnt = tg.query_ball_point(2.0, S, xyz, kp)
        r["fps_ms"] = timeit(lambda: ts.farthest_point_sample(M, xyz))
        r["ball_query_ms"] = timeit(lambda: tg.query_ball_point(2.0, S, xyz, kp))
        r["group_point_ms"] = timeit(lambda: tg.group_point(xyz, idx))
        g = torch.randn((B, M, S, 3), device=dev)
        r["group_point_grad_ms"] = timeit(lambda: tg.group_point_grad(N, idx, g))
        # algorithmic bytes
        r["fps_GBps"] = B * (12 * N + 4 * M) / r["fps_ms"] / 1e6
        r["ball_query_GBps"] = B * (12 * N + 12 * M + 4 * M * S + 4 * M) / r["ball_query_ms"] / 1e6
        r["group_point_GBps"] = B * M * S * 28 / r["group_point_ms"] / 1e6
        r["group_point_grad_GBps"] = (B * M * S * 16 + 12 * B * N) / r["group_point_grad_ms"] / 1e6
        if have_ref:
            r["ref_fps_ms"] = timeit(lambda: oref.gpu_farthest_point_sample(M, xyz), 3)
            r["ref_ball_query_ms"] = timeit(lambda: oref.gpu_query_ball_point(2.0, S, xyz, kp), 3)
            r["ref_group_point_ms"] = timeit(lambda: oref.gpu_group_point(xyz, idx), 3)
            r["ref_group_point_grad_ms"] = timeit(lambda: oref.gpu_group_point_grad(xyz, idx, g), 3)
            assert torch.equal(fps, oref.gpu_farthest_point_sample(M, xyz))
            ridx, rcnt = oref.gpu_query_ball_point(2.0, S, xyz, kp)
            assert torch.equal(idx, ridx) and torch.equal(cnt, rcnt)
        if name in ("C1", "C4"):  # a7: select_top_k / knn_point (unused by the model; (b,m,n) distance matrix)
            k = 32
            kb = min(B, 4)
            x1, x2 = xyz[:kb].contiguous(), kp[:kb].contiguous()
            dist = ((x2[:, :, None, :] - x1[:, None, :, :]) ** 2).sum(-1).contiguous()
            r["select_top_k_ms"] = timeit(lambda: tg.select_top_k(k, dist), 3)
            r["knn_point_ms"] = timeit(lambda: tg.knn_point(k, x1, x2), 3)
            r["gather_point_grad_ms"] = timeit(lambda: ts.gather_point_grad(N, fps, kp), 3)
            if have_ref:
                r["ref_select_top_k_ms"] = timeit(lambda: oref.gpu_select_top_k(k, dist), 2)
                r["ref_gather_point_grad_ms"] = timeit(lambda: oref.gpu_gather_point_grad(xyz, fps, kp), 3)
        res[name] = r
        print(name, json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "op_timing.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Parity of the shipped precision at the shipped configurations, measured against the fp64 oracle (GPU session script).

    python tools/parity_c3.py [--config C3|C4|C5] [--out gpurun_out/parity_c3.json]

C3 = the bench batch itself (64 Oxford-shape clouds x 16384 points, 512 clusters x 64, seed-0 TF random-init weights and a
randomised-BN set); C4 = the training batch shape (18 clouds x 4096 points) in eval mode; C5 = one KITTI-shape scan
(131072 points, 1024 clusters).  For each precision ("bf16x3", "fp32") it reports the error distribution of attention (relative
to the cloud's maximum), orientation (wrapped, rad) and descriptors (absolute, unit-norm vectors), and -- to attribute the
descriptor error -- the descriptor evaluated by the ORACLE at the device path's orientation (what is left is the descriptor
MLP's own error) next to the end-to-end figure (which includes the rotation by a slightly different angle).
"""
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
from oracle import net as onet  # noqa: E402

synth = importlib.import_module("3dfeatnet_b200.synth")
pipe_mod = importlib.import_module("3dfeatnet_b200.pipeline")

CONFIGS = {"C3": dict(B=64, N=16384, M=512, kind="oxford", seed0=1000),
           "C4": dict(B=18, N=4096, M=512, kind="oxford", seed0=2000),
           "C5": dict(B=1, N=131072, M=1024, kind="kitti", seed0=3000)}


def wrap(d):
    return torch.atan2(torch.sin(d), torch.cos(d))


def quantiles(x):
    x = x.reshape(-1).double()
    qs = torch.quantile(x, torch.tensor([0.5, 0.99, 0.995, 0.999, 0.9999], dtype=torch.float64))
    return dict(max=x.max().item(), p50=qs[0].item(), p99=qs[1].item(), p995=qs[2].item(), p999=qs[3].item(), p9999=qs[4].item())


def oracle_chunks(xyz, P64, M, chunk):
    outs = []
    for c0 in range(0, len(xyz), chunk):
        outs.append(onet.inference_model(xyz[c0:c0 + chunk], P64, num_clusters=M, dtype=torch.float64))
    keys = ("attention", "orientation", "features")
    ref = {k: torch.cat([o[k] for o in outs], 0) for k in keys}
    for k in ("xyz", "idx", "pts_cnt", "fps_idx"):
        ref[k] = np.concatenate([o[k] for o in outs], 0)
    return ref


def oracle_descriptor_at(xyz, P64, kp, ori, chunk):
    outs = []
    for c0 in range(0, len(xyz), chunk):
        outs.append(onet.descriptor(xyz[c0:c0 + chunk], P64, kp[c0:c0 + chunk], ori[c0:c0 + chunk].double(), 2.0, 64, dtype=torch.float64)["features"])
    return torch.cat(outs, 0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--out", default=None)
    ap.add_argument("--clouds", type=int, default=0, help="use only the first k clouds (0 = all)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    B = args.clouds or cfg["B"]
    N, M = cfg["N"], cfg["M"]
    torch.set_num_threads(os.cpu_count() or 1)
    xyz = synth.make_batch(B, N, seed0=cfg["seed0"], kind=cfg["kind"])
    chunk = max(1, min(4, (1 << 21) // (M * 64 * 4) or 1))
    report = dict(config=args.config, B=B, N=N, M=M, cases={})
    for rb in (False, True):
        params = onet.init_params(seed=0, randomize_bn=rb)
        P64 = onet.to_torch(params, torch.float64)
        t0 = time.time()
        ref = oracle_chunks(xyz, P64, M, chunk)
        t_oracle = time.time() - t0
        for precision in ("bf16x3", "fp32"):
            pipe = pipe_mod.DetectDescribePipeline(B, N, weights=params, num_clusters=M, precision=precision, device="cuda:0")
            out = {k: v.clone() for k, v in pipe.run(torch.as_tensor(xyz).cuda()).items()}
            torch.cuda.synchronize()
            assert np.array_equal(out["fps_idx"].cpu().numpy(), ref["fps_idx"]), "FPS indices differ"
            assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"]), "ball-query indices differ"
            att, ori, feat = out["attention"].cpu().double(), out["orientation"].cpu().double(), out["features"].cpu().double()
            e_att = (att - ref["attention"]).abs() / ref["attention"].abs().amax(dim=1, keepdim=True).clamp_min(1e-30)
            e_att_self = (att - ref["attention"]).abs() / ref["attention"].abs().clamp_min(1e-30)
            e_ori = wrap(ori - ref["orientation"]).abs()
            e_feat = (feat - ref["features"]).abs()
            # the descriptor MLP alone: oracle descriptor evaluated at the DEVICE path's orientation
            feat_at = oracle_descriptor_at(xyz, P64, ref["xyz"], out["orientation"].cpu(), chunk)
            e_feat_mlp = (feat - feat_at).abs()
            report["cases"]["%s/rb=%d" % (precision, int(rb))] = dict(
                attention_rel_to_cloud_max=quantiles(e_att), attention_rel_to_self=quantiles(e_att_self),
                orientation_rad=quantiles(e_ori), descriptor_abs=quantiles(e_feat.amax(dim=2)),
                descriptor_abs_given_device_orientation=quantiles(e_feat_mlp.amax(dim=2)),
                descriptor_l2=quantiles((feat - ref["features"]).norm(dim=2)),
                oracle_seconds=t_oracle)
            print(precision, "rb=%d" % rb, json.dumps(report["cases"]["%s/rb=%d" % (precision, int(rb))]), flush=True)
            del pipe
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()

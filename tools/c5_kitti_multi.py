#!/usr/bin/env python
"""BASELINE.json configs[4] (C5) on N GPUs: KITTI-shape scans (131 072 points each, synthetic) through the reference's file flow
-- inference.py:99-177: attention at EVERY point (centres in chunks of 30 000), NMS, descriptors at <= 1024 keypoints, one
[xyz | descriptor] .bin per scan -- with the REAL model, the scan list sharded across ranks (no collective).

    python tools/c5_kitti_multi.py --scans 16                                  (1 GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/c5_kitti_multi.py --scans 64

Rank 0 then re-describes a few scans of OTHER ranks' shards in a fresh single-process pass and compares the files byte for byte.
Prints one JSON line (scans/s = all scans / max-over-ranks device+host time of the sharded loop)."""
import argparse
import filecmp
import importlib
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
dist = importlib.import_module("3dfeatnet_b200.dist")
synth = importlib.import_module("3dfeatnet_b200.synth")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet")
inf = importlib.import_module("3dfeatnet_b200.inference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=16)
    ap.add_argument("--points", type=int, default=131072)
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--check", type=int, default=4, help="scans rank 0 re-describes alone for the byte comparison")
    args = ap.parse_args()
    rank, local_rank, world = dist.init("nccl")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    base = os.path.join(tempfile.gettempdir(), "f3d_c5_%s" % os.environ.get("MASTER_PORT", "0"))
    data_dir, out_dir, chk_dir = (os.path.join(base, d) for d in ("scans", "out", "check"))
    if rank == 0:
        shutil.rmtree(base, ignore_errors=True)
        os.makedirs(data_dir)
        for i in range(args.scans):
            xyz = synth.shaped_cloud(args.points, 5000 + i, kind="kitti")
            np.concatenate([xyz, np.zeros_like(xyz)], axis=1).astype(np.float32).tofile(os.path.join(data_dir, "%06d.bin" % i))
    dist.barrier()
    # the reference builds the inference model with num_clusters = -1 (every point a centre until NMS) and Attention = True
    model = f3.Feat3dNet({'num_clusters': -1, 'Attention': True}, device=dev, precision=args.precision, seed=0)
    lo, hi = dist.shard_range(args.scans, rank, world)
    inf.compute_descriptors(model, data_dir, os.path.join(base, "warm%d" % rank), device=dev, rank=0, world=max(args.scans, 1))  # one scan: warm-up
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    done = inf.compute_descriptors(model, data_dir, out_dir, device=dev, rank=rank, world=world)
    torch.cuda.synchronize()
    sec = dist.max_over_ranks(time.perf_counter() - t0, dev)
    dist.barrier()
    if rank == 0:
        files = sorted(os.listdir(out_dir))
        assert len(files) == args.scans, "expected %d output files, found %d" % (args.scans, len(files))
        # byte comparison: scans spread over the shards, re-described by this single process
        pick = sorted({int(round(k * (args.scans - 1) / max(1, args.check - 1))) for k in range(args.check)})
        os.makedirs(chk_dir, exist_ok=True)
        model1 = f3.Feat3dNet({'num_clusters': -1, 'Attention': True}, device=dev, precision=args.precision, seed=0)
        for i in pick:
            name = "%06d.bin" % i
            inf.compute_descriptors_for_file(model1, os.path.join(data_dir, name), os.path.join(chk_dir, name), seed=i, device=dev)
            assert filecmp.cmp(os.path.join(chk_dir, name), os.path.join(out_dir, name), shallow=False), "scan %s differs between the sharded and the single-process run" % name
        rows = [os.path.getsize(os.path.join(out_dir, f)) // (4 * 35) for f in files]
        print(json.dumps(dict(workload="C5: KITTI-shape scans, %d points, attention at every point -> NMS -> <= 1024 keypoints -> descriptors, file flow of inference.py"
                                       % args.points, n_gpus=world, scans=args.scans, seconds=sec, scans_per_s=args.scans / sec,
                              ms_per_scan_per_gpu=sec * 1e3 / max(1, hi - lo), keypoints_per_scan_mean=float(np.mean(rows)), precision=args.precision,
                              byte_identical_to_single_process=True, checked_scans=pick, includes="file read + H2D + two passes + NMS + D2H + file write")))
        shutil.rmtree(base, ignore_errors=True)
    dist.shutdown()


if __name__ == "__main__":
    main()

"""Per-launch CUDA-event times of ONE eager training step at C4 (BASELINE configs[3]: 6 triplets x 4096 points, 512 clusters x 64),
in launch order, with the layers' activations chained (unmaterialised) and materialised.  Measurement aid for DESIGN.md 4b.

    python tools/train_kernel_times.py [--chain 0|1|both]
"""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chain", default="both")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    pkg = lambda n: importlib.import_module("3dfeatnet_b200." + n)
    f3, layers, synth, _lib = pkg("models.feat3dnet"), pkg("models.layers"), pkg("synth"), pkg("_lib")
    dev = torch.device("cuda:0")
    L = _lib.lib()
    B, N, M = 6, 4096, 512
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(dev) for s in (1, 2, 3))
    res = {}
    for chain in ([0, 1] if args.chain == "both" else [int(args.chain)]):
        layers.CHAIN_ACTIVATIONS = bool(chain)
        net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()

        def step():
            xyz, feats, att, ep = net.get_train_model(a, p, n, True)
            loss, ep = net.get_loss(xyz, feats, att, ep)
            net.get_train_op(loss, lr=1e-5, end_points=ep)

        for _ in range(2):
            step()
        torch.cuda.synchronize()
        L.f3d_debug_kernel_timer(1)
        step()
        t = _lib.kernel_timings()
        L.f3d_debug_kernel_timer(0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(5):
            step()
        ev1.record()
        torch.cuda.synchronize()
        res["chain=%d" % chain] = dict(eager_ms_per_step=ev0.elapsed_time(ev1) / 5, sum_timed_ms=sum(ms for _, ms, _ in t),
                                       launches=[dict(kernel=k, ms=round(ms, 4), GBps=round(u / (ms * 1e6), 0) if ms > 0 else None) for k, ms, u in t])
        print("chain=%d: eager %.3f ms/step, timed kernels %.3f ms" % (chain, res["chain=%d" % chain]["eager_ms_per_step"], res["chain=%d" % chain]["sum_timed_ms"]))
        for k, ms, u in t:
            print("   %-48s %8.4f ms %7.0f GB/s" % (k, ms, u / (ms * 1e6) if ms > 0 else 0))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

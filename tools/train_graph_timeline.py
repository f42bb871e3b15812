"""Timeline of ONE replay of the captured training step at C4 (CUPTI activity records through torch.profiler): every kernel / memcpy in
start order with its duration and the idle gap before it, then the totals per kernel name.  Shows what the CUDA-event table of
bench.py (named launches only) leaves out: torch glue kernels and the gaps between graph nodes.

    python tools/train_graph_timeline.py [--out profiles/x.txt]
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
    ap.add_argument("--out", default=None)
    ap.add_argument("--chain", type=int, default=1)
    ap.add_argument("--pipelined", type=int, default=0, help="1: the software-pipelined step (next batch's sampling beside the backward pass)")
    args = ap.parse_args()
    pkg = lambda n: importlib.import_module("3dfeatnet_b200." + n)
    f3, layers, synth = pkg("models.feat3dnet"), pkg("models.layers"), pkg("synth")
    layers.CHAIN_ACTIVATIONS = bool(args.chain)
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, M = 6, 4096, 512
    a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(dev) for s in (1, 2, 3))
    net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()
    replay = net.capture_train_step(a, p, n, lr=1e-5, warmup=2, pipelined=bool(args.pipelined))
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        replay()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    lines = []
    t0 = ev[0].time_range.start
    prev_end = t0
    busy = 0.0
    agg = {}
    for e in ev:
        st, en = e.time_range.start, e.time_range.end
        gap = st - prev_end
        dur = en - st
        busy += dur
        name = e.name[:90]
        lines.append("%9.1f us  +%6.1f gap  %8.1f us  %s" % (st - t0, gap, dur, name))
        k = agg.setdefault(name, [0, 0.0])
        k[0] += 1
        k[1] += dur
        prev_end = max(prev_end, en)
    total = prev_end - t0
    out = ["one replay: %.1f us wall on the device, %.1f us in kernels, %.1f us idle, %d kernels" % (total, busy, total - busy, len(ev)), ""]
    out += ["%5d x %9.1f us  %s" % (v[0], v[1], k) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    out += ["", "timeline:"] + lines
    text = "\n".join(out)
    print(text)
    if args.out:
        open(args.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()

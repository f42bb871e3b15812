#!/usr/bin/env python
"""Tensor-pipe time per tcgen05.mma as a function of the instruction shape (GPU session script; f3d_debug_umma_bench).

    python tools/umma_bench.py [--cta-group 1|2|both] [--out gpurun_out/umma_bench.json]

All 148 SMs issue at once.  Reported per shape: cycles per instruction (median over CTAs), the math floor
M_cta * N / 256 cycles (128 x N x 16 MACs per CTA at 4096 MAC/cycle/SM), their ratio, and the wall-clock TFLOP/s of the launch."""
import argparse
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
_lib = importlib.import_module("3dfeatnet_b200._lib")


def run(L, cg, N, mn, ctas, groups=400, per_group=24):
    out = torch.zeros((ctas, 2), dtype=torch.int64, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for rep in range(3):
        ev[0].record()
        _lib.check(L.f3d_debug_umma_bench(cg, N, groups, per_group, mn, ctas, _lib.ptr(out), _lib.stream()), "umma_bench")
        ev[1].record()
        torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    o = out.cpu()
    rows = o[o[:, 1] > 0]
    cyc = (rows[:, 0].double() / rows[:, 1].double())
    floor = 128 * N / 256.0 if cg == 1 else 256 * N / 512.0
    flops = 2.0 * (128 * cg) * N * 16 * groups * per_group * (ctas // cg)
    return dict(cta_group=cg, N=N, b_mn_major=mn, ctas=ctas, cycles_per_mma=cyc.median().item(), cycles_per_mma_max=cyc.max().item(),
                floor_cycles=floor, pipe_frac=floor / cyc.median().item(), ms=ms, tflops=flops / (ms * 1e-3) / 1e12,
                implied_mhz=(rows[:, 0].double().median().item() / (ms * 1e-3)) / 1e6)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cta-group", default="1")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    L = _lib.lib()
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    res = []
    groups = ("1", "2") if args.cta_group == "both" else (args.cta_group,)
    for cg in groups:
        cg = int(cg)
        for mn in (0, 1):
            for N in ((8, 16, 32, 64, 96, 128, 192, 256) if cg == 1 else (16, 32, 64, 128, 256)):
                r = run(L, cg, N, mn, sms if cg == 1 else sms - (sms & 1))
                print(json.dumps(r), flush=True)
                res.append(r)
        # a single CTA / pair alone: the same instruction stream without chip-wide power effects
        for N in (64, 128, 256):
            r = run(L, cg, N, 1, cg)
            print(json.dumps(r), flush=True)
            res.append(r)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

"""Which phase bounds the lin_tc kernels (forward z = xW, dgrad dx = dz W^T of the training layers, csrc/train_tc.cu)?  Each shape of the C4
training step is timed with phases of the kernel switched off (f3d_debug_set_lin_tc_phases: operand conversion / MMAs / epilogue stores /
TMA fetches): what the time falls to when a phase is removed says how much of it is exposed.  Measurement aid for DESIGN.md 4b.

    python tools/lin_tc_phases.py [--out profiles/x.json]
"""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [  # (label, k, nout, nsplit)
    ("det conv1 fwd  64->128 (lin_tc_kernel<3>)", 64, 128, 3),
    ("det conv2 fwd 128->256 (pipe<3,64>)", 128, 256, 3),
    ("desc conv1 fwd 32->64 (lin_tc_kernel<3>)", 32, 64, 3),
    ("det conv2 dgrad 256->128 (pipe<2,32>)", 256, 128, 2),
    ("det conv1 dgrad 128->64 (pipe<2,64>)", 128, 64, 2),
    ("desc conv1 dgrad 64->32 (lin_tc_kernel<2>)", 64, 32, 2),
]
MASKS = [(0, "all phases"), (1, "no conversion"), (2, "no MMAs"), (4, "no stores"), (8, "no TMA"), (3, "no conversion, no MMAs"),
         (5, "no conversion, no stores"), (6, "no MMAs, no stores"), (9, "no conversion, no TMA"), (7, "TMA only"), (11, "stores only"), (15, "nothing")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--rows", type=int, default=18 * 512 * 64)
    args = ap.parse_args()
    _lib = importlib.import_module("3dfeatnet_b200._lib")
    L = _lib.lib()
    dev = torch.device("cuda:0")
    rows = args.rows
    res = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for label, k, nout, nsplit in SHAPES:
        x = torch.randn(rows, k, device=dev)
        W = torch.randn(nout, k, device=dev) * 0.1
        out = torch.empty(rows, nout, device=dev)
        part = torch.empty(4 * nout * 2 * 148 * 4, device=dev)
        wimg = torch.empty(L.f3d_debug_lin_tc_weight_bytes(k, nout), dtype=torch.uint8, device=dev)
        row = dict(shape=label, rows=rows, k=k, nout=nout, nsplit=nsplit, bytes=4.0 * rows * (k + nout), ms={})
        for mask, name in MASKS:
            _lib.check(L.f3d_debug_set_lin_tc_phases(mask), "phases")
            ts = []
            for it in range(4):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                _lib.check(L.f3d_debug_lin_tc(rows, k, nout, _lib.ptr(x), _lib.ptr(W), _lib.ptr(out), _lib.ptr(part), _lib.ptr(wimg), nsplit,
                                              _lib.stream()), "lin_tc")
                e.record()
                torch.cuda.synchronize()
                if it:
                    ts.append(s.elapsed_time(e))
            row["ms"][name] = min(ts)
        _lib.check(L.f3d_debug_set_lin_tc_phases(0), "phases")
        res.append(row)
        print("%s   (%.0f MB, HBM floor %.3f ms at 6.5 TB/s)" % (label, row["bytes"] / 1e6, row["bytes"] / 6.5e9))
        for mask, name in MASKS:
            print("      %-28s %.4f ms" % (name, row["ms"][name]))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()

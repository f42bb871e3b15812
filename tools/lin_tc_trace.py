"""clock64() trace of one CTA of lin_tc_kernel (the serial-phase forward / dgrad contraction of the training layers): cycles per phase of
a tile, averaged over the CTA's tiles, with all phases on and with phases switched off.  Measurement aid.
    python tools/lin_tc_trace.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_lib = importlib.import_module("3dfeatnet_b200._lib")
L = _lib.lib()
dev = torch.device("cuda:0")
rows = 18 * 512 * 64
NAMES = ["wait ring", "convert", "fence+sync", "issue MMAs+fetch", "wait MMAs", "tmem ld", "stores+stats", "sync"]
for label, k, nout, nsplit in [("det conv1 fwd 64->128", 64, 128, 3), ("desc conv1 fwd 32->64", 32, 64, 3)]:
    x = torch.randn(rows, k, device=dev)
    W = torch.randn(nout, k, device=dev) * 0.1
    out = torch.empty(rows, nout, device=dev)
    part = torch.empty(4 * nout * 2 * 148 * 4, device=dev)
    wimg = torch.empty(L.f3d_debug_lin_tc_weight_bytes(k, nout), dtype=torch.uint8, device=dev)
    for mask in (0, 15, 7, 11, 13, 14):
        buf = torch.zeros(2 * 64 * 16, dtype=torch.int64, device=dev)
        _lib.check(L.f3d_debug_set_lin_tc_phases(mask), "phases")
        for _ in range(2):
            buf.zero_()
            _lib.check(L.f3d_debug_lin_tc_trace(_lib.ptr(buf)), "trace")
            _lib.check(L.f3d_debug_lin_tc(rows, k, nout, _lib.ptr(x), _lib.ptr(W), _lib.ptr(out), _lib.ptr(part), _lib.ptr(wimg), nsplit, _lib.stream()), "lin_tc")
            torch.cuda.synchronize()
        _lib.check(L.f3d_debug_lin_tc_trace(None), "trace")
        t = buf.cpu().view(2, 64, 16)
        print("%s  skip mask %d" % (label, mask))
        for th, name in ((0, "thread 0"), (1, "thread 255")):
            tt = t[th]
            n = int((tt[:, 8] > 0).sum())
            if n < 3:
                continue
            d = (tt[1:n - 1, 1:9] - tt[1:n - 1, 0:8]).double()
            per_tile = (tt[2:n, 0] - tt[1:n - 1, 0]).double().mean().item()
            print("   %-10s tiles %d, %.0f cycles per tile: " % (name, n, per_tile) + ", ".join("%s %.0f" % (NAMES[i], d[:, i].mean().item()) for i in range(8)))
_lib.check(L.f3d_debug_set_lin_tc_phases(0), "phases")

"""clock64() trace of CTA 0 of wgrad_tc_kernel: cycles per phase of a 32-row stage for the issuer warp and one converter warp.
Measurement aid.   python tools/wgrad_tc_trace.py"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_lib = importlib.import_module("3dfeatnet_b200._lib")
L = _lib.lib()
dev = torch.device("cuda:0")
rows = 18 * 512 * 64
for cin, cout in ((32, 64), (64, 128), (128, 256)):
    x = torch.randn(rows, cin, device=dev)
    dz = torch.randn(rows, cout, device=dev)
    partW = torch.empty(2 * 148 * cin * cout, device=dev)
    for dbg in (0, 2):
        buf = torch.zeros(2 * 64 * 16, dtype=torch.int64, device=dev)
        for _ in range(2):
            buf.zero_()
            _lib.check(L.f3d_debug_lin_tc_trace(_lib.ptr(buf)), "trace")
            _lib.check(L.f3d_debug_wgrad_tc(rows, cin, cout, _lib.ptr(x), _lib.ptr(dz), _lib.ptr(partW), dbg, _lib.stream()), "wgrad")
            torch.cuda.synchronize()
        _lib.check(L.f3d_debug_lin_tc_trace(None), "trace")
        t = buf.cpu().view(2, 64, 16).double()
        iss, con = t[0], t[1]
        n = 60
        per = (iss[40:n, 0] - iss[39:n - 1, 0]).mean().item()
        print("wgrad %d x %d dbg %d: stage period %.0f cycles" % (cin, cout, dbg, per))
        print("   issuer   : wait image %.0f, fetch (TMA issue) %.0f, MMA issue + commit %.0f" % ((iss[20:n, 1] - iss[20:n, 0]).mean().item(),
              (iss[20:n, 3] - iss[20:n, 1]).mean().item(), (iss[20:n, 2] - iss[20:n, 3]).mean().item()))
        print("   converter: wait MMAs(s-2) %.0f, wait ring %.0f, convert %.0f, fence + arrive %.0f" % tuple(
            (con[20:n, i + 1] - con[20:n, i]).mean().item() for i in range(4)))
        # how long after the converter's arrive does the issuer see the image, and how long after the commit is the image free again
        print("   image converted -> issuer past its wait: %.0f;  issuer stage start -> converter stage start (same s): %.0f" % (
            (iss[20:n, 1] - con[20:n, 4]).mean().item(), (con[20:n, 0] - iss[20:n, 0]).mean().item()))

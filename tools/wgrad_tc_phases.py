"""wgrad_tc_kernel (dW = x^T dz on tcgen05) with phases switched off (f3d_debug_wgrad_tc: bit 0 = no TMA + no conversion, bit 1 = no MMAs):
how much of a stage is the conversion, the MMAs, the skeleton.  Measurement aid.   python tools/wgrad_tc_phases.py"""
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
    line = "wgrad %3d x %3d (%4.0f MB, floor %.3f ms):" % (cin, cout, 4e-6 * rows * (cin + cout), 4.0 * rows * (cin + cout) / 6.5e9)
    for dbg, name in ((0, "all"), (1, "no TMA/convert"), (2, "no MMAs"), (3, "skeleton"), (4, "convert w/o LDS"), (8, "convert w/o STS"), (12, "convert w/o LDS+STS"), (16, "TMA but no convert")):
        ts = []
        for it in range(4):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            _lib.check(L.f3d_debug_wgrad_tc(rows, cin, cout, _lib.ptr(x), _lib.ptr(dz), _lib.ptr(partW), dbg, _lib.stream()), "wgrad")
            e.record()
            torch.cuda.synchronize()
            if it:
                ts.append(s.elapsed_time(e))
        line += "  %s %.4f" % (name, min(ts))
    print(line)

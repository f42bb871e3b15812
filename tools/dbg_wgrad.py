import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
_lib = importlib.import_module("3dfeatnet_b200._lib")
L = _lib.lib()
dev = torch.device("cuda:0")
R = 589824
for cin, cout in ((128, 256), (128, 128), (64, 128), (32, 64)):
    x = torch.randn(R, cin, device=dev); dz = torch.randn(R, cout, device=dev)
    part = torch.empty(2 * 148 * cin * cout, device=dev)
    for dbg in (0, 1, 2, 3):
        for _ in range(2):
            _lib.check(L.f3d_debug_wgrad_tc(R, cin, cout, _lib.ptr(x), _lib.ptr(dz), _lib.ptr(part), dbg, _lib.stream()), "dbg")
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5):
            L.f3d_debug_wgrad_tc(R, cin, cout, _lib.ptr(x), _lib.ptr(dz), _lib.ptr(part), dbg, _lib.stream())
        e.record(); torch.cuda.synchronize()
        us = s.elapsed_time(e) / 5 * 1e3
        print("wgrad %d->%d dbg=%d: %.0f us  (%.2f TB/s of x+dz)" % (cin, cout, dbg, us, R * (cin + cout) * 4 / us / 1e6))

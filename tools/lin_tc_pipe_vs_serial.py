"""lin_tc: the serial-phase kernel (2-4 CTAs per SM) against the warp-specialised one (one CTA per SM, 17 warps) on the narrow layers of the C4
training step (f3d_debug_set_lin_tc_pipe_min_k).  Measured: K = 64 -> 128 channels 0.129 vs 0.119 ms, 32 -> 64 0.074 vs 0.105 ms, 64 -> 32 0.081 vs 0.094:
the launcher keeps the serial-phase kernel below K = 128.   python tools/lin_tc_pipe_vs_serial.py"""
import importlib, os, sys, torch
sys.path.insert(0, "/root/repo" if os.path.isdir("/root/repo") else ".")
_lib = importlib.import_module("3dfeatnet_b200._lib"); L = _lib.lib()
dev = torch.device("cuda:0"); rows = 18*512*64
for label, k, nout, ns in [("c1 fwd 64->128",64,128,3),("convmid-like fwd 64->128",64,128,3),("l1 fwd 32->64",32,64,3),("c1 dgrad 128->64",128,64,2),("l1 dgrad 64->32",64,32,2),("post 256->128 rows 9216",256,128,3)]:
    r = 9216 if "9216" in label else rows
    x = torch.randn(r,k,device=dev); W = torch.randn(nout,k,device=dev)*0.1; out = torch.empty(r,nout,device=dev)
    part = torch.empty(4*nout*2*148*4,device=dev); wimg = torch.empty(L.f3d_debug_lin_tc_weight_bytes(k,nout),dtype=torch.uint8,device=dev)
    res=[]
    for mink in (128, 16):
        L.f3d_debug_set_lin_tc_pipe_min_k(mink)
        ts=[]
        for it in range(5):
            s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            s.record(); _lib.check(L.f3d_debug_lin_tc(r,k,nout,_lib.ptr(x),_lib.ptr(W),_lib.ptr(out),_lib.ptr(part),_lib.ptr(wimg),ns,_lib.stream()),"x"); e.record(); torch.cuda.synchronize()
            if it: ts.append(s.elapsed_time(e))
        res.append(min(ts))
    print("%-28s serial-phase kernel %.4f ms   warp-specialised %.4f ms" % (label, res[0], res[1]))
L.f3d_debug_set_lin_tc_pipe_min_k(128)

"""Bring-up tool: dump CTA 0's clock64() timeline of the detector tensor kernel (gpurun_out/tc_timeline.txt)."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
pm = importlib.import_module("3dfeatnet_b200.pipeline"); synth = importlib.import_module("3dfeatnet_b200.synth")
lib = importlib.import_module("3dfeatnet_b200._lib")
B, N = 64, 16384
pipe = pm.DetectDescribePipeline(B, N, precision="bf16x3")
pipe.xyz.copy_(torch.as_tensor(synth.make_batch(B, N)).cuda())
pipe.step(); pipe.step(); torch.cuda.synchronize()
T = (B * 512 + 147) // 148
which = sys.argv[1] if len(sys.argv) > 1 else "det"
setter = lib.lib().f3d_debug_set_timeline if which == "det" else lib.lib().f3d_debug_set_timeline_desc
buf = torch.zeros((T + 2) * 16, dtype=torch.int64, device="cuda")
setter(lib.ptr(buf))
pipe.step(); torch.cuda.synchronize()
setter(None)
a = buf.cpu().numpy().reshape(-1, 16)[:T]
t0 = a[a > 0].min()
names = {0: "mma1_go", 1: "mma2_go", 2: "mma2_issued", 4: "P_start", 5: "P_computed", 6: "P_x1free", 8: "E_d1full", 9: "E1_computed",
         10: "E_x2free", 11: "E_x2full", 12: "E_d2full", 13: "E_done"}
with open(os.path.join(ROOT, "gpurun_out", "tc_timeline_%s.txt" % which), "w") as f:
    f.write("tile " + " ".join("%12s" % names[k] for k in sorted(names)) + "\n")
    for t in list(range(0, 12)) + list(range(100, 112)):
        f.write("%4d " % t + " ".join("%12d" % (a[t, k] - t0 if a[t, k] else -1) for k in sorted(names)) + "\n")
    rows = [t for t in range(20, min(200, T)) if a[t, 1] > 0]   # the detector issues conv2 per PAIR of tiles: even tiles carry the stamp
    per = (a[rows[-1], 1] - a[rows[0], 1]) / float(rows[-1] - rows[0])
    f.write("mean tile period (mma2_go): %.1f cycles\n" % per)
print(open(os.path.join(ROOT, "gpurun_out", "tc_timeline_%s.txt" % which)).read())

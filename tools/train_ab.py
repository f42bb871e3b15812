"""A/B of a debug switch of the library on the captured C4 training step, in ONE process on ONE box (the boxes of the pool differ by +-2 %,
more than most single changes): alternates the two settings, several rounds, serial graph.
    python tools/train_ab.py f3d_debug_set_row_walk [value_a value_b]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = lambda n: importlib.import_module("3dfeatnet_b200." + n)
f3, synth, _lib = pkg("models.feat3dnet"), pkg("synth"), pkg("_lib")
L = _lib.lib()
fn = getattr(L, sys.argv[1] if len(sys.argv) > 1 else "f3d_debug_set_row_walk")
dev = torch.device("cuda:0")
B, N, M = 6, 4096, 512
a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(dev) for s in (1, 2, 3))
va, vb = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 0)
res = {va: [], vb: []}
for rnd in range(3):
    for on in (va, vb):
        fn(on)
        net = f3.Feat3dNet({'num_clusters': M}, device=dev, seed=0).train_mode()
        replay = net.capture_train_step(a, p, n, lr=1e-5, warmup=1)
        for _ in range(5):
            replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(30):
            replay()
        e.record(); torch.cuda.synchronize()
        res[on].append(s.elapsed_time(e) / 30)
        del net, replay
        torch.cuda.empty_cache()
fn(va)
print("value %d: %s ms per step" % (va, ["%.4f" % v for v in res[va]]))
print("value %d: %s ms per step" % (vb, ["%.4f" % v for v in res[vb]]))

import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
layers = importlib.import_module("3dfeatnet_b200.models.layers")
cuda = torch.device("cuda:0")
def run(rows, cin, cout):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(rows, cin, generator=g); w = torch.randn(cin, cout, generator=g) * (2.0 / cin) ** 0.5
    b = torch.zeros(cout); ga = torch.ones(cout); be = torch.zeros(cout); gy = torch.randn(rows, cout, generator=g)
    outs = {}
    for prec in ("fp32", "bf16x3"):
        layers.TRAIN_PRECISION = prec
        ins = [t.to(cuda).requires_grad_(True) for t in (x, w, b, ga, be)]
        y, mean, var = layers.conv_bn_train(*ins, True)
        grads = torch.autograd.grad((y * gy.to(cuda)).sum(), ins)
        outs[prec] = [y.detach().cpu()] + [t.cpu() for t in grads]
    for name, a, r in zip(("y", "dx", "dW", "db", "dgamma", "dbeta"), outs["bf16x3"], outs["fp32"]):
        d = (a - r).abs()
        bad = (d > 1e-3 * (r.abs().max() + 1e-9)).nonzero()
        print(rows, cin, cout, name, "max err %.3e scale %.3e nbad %d" % (d.max(), r.abs().max(), len(bad)))
        if name in ("y", "dx", "dW"):
            print("      cos %.8f  rel-l2 %.3e" % (torch.nn.functional.cosine_similarity(a.double().flatten(), r.double().flatten(), dim=0), (a.double() - r.double()).norm() / r.double().norm()))
        if len(bad) and name in ("y", "dx") and len(bad) < 20000:
            rr = bad[:, 0]; cc = bad[:, 1]
            print("   rows%64:", sorted(set((rr % 64).tolist()))[:70], "\n   cols:", sorted(set(cc.tolist()))[:140], "\n   tiles:", sorted(set((rr // 64).tolist()))[:40])
for shp in ((589824, 128, 256), (589824, 64, 128), (589824, 3, 64), (589824, 128, 128), (9216, 256, 128)):
    run(*shp)

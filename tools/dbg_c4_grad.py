import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import net as onet
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet"); layers = importlib.import_module("3dfeatnet_b200.models.layers")
synth = importlib.import_module("3dfeatnet_b200.synth"); inf = importlib.import_module("3dfeatnet_b200.inference")
cuda = torch.device("cuda:0")
B, N, M = 6, 4096, 512
a, p, n = (torch.as_tensor(synth.make_batch(B, N, seed0=s)).to(cuda) for s in (11, 12, 13))
params = onet.init_params(seed=1, randomize_bn=True)
torch.backends.cuda.matmul.allow_tf32 = False
def run(fused, precision="bf16x3"):
    layers.FUSED_TRAINING, layers.TRAIN_PRECISION = fused, precision
    net = f3.Feat3dNet({'num_clusters': M, 'fused_loss': fused}, weights=params, device=cuda).train_mode()
    xyz, feats, att, ep = net.get_train_model(a, p, n, True)
    loss, ep = net.get_loss(xyz, feats, att, ep)
    flat = net.get_train_op(loss, lr=1e-5, end_points=ep)
    names = [k for k in net.trainable_variables()]
    sizes = [net.weights[k].numel() for k in names]
    return loss.item(), flat.detach().clone(), names, sizes
l0, g0, names, sizes = run(False)
l1, g1, _, _ = run(True, "bf16x3")
l2, g2, _, _ = run(True, "fp32")
print("loss", l0, l1, l2)
off = 0
cs = torch.nn.functional.cosine_similarity
for k, s in zip(names, sizes):
    A, Bt, C = g0[off:off+s].double(), g1[off:off+s].double(), g2[off:off+s].double()
    print("%-45s n=%6d |ref|=%.3e cos(tc,ref)=%.5f cos(f32,ref)=%.5f cos(tc,f32)=%.5f" % (k, s, A.norm(), cs(Bt, A, dim=0), cs(C, A, dim=0), cs(Bt, C, dim=0)))
    off += s
# W4 component timing at N=131072
pc = torch.as_tensor(synth.make_batch(1, 131072, seed0=5, kind="kitti")).to(cuda)
net = f3.Feat3dNet({'num_clusters': 1024}, device=cuda, seed=0, precision="bf16x3")
xyz = pc[:, :, :3].contiguous()
def t(fn, k=3):
    fn(); torch.cuda.synchronize(); t0 = time.time()
    for _ in range(k): r = fn()
    torch.cuda.synchronize(); return (time.time() - t0) / k * 1e3, r
ms, out = t(lambda: net.get_inference_model(pc, False, keypoints=xyz[:, :30000].contiguous()))
print("detector+descriptor chunk of 30000 centres: %.2f ms" % ms)
att = torch.rand(1, 131072, device=cuda)
ms, _ = t(lambda: inf.nms(xyz, att))
print("nms 131072: %.2f ms" % ms)
ms, _ = t(lambda: inf.nms(xyz[:, :29291].contiguous(), att[:, :29291].contiguous()))
print("nms 29291: %.2f ms" % ms)

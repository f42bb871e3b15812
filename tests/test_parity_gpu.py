"""GPU parity tests (-m gpu) of the SHIPPED precision at the SHIPPED configurations (BASELINE.json configs[2], [3], [4]) against
the fp64 oracle, under the one tolerance table of oracle/parity.py:

  C3  the bench batch itself: 64 Oxford-shape clouds x 16384 points (synth.make_batch(64, 16384, seed0=1000)), 512 clusters x 64,
      seed-0 weights in the TF random-init state (what bench.py times) and with randomised BN statistics;
  C4  the training batch shape (6 triplets = 18 clouds x 4096 points, 512 clusters x 64), forward in eval mode;
  C5  one KITTI-shape scan: 131 072 points, 1024 clusters.

Both device paths are held to their own row of the table: "bf16x3" (tcgen05, the default and the one the bench measures) and "fp32"
(CUDA-core FFMA).  FPS / ball-query indices must equal the oracle's bit for bit.  Matches models/feat3dnet.py:258-313.
"""
import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import parity
from tests.conftest import pkg
from tests.test_model_gpu import run_pipeline

pytestmark = pytest.mark.gpu

CONFIGS = {"C3": dict(B=64, N=16384, M=512, kind="oxford", seed0=1000),
           "C4": dict(B=18, N=4096, M=512, kind="oxford", seed0=2000),
           "C5": dict(B=1, N=131072, M=1024, kind="kitti", seed0=3000)}


@pytest.mark.parametrize("randomize_bn", [False, True])
@pytest.mark.parametrize("config", ["C3", "C4", "C5"])
def test_shipped_precisions_at_shipped_configs_vs_fp64_oracle(cuda, config, randomize_bn):
    cfg = CONFIGS[config]
    xyz = pkg("synth").make_batch(cfg["B"], cfg["N"], seed0=cfg["seed0"], kind=cfg["kind"])
    params = onet.init_params(seed=0, randomize_bn=randomize_bn)
    P64 = onet.to_torch(params, torch.float64)
    ref = parity.oracle_forward(onet, xyz, P64, cfg["M"])
    for precision in ("bf16x3", "fp32"):
        out, _ = run_pipeline(xyz, params, cfg["M"], precision=precision)
        assert np.array_equal(out["fps_idx"].cpu().numpy(), ref["fps_idx"]), "FPS indices differ from the oracle"
        assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"]), "ball-query indices differ from the oracle"
        assert np.array_equal(out["pts_cnt"].cpu().numpy(), ref["pts_cnt"])
        assert np.array_equal(out["xyz"].cpu().numpy(), ref["xyz"])
        own = parity.oracle_descriptor_at(onet, xyz, P64, ref["xyz"], out["orientation"])
        e = parity.check(parity.errors(out, ref, own), precision, "%s %s rb=%d" % (config, precision, randomize_bn))
        print(config, precision, "rb=%d" % randomize_bn, {k: "%.2e" % v for k, v in e.items()})


def test_c4_training_step_vs_fp64_oracle(cuda):
    """BASELINE configs[3] at its full size (6 triplets = 18 clouds x 4096 points, 512 clusters x 64) against the fp64 ORACLE
    (oracle/net.py train_step, itself pinned to the reference's graph files incl. the loss gradients): training-mode forward (BN batch
    statistics) under the tolerance table, loss, the 107 619-entry gradient (direction and scale: ReLU / max-pool routing flips on
    rounding-level differences, so single entries are not comparable at this size), BN shadow updates and the TF-1 Adam update.
    Matches models/feat3dnet.py:227-256,315-375 and train.py:142-158."""
    f3, synth = pkg("models.feat3dnet"), pkg("synth")
    B, N, M = 6, 4096, 512
    a, p, n = (synth.make_batch(B, N, seed0=s) for s in (2000, 2006, 2012))
    params = onet.init_params(seed=0, randomize_bn=True)
    net = f3.Feat3dNet({'num_clusters': M}, weights=params, device=cuda).train_mode()
    A, P_, Nn = (torch.as_tensor(t).to(cuda) for t in (a, p, n))
    xyz, feats, att, ep = net.get_train_model(A, P_, Nn, True)
    loss, ep = net.get_loss(xyz, feats, att, ep)
    before = {k: v.detach().clone() for k, v in net.weights.items()}
    flat = net.get_train_op(loss, lr=1e-5, end_points=ep).detach().cpu().double()
    oP = onet.to_torch(params, torch.float64, requires_grad=True)
    oloss, ograds, oout = onet.train_step(a, p, n, oP, {}, num_clusters=M, lr=1e-5, dtype=torch.float64)
    # forward in training mode: keypoints exact, float outputs under the shipped precision's row of the table
    assert np.array_equal(torch.cat(xyz).cpu().numpy(), np.asarray(oout["xyz"], dtype=np.float32))
    out = dict(attention=ep['attention'], orientation=ep['orientation'], features=torch.cat(feats))
    ref = {k: oout[k].detach() for k in ("attention", "orientation", "orientation_xy", "features")}
    e = parity.errors(out, ref)
    tol = parity.TOL["bf16x3"]
    assert e["att"] < tol["att"] and e["ori_w"] < tol["ori_w"] and e["desc_e2e_p999"] < tol["desc_e2e"], e
    assert abs(loss.item() - oloss.item()) < 1e-4 * max(1.0, abs(oloss.item()))
    names = list(net.trainable_variables())
    og = torch.cat([ograds[k].reshape(-1) for k in names]).double()
    assert flat.numel() == og.numel() == 107619
    cos = torch.nn.functional.cosine_similarity(flat, og, dim=0).item()
    scale = flat.norm().item() / og.norm().item()
    worst = 1.0
    off = 0
    for k in names:  # every variable's own gradient points the oracle's way
        cnt = ograds[k].numel()
        gk, ok = flat[off:off + cnt], og[off:off + cnt]
        off += cnt
        if ok.norm().item() > 1e-6 * og.norm().item():
            worst = min(worst, torch.nn.functional.cosine_similarity(gk, ok, dim=0).item())
    print("C4 train vs fp64 oracle: loss %.6f / %.6f, grad cos %.6f, |g|/|g_ref| %.5f, worst per-variable cos %.5f, forward %s"
          % (loss.item(), oloss.item(), cos, scale, worst, {k: "%.1e" % v for k, v in e.items() if not k.startswith("_")}))
    assert cos > 0.9999 and abs(scale - 1.0) < 2e-3 and worst > 0.999  # measured on a B200: 0.999986, 0.99953, 0.99997
    for k in ("detection/conv0/bn/moving_mean", "detection/conv2/bn/moving_variance", "description/layer1/conv_mid_0/bn/moving_variance"):
        assert torch.allclose(net.weights[k].cpu().double(), oP[k].detach(), rtol=1e-4, atol=1e-6), k
    # Adam (TF-1 form): the first step moves every trainable entry by lr * sign(g) up to eps
    moved = torch.cat([(net.weights[k] - before[k]).reshape(-1) for k in names]).cpu().double()
    omoved = torch.cat([(oP[k].detach() - torch.as_tensor(params[k], dtype=torch.float64)).reshape(-1) for k in names])
    big = og.abs() > 1e-2 * og.abs().max()
    assert big.sum().item() > 1000 and ((moved[big] - omoved[big]).abs() < 2e-7).double().mean().item() > 0.999

"""A miniature of the reference's train.py loop (train.py:93-184) on SYNTHETIC triplets, to show that the hand-written
training path learns: anchors / positives are two independently augmented, re-sampled views of the same synthetic "place",
negatives are views of another place; stage-2 model (attention + rotation), the reference's augmentations on the device,
TF-1 Adam, checkpoints written as .npz every --save-every steps and restored through checkpoint.initialize_model.
Reports the loss curve and, before / after training, how often the nearest positive descriptor of an anchor keypoint lies
within 1 m of the keypoint's true location in the positive cloud (the quantity the descriptor is trained for).

    python tools/train_synthetic.py --steps 300 --lr 1e-3
"""
import argparse, importlib, json, os, sys, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
synth = importlib.import_module("3dfeatnet_b200.synth"); aug = importlib.import_module("3dfeatnet_b200.augment")
f3 = importlib.import_module("3dfeatnet_b200.models.feat3dnet"); ck = importlib.import_module("3dfeatnet_b200.checkpoint")

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300); ap.add_argument("--lr", type=float, default=1e-3)
ap.add_argument("--places", type=int, default=48); ap.add_argument("--batch", type=int, default=6)
ap.add_argument("--points", type=int, default=4096); ap.add_argument("--clusters", type=int, default=512)
ap.add_argument("--save-every", type=int, default=100); ap.add_argument("--out", type=str, default=None)
args = ap.parse_args()
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev); gen.manual_seed(0)
rng = np.random.default_rng(0)
# "places": dense synthetic clouds; a view = random subset of `points` points + augmentation
DENSE = 4 * args.points
places = torch.as_tensor(np.stack([synth.shaped_cloud(DENSE, seed=100 + i) for i in range(args.places)])).to(dev)


def view(idx, augment=True):
    sel = torch.stack([torch.randperm(DENSE, generator=gen, device=dev)[:args.points] for _ in idx])
    xyz = torch.gather(places[idx], 1, sel[:, :, None].expand(-1, -1, 3))
    return aug.apply_augmentations(xyz, ("Jitter", "RotateSmall", "Shift"), gen=gen) if augment else xyz


def triplets():
    a = torch.as_tensor(rng.choice(args.places, args.batch, replace=False))
    n = torch.as_tensor([(int(i) + 1 + rng.integers(args.places - 1)) % args.places for i in a])
    return view(a), view(a), view(n)


@torch.no_grad()
def matching_rate(net, trials=4):
    """fraction of anchor keypoints whose nearest positive descriptor belongs to a keypoint within 1 m (views un-augmented)"""
    hits = total = 0
    for t in range(trials):
        idx = torch.arange(t * args.batch, (t + 1) * args.batch) % args.places
        xa, fa, _, _ = net.get_inference_model(view(idx, False), False)
        xp, fp, _, _ = net.get_inference_model(view(idx, False), False)
        nn = torch.cdist(fa, fp).argmin(dim=2)
        d = (xa - torch.gather(xp, 1, nn[:, :, None].expand(-1, -1, 3))).norm(dim=2)
        hits += int((d < 1.0).sum()); total += d.numel()
    return hits / total


net = f3.Feat3dNet({'num_clusters': args.clusters}, device=dev, seed=0, precision="bf16x3").train_mode()
before = matching_rate(net)
out_dir = args.out or tempfile.mkdtemp(prefix="f3d_ckpt_")
losses = []
for step in range(1, args.steps + 1):
    a, p, n = triplets()
    xyz, feats, att, ep = net.get_train_model(a, p, n, True)
    loss, ep = net.get_loss(xyz, feats, att, ep)
    net.get_train_op(loss, lr=args.lr, end_points=ep)
    losses.append(float(loss.detach()))
    if step % args.save_every == 0 or step == args.steps:
        ck.save_npz(net.weights, os.path.join(out_dir, "model-%d.npz" % step))
after = matching_rate(net)
# restore the last checkpoint into a fresh model (reference: initialize_model) and check it reproduces the trained one
fresh = f3.Feat3dNet({'num_clusters': args.clusters}, device=dev, seed=123, precision="bf16x3")
ck.initialize_model(fresh, os.path.join(out_dir, "model-%d.npz" % args.steps))
restored = matching_rate(fresh)
k = max(1, args.steps // 10)
print(json.dumps(dict(steps=args.steps, lr=args.lr, loss_first=float(np.mean(losses[:k])), loss_last=float(np.mean(losses[-k:])),
                      match_rate_before=before, match_rate_after=after, match_rate_restored_checkpoint=restored,
                      loss_curve=[round(float(np.mean(losses[i:i + k])), 5) for i in range(0, args.steps, k)], checkpoints=out_dir)))

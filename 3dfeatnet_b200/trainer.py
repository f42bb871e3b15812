"""The training loop of the reference's train.py (train():93-184) around the hand-written training step: epochs over the
triplet generator (shuffle, next_triplet, stop at the first short batch), the reference's augmentations applied on the device,
get_train_model -> get_loss -> get_train_op, a checkpoint every `checkpoint_every_n_steps`, validate() at step 1 and every
`validate_every_n_steps`; train_two_stage() is train.sh (descriptor pretraining, then the full model restored without the
`detection` scope).  TF summaries / logging configuration are out of scope; progress comes back as a history dict."""
import importlib
import os

import torch

_ROOT = __name__.split(".")[0]
_pfx = "3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else ""
_aug = importlib.import_module(_pfx + "augment")
_ck = importlib.import_module(_pfx + "checkpoint")
_val = importlib.import_module(_pfx + "validation")

BATCH_SIZE = 6  # train.py:21


def train(model, train_data, num_epochs=1, batch_size=BATCH_SIZE, num_points=4096, augmentation=("Jitter", "RotateSmall", "Shift"),
          lr=1e-5, checkpoint_dir=None, checkpoint_every_n_steps=500, val_folder=None, val_groundtruths=None,
          validate_every_n_steps=250, data_dim=6, seed=0, grad_hook=None, grad_scale=1.0, max_steps=None, on_step=None):
    """model: Feat3dNet in training mode; train_data: data.datagenerator.DataGenerator.  `augmentation` names follow
    get_augmentations_from_list (data/augment.py:4-29) and are applied per cloud on the device (augment.py); a list of the
    objects that function returns (the reference's `.apply(xyz)` interface, data/augment.py here) is applied on the device too.
    Returns {'steps', 'losses', 'fp_rates': [(step, fp_rate)], 'checkpoints': [paths]}."""
    dev = model.device if hasattr(model, "device") else torch.device("cuda")
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    history = dict(steps=0, losses=[], fp_rates=[], checkpoints=[])
    if checkpoint_dir:
        os.makedirs(checkpoint_dir, exist_ok=True)
    step = 0
    for _ in range(num_epochs):
        train_data.shuffle()  # train.py:141
        while True:
            anchors, positives, negatives = train_data.next_triplet(k=batch_size, num_points=num_points)
            if anchors is None or anchors.shape[0] != batch_size:  # train.py:148-149: the short last batch ends the epoch
                break
            clouds = [torch.as_tensor(a[:, :, :3]).to(dev, non_blocking=True) for a in (anchors, positives, negatives)]
            if augmentation and all(hasattr(a, "apply") for a in augmentation):  # objects of data/augment.py (train.py:45)
                for a in augmentation:
                    clouds = [a.apply(c) for c in clouds]
            elif augmentation:
                clouds = [_aug.apply_augmentations(c, augmentation, gen=gen) for c in clouds]
            xyz, feats, att, ep = model.get_train_model(clouds[0], clouds[1], clouds[2], True)
            loss, ep = model.get_loss(xyz, feats, att, ep)
            model.get_train_op(loss, lr=lr, end_points=ep, grad_hook=grad_hook, grad_scale=grad_scale)
            step += 1
            history["losses"].append(float(loss.detach()))
            if checkpoint_dir and step % checkpoint_every_n_steps == 0:  # train.py:158-159
                path = os.path.join(checkpoint_dir, "checkpoint.ckpt-%d.npz" % step)
                _ck.save_npz(model.weights, path)
                history["checkpoints"].append(path)
            if val_groundtruths and (step % validate_every_n_steps == 0 or step == 1):  # train.py:164
                history["fp_rates"].append((step, _val.validate(model, val_folder, val_groundtruths, data_dim, dev)))
            if on_step is not None:
                on_step(step, history)
            if max_steps is not None and step >= max_steps:
                history["steps"] = step
                return history
    history["steps"] = step
    return history


latest_checkpoint = _ck.latest_checkpoint  # tf.train.latest_checkpoint's role; lives with the checkpoint readers


def train_two_stage(make_model, train_data, log_dir, param=None, pretrain_epochs=2, num_epochs=70, **train_kwargs):
    """The reference's two-stage recipe (train.sh): (1) pretrain the descriptor alone -- `--noattention --noregress`, i.e.
    param NoRegress=True / Attention=False, augmentations Jitter RotateSmall Shift, 2 epochs, checkpoints under
    <log_dir>/pretrain/ckpt; (2) build the full model (attention + orientation), restore the pretrain checkpoint except the
    `detection` scope (`--restore_exclude detection`), add the Rotate1D augmentation and train for 70 epochs, checkpoints under
    <log_dir>/secondstage/ckpt.  `make_model(param)` returns a Feat3dNet; other keywords go to train() for both stages.
    The final state of stage 1 is always saved, so that the checkpoint stage 2 restores exists even when the run is shorter than
    checkpoint_every_n_steps.  Returns {'pretrain', 'secondstage': train() histories, 'restored': names, 'model': stage-2 model}."""
    base = dict(param or {})
    dirs = {s: os.path.join(log_dir, s, "ckpt") for s in ("pretrain", "secondstage")}
    model1 = make_model(dict(base, NoRegress=True, Attention=False))
    h1 = train(model1, train_data, num_epochs=pretrain_epochs, augmentation=("Jitter", "RotateSmall", "Shift"),
               checkpoint_dir=dirs["pretrain"], **train_kwargs)
    final = os.path.join(dirs["pretrain"], "checkpoint.ckpt-%d.npz" % h1["steps"])
    if final not in h1["checkpoints"]:
        _ck.save_npz(model1.weights, final)
        h1["checkpoints"].append(final)
    model2 = make_model(dict(base, NoRegress=False, Attention=True))
    restored = _ck.initialize_model(model2, latest_checkpoint(dirs["pretrain"]), restore_exclude=["detection"])
    h2 = train(model2, train_data, num_epochs=num_epochs, augmentation=("Jitter", "RotateSmall", "Shift", "Rotate1D"),
               checkpoint_dir=dirs["secondstage"], **train_kwargs)
    return dict(pretrain=h1, secondstage=h2, restored=restored, model=model2)

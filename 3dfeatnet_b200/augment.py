"""Training-time augmentations on the device, batched: the semantics of the reference's data/augment.py (applied there per
cloud on the CPU with NumPy, one object per type) as functions over a (B,N,3) CUDA tensor with one random draw per cloud
from an explicit torch.Generator.  Order and defaults follow get_augmentations_from_list (:4-29) and train.py:45
(Rotate1D -> Jitter -> Scale -> RotateSmall -> Shift)."""
import math

import torch


def _rand(gen, shape, device):
    return torch.rand(shape, generator=gen, device=device)


def _randn(gen, shape, device):
    return torch.randn(shape, generator=gen, device=device)


def rotate_z(xyz, gen=None):
    """RotateZ (:67-81): one angle U(0, 2 pi) per cloud, data @ [[c,s,0],[-s,c,0],[0,0,1]]  ->  x' = x c - y s, y' = x s + y c."""
    a = _rand(gen, (xyz.shape[0], 1), xyz.device) * (2 * math.pi)
    c, s = torch.cos(a), torch.sin(a)
    x, y, z = xyz.unbind(dim=2)
    return torch.stack((x * c - y * s, x * s + y * c, z), dim=2)


def rotate_y(xyz, gen=None):
    """RotateY (:84-98, upright_axis=1 data sets): data @ [[c,0,s],[0,1,0],[-s,0,c]]  ->  x' = x c - z s, z' = x s + z c."""
    a = _rand(gen, (xyz.shape[0], 1), xyz.device) * (2 * math.pi)
    c, s = torch.cos(a), torch.sin(a)
    x, y, z = xyz.unbind(dim=2)
    return torch.stack((x * c - z * s, y, x * s + z * c), dim=2)


def jitter(xyz, sigma=0.01, clip=0.05, gen=None):
    """Jitter (:38-52): N(0, sigma) per coordinate, clipped at +-clip."""
    return xyz + torch.clamp(sigma * _randn(gen, xyz.shape, xyz.device), -clip, clip)


def scale(xyz, low=0.8, high=1.25, gen=None):
    """Scale (:127-137): one factor U(low, high) per cloud."""
    return xyz * (low + (high - low) * _rand(gen, (xyz.shape[0], 1, 1), xyz.device))


def rotate_small(xyz, angle_sigma=0.06, angle_clip=0.18, gen=None):
    """RotateSmall (:101-124): clipped N(0, sigma) angles about x, y, z; data @ (Rz Ry Rx)."""
    ang = torch.clamp(angle_sigma * _randn(gen, (xyz.shape[0], 3), xyz.device), -angle_clip, angle_clip)
    cx, sx, cy, sy, cz, sz = (f(ang[:, i]) for i in range(3) for f in (torch.cos, torch.sin))
    one, zero = torch.ones_like(cx), torch.zeros_like(cx)
    rx = torch.stack((one, zero, zero, zero, cx, -sx, zero, sx, cx), dim=1).view(-1, 3, 3)
    ry = torch.stack((cy, zero, sy, zero, one, zero, -sy, zero, cy), dim=1).view(-1, 3, 3)
    rz = torch.stack((cz, -sz, zero, sz, cz, zero, zero, zero, one), dim=1).view(-1, 3, 3)
    return torch.bmm(xyz, torch.bmm(rz, torch.bmm(ry, rx)))


def shift(xyz, shift_range=0.1, gen=None):
    """Shift (:55-64): one offset U(-range, range)^3 per cloud."""
    return xyz + (2 * _rand(gen, (xyz.shape[0], 1, 3), xyz.device) - 1) * shift_range


_ORDER = (("Rotate1D", rotate_z), ("Jitter", jitter), ("Scale", scale), ("RotateSmall", rotate_small), ("Shift", shift))


def apply_augmentations(xyz, names=("Jitter", "RotateSmall", "Shift", "Rotate1D"), gen=None, upright_axis=2):
    """get_augmentations_from_list + the per-cloud application of data/datagenerator.py, on a (B,N,3) device tensor.
    upright_axis=1 turns Rotate1D into the rotation about y (:17-18); any other value but 2 drops it, like the reference."""
    for name, fn in _ORDER:
        if names is not None and name in names:
            if name == "Rotate1D" and upright_axis != 2:
                if upright_axis != 1:
                    continue
                fn = rotate_y
            xyz = fn(xyz, gen=gen)
    return xyz

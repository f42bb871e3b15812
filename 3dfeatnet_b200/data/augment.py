"""The reference's augmentation objects (data/augment.py:4-137: get_augmentations_from_list, Jitter, Shift, RotateZ, RotateY,
RotateSmall, Scale, each with `.apply(data)`) as thin holders of their parameters around the batched functions of
3dfeatnet_b200/augment.py.  `data` may be what the reference passes -- one (N,3) NumPy cloud, as DataGenerator.next_triplet
does -- or a (N,3) / (B,N,3) torch tensor on any device; the result has the input's type, shape and dtype.  Each cloud of a
batch gets its own random draw.  `gen` (a torch.Generator on the data's device) makes the draws reproducible; without it the
global torch generator is used, as the reference uses the global NumPy one."""
import importlib

import numpy as np
import torch

_ROOT = __name__.split(".")[0]
_fn = importlib.import_module(("3dfeatnet_b200." if _ROOT == "3dfeatnet_b200" else "") + "augment")


def get_augmentations_from_list(str_list, upright_axis=2):
    """data/augment.py:4-29: the objects for the named augmentations, always in the order Rotate1D, Jitter, Scale,
    RotateSmall, Shift; Rotate1D turns about y when upright_axis == 1 and about z when it is 2."""
    if str_list is None:
        return []
    out = []
    if 'Rotate1D' in str_list and upright_axis in (1, 2):
        out.append(RotateY() if upright_axis == 1 else RotateZ())
    for name, cls in (('Jitter', Jitter), ('Scale', Scale), ('RotateSmall', RotateSmall), ('Shift', Shift)):
        if name in str_list:
            out.append(cls())
    return out


class Augmentation(object):
    """Base class (:32-35).  Subclasses set `_op` (a function of 3dfeatnet_b200/augment.py) and `_kwargs()`."""
    gen = None
    _op = None

    def _kwargs(self):
        return {}

    def apply(self, data):
        if self._op is None:
            raise NotImplementedError
        is_numpy = isinstance(data, np.ndarray)
        t = torch.as_tensor(data)
        if t.dim() not in (2, 3) or t.shape[-1] != 3:
            raise ValueError("augmentation expects (N,3) or (B,N,3) coordinates, got %s" % (tuple(t.shape),))
        x = t if t.is_floating_point() else t.float()
        y = type(self)._op(x if x.dim() == 3 else x.unsqueeze(0), gen=self.gen, **self._kwargs())
        y = (y if t.dim() == 3 else y.squeeze(0)).to(x.dtype)
        return y.numpy() if is_numpy else y


class Jitter(Augmentation):
    """N(0, sigma) noise per coordinate, clipped at +-clip (:38-52)."""
    _op = staticmethod(_fn.jitter)

    def __init__(self, sigma=0.01, clip=0.05, gen=None):
        assert clip > 0
        self.sigma, self.clip, self.gen = sigma, clip, gen

    def _kwargs(self):
        return dict(sigma=self.sigma, clip=self.clip)


class Shift(Augmentation):
    """One U(-shift_range, shift_range)^3 offset per cloud (:55-64)."""
    _op = staticmethod(_fn.shift)

    def __init__(self, shift_range=0.1, gen=None):
        self.shift_range, self.gen = shift_range, gen

    def _kwargs(self):
        return dict(shift_range=self.shift_range)


class RotateZ(Augmentation):
    """Uniform rotation about z (:67-81)."""
    _op = staticmethod(_fn.rotate_z)

    def __init__(self, gen=None):
        self.gen = gen


class RotateY(Augmentation):
    """Uniform rotation about y (:84-98)."""
    _op = staticmethod(_fn.rotate_y)

    def __init__(self, gen=None):
        self.gen = gen


class RotateSmall(Augmentation):
    """Small clipped-normal rotation about all three axes (:101-124)."""
    _op = staticmethod(_fn.rotate_small)

    def __init__(self, angle_sigma=0.06, angle_clip=0.18, gen=None):
        self.angle_sigma, self.angle_clip, self.gen = angle_sigma, angle_clip, gen

    def _kwargs(self):
        return dict(angle_sigma=self.angle_sigma, angle_clip=self.angle_clip)


class Scale(Augmentation):
    """One U(scale_low, scale_high) factor per cloud (:127-137)."""
    _op = staticmethod(_fn.scale)

    def __init__(self, scale_low=0.8, scale_high=1.25, gen=None):
        self.scale_low, self.scale_high, self.gen = scale_low, scale_high, gen

    def _kwargs(self):
        return dict(low=self.scale_low, high=self.scale_high)

    def apply(self, data, keypoints=None):
        return super().apply(data)

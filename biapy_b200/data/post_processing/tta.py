"""Orientation bookkeeping of the test-time augmentation: host-side mirror of ``biapy/data/post_processing/tta.py:58-260``.

Only the group elements are needed by the device path (``AxisTransform``, ``build_axis_transform_group``); the
representation-aware channel specs (``TTASpec``: flows, rays, affinities ...) belong to the instance-segmentation workflows
and are out of scope -- semantic segmentation and denoising run with ``tta_spec=None`` (every channel a scalar field).
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

#: valid values of ``TEST.AUGMENTATION_GROUP`` (reference tta.py:58)
TTA_GROUPS = ("auto", "full", "flips", "none")


@dataclass(frozen=True)
class AxisTransform:
    """Signed axis permutation over the spatial axes of a ``(spatial..., channels)`` array: output axis ``a`` is input axis
    ``perm[a]``, walked backwards when ``sign[a] == -1`` (reference ``tta.py:65-190``)."""

    perm: Tuple[int, ...]
    sign: Tuple[int, ...]

    def __post_init__(self):
        if sorted(self.perm) != list(range(len(self.perm))):
            raise ValueError("'perm' must be a permutation of range(ndim); got {}".format(self.perm))
        if len(self.sign) != len(self.perm):
            raise ValueError("'sign' and 'perm' must have the same length")
        if any(s not in (1, -1) for s in self.sign):
            raise ValueError("'sign' entries must be +1 or -1; got {}".format(self.sign))

    @property
    def ndim(self) -> int:
        return len(self.perm)

    @property
    def is_identity(self) -> bool:
        return self.perm == tuple(range(self.ndim)) and all(s == 1 for s in self.sign)

    @property
    def permutes_axes(self) -> bool:
        return self.perm != tuple(range(self.ndim))

    @classmethod
    def identity(cls, ndim: int) -> "AxisTransform":
        return cls(tuple(range(ndim)), (1,) * ndim)

    @property
    def inverse(self) -> "AxisTransform":
        inv = [0] * self.ndim
        for a, src in enumerate(self.perm):
            inv[src] = a
        return AxisTransform(tuple(inv), tuple(self.sign[inv[b]] for b in range(self.ndim)))

    def as_zyx(self) -> Tuple[Tuple[int, int, int], Tuple[int, int, int]]:
        """(perm, sign) over (z, y, x) for the C ABI: a 2D transform leaves a unit z axis in place."""
        if self.ndim == 3:
            return tuple(self.perm), tuple(self.sign)
        return (0,) + tuple(p + 1 for p in self.perm), (1,) + tuple(self.sign)

    def apply(self, arr):
        """Spatial transform of a CUDA tensor laid out ``(spatial..., channels)`` (reference ``tta.py:141-166``)."""
        from . import post_processing as pp
        return pp.orient_apply(arr[None], self, (0,) * self.ndim, "constant")[0]

    def describe(self) -> str:
        names = ("y", "x") if self.ndim == 2 else ("z", "y", "x")
        return ", ".join("{}<-{}{}".format(names[a], "+" if self.sign[a] > 0 else "-", names[self.perm[a]])
                         for a in range(self.ndim))


def build_axis_transform_group(ndim: int, level: str = "full",
                               interchangeable_axes: Optional[Sequence[int]] = None) -> List[AxisTransform]:
    """The orientations in the reference's deterministic order (``tta.py:197-260``): sign combinations outermost, axis
    permutations innermost, identity first.  8 (2D) / 16 (3D) for ``full``, 4 / 8 for ``flips``, 1 for ``none``; Z is never
    swapped with Y / X."""
    if ndim not in (2, 3):
        raise ValueError("ndim must be 2 or 3; got {}".format(ndim))
    if level not in ("full", "flips", "none"):
        raise ValueError("level must be one of 'full', 'flips', 'none'; got '{}'".format(level))
    if level == "none":
        return [AxisTransform.identity(ndim)]
    if interchangeable_axes is None:
        interchangeable_axes = (0, 1) if ndim == 2 else (1, 2)
    slots = tuple(sorted(interchangeable_axes))
    perms = []
    if level == "flips":
        perms.append(tuple(range(ndim)))
    else:
        for order in itertools.permutations(slots):
            p = list(range(ndim))
            for slot, src in zip(slots, order):
                p[slot] = src
            perms.append(tuple(p))
    group = [AxisTransform(p, signs) for signs in itertools.product((1, -1), repeat=ndim) for p in perms]
    group.sort(key=lambda t: not t.is_identity)            # stable: identity first, the rest keep their order
    return group

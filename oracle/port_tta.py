"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's scalar test-time augmentation.

Follows ``biapy/data/post_processing/tta.py`` (AxisTransform :65-190, build_axis_transform_group :197-260) and
``biapy/data/post_processing/post_processing.py`` (_pad_for_orientations :1285-1339, _crop_padding :1342-1346,
_reduce_orientations :1349-1383, ensemble_predictions :1386-1555) for ``tta_spec=None``.  Pinned against the reference's own
functions by ``tests/test_oracle_golden.py`` (container only) and the committed ``tests/golden/tta_*.npz`` fixtures.
Only tests, ``__graft_entry__.smoke()`` and bench baselines may import this module.
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np


def orientations(ndim: int, level: str = "full") -> List[Tuple[Tuple[int, ...], Tuple[int, ...]]]:
    """(perm, sign) pairs in the reference's order (tta.py:236-260)."""
    if level == "none":
        return [(tuple(range(ndim)), (1,) * ndim)]
    inter = (0, 1) if ndim == 2 else (1, 2)
    if level == "flips":
        perms = [tuple(range(ndim))]
    else:
        perms = []
        for sub in itertools.permutations(inter):
            p = list(range(ndim))
            for slot, src in zip(inter, sub):
                p[slot] = src
            perms.append(tuple(p))
    out = [(p, s) for s in itertools.product((1, -1), repeat=ndim) for p in perms]
    ident = (tuple(range(ndim)), (1,) * ndim)
    out.sort(key=lambda t: t != ident)
    return out


def inverse(perm, sign):
    """tta.py:123-139."""
    n = len(perm)
    inv = [0] * n
    for a, p in enumerate(perm):
        inv[p] = a
    return tuple(inv), tuple(sign[inv[b]] for b in range(n))


def apply(arr: np.ndarray, perm, sign) -> np.ndarray:
    """tta.py:141-166: transpose the spatial axes, then flip the reversed ones."""
    n = len(perm)
    out = np.transpose(arr, tuple(perm) + (n,))
    flips = tuple(a for a in range(n) if sign[a] < 0)
    if flips:
        out = np.flip(out, axis=flips)
    return np.ascontiguousarray(out)


def pad_for_orientations(img: np.ndarray, orients, pad_mode: str):
    """post_processing.py:1285-1339."""
    n = img.ndim - 1
    moved = set()
    for perm, _ in orients:
        for a in range(n):
            if perm[a] != a:
                moved.update((a, perm[a]))
    if not moved:
        return img, None
    target = max(img.shape[a] for a in moved)
    if all(img.shape[a] == target for a in moved):
        return img, None
    pad_before = [0] * n
    for a in moved:
        pad_before[a] = target - img.shape[a]
    if pad_mode == "reflect" and any(pad_before[a] >= img.shape[a] for a in moved):
        pad_mode = "edge"
    pad_w = [(pad_before[a], 0) for a in range(n)] + [(0, 0)]
    return np.pad(img, pad_w, mode=pad_mode), tuple(pad_before)


def reduce_orientations(stack: np.ndarray, mode: str) -> np.ndarray:
    """post_processing.py:1349-1383 with mode_channels=None; the mean is written out the way numpy evaluates it for a
    float32 stack (orientation after orientation, one division)."""
    if mode == "mean":
        acc = stack[0].astype(np.float32, copy=True)
        for k in range(1, stack.shape[0]):
            acc = acc + stack[k]
        return acc / np.float32(stack.shape[0])
    return (np.min if mode == "min" else np.max)(stack, axis=0)


def ensemble_predictions(o_img: np.ndarray, pred_func: Callable[[np.ndarray], np.ndarray], ndim: int, batch_size_value: int = 1,
                         mode: str = "mean", group: str = "auto") -> np.ndarray:
    """post_processing.py:1466-1531 for a pred_func that maps ``(batch, spatial..., C)`` numpy -> ``(batch, spatial..., C_out)``
    numpy.  Returns ``(spatial..., C_out)`` float32."""
    orients = orientations(ndim, "full" if group == "auto" else group)
    img, pad_before = pad_for_orientations(o_img, orients, "reflect")
    aug = np.stack([apply(img, p, s) for p, s in orients], axis=0)
    preds = []
    total = aug.shape[0]
    for i in range(int(math.ceil(total / batch_size_value))):
        low, top = i * batch_size_value, min((i + 1) * batch_size_value, total)
        preds.append(pred_func(aug[low:top]))
    pred = np.concatenate(preds, axis=0).astype(np.float32)
    for k, (p, s) in enumerate(orients):
        pred[k] = apply(pred[k], *inverse(p, s))
    out = reduce_orientations(pred, mode)
    if pad_before is not None:
        out = out[tuple(slice(q, None) for q in pad_before) + (slice(None),)]
    return out


def toy_pred_func(batch: np.ndarray) -> np.ndarray:
    """A deliberately non-equivariant 'network' for fixtures: position ramps, a shifted copy and a channel mix, so that every
    orientation contributes a different value at every voxel.  ``(b, spatial..., C)`` -> ``(b, spatial..., 2)`` float32."""
    x = batch.astype(np.float32)
    nd = x.ndim - 2
    out = np.zeros(x.shape[:-1] + (2,), np.float32)
    ramp = np.float32(0.0)
    for a in range(nd):
        shp = [1] * x.ndim
        shp[1 + a] = x.shape[1 + a]
        ramp = ramp + (np.arange(x.shape[1 + a], dtype=np.float32).reshape(shp) * np.float32(0.37 * (a + 1)))
    out[..., 0] = x[..., 0] * np.float32(1.5) + np.roll(x[..., -1], 1, axis=nd) * np.float32(0.25) + ramp[..., 0] * np.float32(0.01)
    out[..., 1] = np.tanh(x.sum(-1)) - np.roll(x[..., 0], 2, axis=1) * np.float32(0.5)
    return out

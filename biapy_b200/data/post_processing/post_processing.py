"""Test-time augmentation on the device: mirror of ``ensemble_predictions``
(``biapy/data/post_processing/post_processing.py:1386-1555``) for scalar predictions (``tta_spec=None``: semantic
segmentation, denoising, the workflows of the hot path).

The reference stacks the orientations with numpy, sends every batch host -> device -> host and un-transforms / reduces in
numpy.  Here the image stays in HBM: ``b200_orient_apply`` writes each padded orientation, the predictions stay on the device
and ``b200_orient_reduce`` reads all of them through their inverse transforms, reduces (mean / min / max, numpy's float32
arithmetic) and crops the padding in one pass.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import _lib, ops
from .. import _stitch
from .tta import TTA_GROUPS, AxisTransform, build_axis_transform_group

_MODES = {"mean": 0, "min": 1, "max": 2}


def _i3(v) -> "C.Array":
    return (C.c_int32 * 3)(*[int(x) for x in v])


def _pad_plan(shape: Sequence[int], orientations: List[AxisTransform], pad_mode: str) -> Tuple[Optional[Tuple[int, ...]], str]:
    """``_pad_for_orientations`` (reference ``:1285-1339``) without touching data: front padding per spatial axis (or None)
    and the pad mode actually used (``reflect`` degrades to ``edge`` when a pad reaches the axis length)."""
    n = len(shape)
    moved = set()
    for t in orientations:
        for a in range(n):
            if t.perm[a] != a:
                moved.update((a, t.perm[a]))
    if not moved:
        return None, pad_mode
    target = max(shape[a] for a in moved)
    if all(shape[a] == target for a in moved):
        return None, pad_mode
    pad = [0] * n
    for a in moved:
        pad[a] = target - shape[a]
    if pad_mode == "reflect" and any(pad[a] >= shape[a] for a in moved):
        pad_mode = "edge"
    return tuple(pad), pad_mode


def orient_apply(batch: torch.Tensor, t: AxisTransform, pad_before: Sequence[int], pad_mode: str) -> torch.Tensor:
    """batch: CUDA ``(N, spatial..., C)``.  Returns ``T(pad_front(batch))`` per element, contiguous."""
    _lib.require_cuda(batch, "TTA input")
    nd = t.ndim
    x5 = batch if nd == 3 else batch[:, None]
    x5 = x5.contiguous()
    perm, sign = t.as_zyx()
    pad3 = tuple(pad_before) if nd == 3 else (0,) + tuple(pad_before)
    sp = [x5.shape[1 + a] + pad3[a] for a in range(3)]
    out = torch.empty((x5.shape[0],) + tuple(sp[perm[a]] for a in range(3)) + (x5.shape[4],), dtype=x5.dtype, device=x5.device)
    ops._launch("b200_orient_apply", ops._ref(x5), ops._ref(out), _i3(perm), _i3(sign), _i3(pad3), _lib.PAD_MODE[pad_mode],
                _lib.stream_ptr())
    return out if nd == 3 else out[:, 0]


def orient_reduce(pred: torch.Tensor, orientations: List[AxisTransform], mode: str, pad_before: Optional[Sequence[int]]) -> torch.Tensor:
    """pred: CUDA ``(n_orientations, spatial..., C)``, prediction k made on orientation k.  Returns the float32
    ``(spatial..., C)`` ensemble with the front padding cropped."""
    nd = orientations[0].ndim
    p5 = (pred if nd == 3 else pred[:, None]).contiguous()
    k = len(orientations)
    assert p5.shape[0] == k, (p5.shape, k)
    perms = (C.c_int32 * (3 * k))(*[v for t in orientations for v in t.as_zyx()[0]])
    signs = (C.c_int32 * (3 * k))(*[v for t in orientations for v in t.as_zyx()[1]])
    pad = tuple(pad_before) if pad_before is not None else (0,) * nd
    pad3 = pad if nd == 3 else (0,) + pad
    out = torch.empty((1,) + tuple(p5.shape[1 + a] - pad3[a] for a in range(3)) + (p5.shape[4],), dtype=torch.float32, device=p5.device)
    ops._launch("b200_orient_reduce", ops._ref(p5), perms, signs, _MODES[mode], _i3(pad3), ops._ref(out), _lib.stream_ptr())
    return out[0] if nd == 3 else out[0, 0]


def ensemble_predictions(o_img, pred_func: Callable, axes_order_back: Tuple[int, ...], axes_order: Tuple[int, ...], device,
                         ndim: int, batch_size_value: int = 1, mode: str = "mean", tta_spec=None, group: str = "auto",
                         verbose: bool = False):
    """Same signature and result as the reference function.  ``o_img``: numpy array or CUDA tensor ``(spatial..., C)``;
    ``pred_func`` receives a CUDA tensor ``(batch, spatial..., C)`` and returns ``(batch, C_out, spatial...)`` (or a dict with
    ``"pred"``) like ``model_call_func``.  Returns a float32 CUDA tensor ``(1, C_out, spatial...)`` (``axes_order`` layout),
    or a dict when the model returns extra outputs."""
    assert mode in ["mean", "min", "max"], "Get unknown ensemble mode {}".format(mode)
    assert ndim in (2, 3), "ndim must be 2 or 3, got {}".format(ndim)
    assert group in TTA_GROUPS, "group must be one of {}, got '{}'".format(TTA_GROUPS, group)
    if o_img.ndim != ndim + 1:
        raise ValueError("Expected a {}D input (spatial..., channels); got shape {}".format(ndim, tuple(o_img.shape)))
    if tta_spec is not None:
        raise NotImplementedError("representation-aware TTA specs (flows, rays, affinities ...) belong to the instance-segmentation "
                                  "workflows; the B200 hot path ensembles scalar predictions (tta_spec=None)")
    orientations = build_axis_transform_group(ndim, level=("full" if group == "auto" else group))
    if verbose:
        print("TTA: {} orientation(s) over scalar channels".format(len(orientations)))
    img = _stitch.to_device(o_img, device)
    pad_before, pad_mode = _pad_plan(tuple(img.shape[:ndim]), orientations, "reflect")
    pads = pad_before if pad_before is not None else (0,) * ndim

    # every orientation of the (front-padded) image, stacked on the batch axis: (n_orient, spatial', C)
    aug = torch.cat([orient_apply(img[None], t, pads, pad_mode) for t in orientations], 0)
    total = aug.shape[0]
    preds: List[torch.Tensor] = []
    extra: Dict[str, List] = {}
    for i in range(int(math.ceil(total / batch_size_value))):
        low, top = i * batch_size_value, min((i + 1) * batch_size_value, total)
        r = pred_func(aug[low:top])
        if isinstance(r, dict):
            for key, val in r.items():
                if key != "pred":
                    extra.setdefault(key, []).append(val)
            r = r["pred"]
        r = r.permute(axes_order_back)                      # channels-last view, still on the device
        if r.dim() == ndim + 1:
            r = r[None]
        preds.append(r)
    pred = torch.cat([p.float() for p in preds], 0).contiguous()
    if tuple(pred.shape[1:1 + ndim]) != tuple(aug.shape[1:1 + ndim]):
        raise ValueError("TTA needs the prediction to keep the input's spatial shape to undo the augmentation; "
                         "got {} for an input of {}".format(tuple(pred.shape[1:1 + ndim]), tuple(aug.shape[1:1 + ndim])))
    out = orient_reduce(pred, orientations, mode, pad_before)[None].permute(axes_order)
    if not extra:
        return out
    rest: Dict = {}
    for key, chunks in extra.items():
        arrs = []
        for c in chunks:
            a = c.permute(axes_order_back) if (isinstance(c, torch.Tensor) and c.dim() == ndim + 2) else c
            arrs.append(a)
        ok = all(isinstance(a, torch.Tensor) and a.dim() == ndim + 2 for a in arrs)
        stacked = torch.cat([a.float() for a in arrs], 0).contiguous() if ok else None
        if stacked is None or stacked.shape[0] != len(orientations) or tuple(stacked.shape[1:1 + ndim]) != tuple(aug.shape[1:1 + ndim]):
            rest[key] = chunks[0]                           # not a matching spatial map: keep it as the model gave it
            continue
        rest[key] = orient_reduce(stacked, orientations, mode, pad_before)[None].permute(axes_order)
    rest["pred"] = out
    return rest

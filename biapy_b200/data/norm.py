"""Image normalisation at the two ends of the inference path, on the device: mirror of ``normalize_image`` /
``undo_image_norm`` (``biapy/data/norm.py:44-220, 408-780``) with the reference's ``norm_module`` / ``norm_info`` dictionaries.

The statistics (min / max / mean / std / "is this channel binary") come from one pass of ``b200_image_stats``; the
normalisation itself is one float32 stream (``b200_image_norm_apply``: clip -> subtract -> divide in the reference's operation
order, so given the same statistics the result is bit-identical to numpy's).  Differences kept deliberately small and loud:

* percentile bounds must be given as values (``lower_bound_val`` / ``upper_bound_val`` or a ``per_channel_info`` from a
  previous call); computing percentiles from the data (a global selection) raises ``NotImplementedError``;
* ``out_dtype`` must be ``float32`` (what the engine consumes).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import _lib, ops
from . import _stitch

_TORCH_DTYPE = {"uint8": torch.uint8, "uint16": torch.uint16, "float32": torch.float32}
_EPS = 1e-6


def image_stats(img: torch.Tensor, clip: Optional[np.ndarray] = None) -> np.ndarray:
    """Per-channel (min, max, sum, sum of squares, is_binary) of a CUDA image ``(..., C)`` as a float64 host array;
    `clip` (C, 3) float32 = (on, lo, hi) per channel: statistics of the clipped values."""
    _lib.require_cuda(img, "image")
    img = img.contiguous()
    c = img.shape[-1]
    out = torch.empty((c, 5), dtype=torch.float64, device=img.device)
    cp = np.ascontiguousarray(clip, np.float32).ctypes.data_as(C.POINTER(C.c_float)) if clip is not None else None
    ops._launch("b200_image_stats", ops._ptr(img), _lib.torch_dtype_code(img.dtype), img.numel() // c, c, cp, ops._ptr(out),
                _lib.stream_ptr())
    return out.cpu().numpy()


def _per_channel(norm_module: Dict, key: str, c: int) -> Optional[List[float]]:
    if key not in norm_module:
        return None
    v = norm_module[key]
    assert isinstance(v, list), f"'{key}' should be a list of float/integer values"
    if v[0] == -1:
        return None
    if len(v) == 1:
        return [float(v[0])] * c
    assert len(v) == c, f"If more that one {key} value is provided, the number of values should be the same as the number of channels"
    return list(v)


def normalize_image(img, norm_module: Dict, apply_norm: bool = True) -> Tuple[object, Dict]:
    """Same contract as the reference: returns ``(normalised image, norm_info)``.  ``img``: numpy array or CUDA tensor
    ``([z,] y, x, C)`` of dtype uint8 / uint16 / float32; the result has the container type of the input."""
    assert img.ndim >= 3, "Data should be at least 3D. E.g. (y, x, channels) in 2D and (z, y, x, channels) in 3D"
    assert "type" in norm_module, "'type' key should be in 'norm_module' dict"
    assert norm_module["type"] in ["div", "scale_range", "zero_mean_unit_variance"], (
        "Invalid normalization type. Expected values are: 'div', 'scale_range' and 'zero_mean_unit_variance'")
    assert "percentile_clip" in norm_module, "'percentile_clip' key should be in 'norm_module' dict"
    assert isinstance(norm_module["percentile_clip"], bool), "'percentile_clip' should be a boolean value"
    assert "out_dtype" in norm_module, "'out_dtype' key should be in 'norm_module' dict"
    if norm_module["out_dtype"] != "float32":
        raise NotImplementedError("the B200 engine consumes float32 images: norm_module['out_dtype'] must be 'float32'")
    dev = _stitch.to_device(img)
    if str(dev.dtype).replace("torch.", "") not in _TORCH_DTYPE:
        raise NotImplementedError(f"image dtype {dev.dtype} is not supported (uint8, uint16, float32)")
    c = dev.shape[-1]
    kind = norm_module["type"]
    info: Dict = {"type": kind, "percentile_clip": norm_module["percentile_clip"], "orig_dtype": str(img.dtype).replace("torch.", ""),
                  "out_dtype": norm_module["out_dtype"], "per_channel_info": {}}
    pci = norm_module.get("per_channel_info")
    if pci is not None:
        assert isinstance(pci, dict) and len(pci) == c, "The number of channels in 'per_channel_info' should be the same as the number of channels in the input data"
    lo = hi = None
    if norm_module["percentile_clip"]:
        if pci is not None:
            lo = [pci[str(k)].get("lower_bound_val") for k in range(c)]
            hi = [pci[str(k)].get("upper_bound_val") for k in range(c)]
        else:
            lo, hi = _per_channel(norm_module, "lower_bound_val", c), _per_channel(norm_module, "upper_bound_val", c)
        if lo is None or hi is None or any(v is None for v in lo + hi):
            raise NotImplementedError("percentile bounds computed from the data (per_lower_bound / per_upper_bound) are not implemented "
                                      "on the device; pass lower_bound_val / upper_bound_val or a per_channel_info")
    raw = image_stats(dev)
    nvox = dev.numel() // c
    params = np.zeros((c, 6), np.float32)
    chans: List[Dict] = [dict() for _ in range(c)]
    if norm_module["percentile_clip"]:
        for k in range(c):
            if raw[k, 4]:                          # binary channels are never clipped (norm.py:441-443)
                chans[k]["lower_bound_val"], chans[k]["upper_bound_val"] = 0.0, 1.0
            else:
                chans[k]["lower_bound_val"], chans[k]["upper_bound_val"] = lo[k], hi[k]
                if apply_norm:
                    params[k, 0:3] = (1.0, lo[k], hi[k])
    # the reference clips each channel in place first: min / max / mean / std / is_binary below are those of the clipped channel
    stats = image_stats(dev, params[:, 0:3]) if params[:, 0].any() else raw
    for k in range(c):
        ch = chans[k]
        mn, mx, s1, s2, is_bin = stats[k]
        if kind in ("div", "scale_range"):
            a = pci[str(k)].get("max_val_to_div") if pci is not None else None
            b = pci[str(k)].get("min_val_to_div") if pci is not None else None
            if (a is None) != (b is None):
                raise ValueError("If 'max_val_to_div' is provided, 'min_val_to_div' should also be provided")
            if is_bin:
                mx_v, mn_v = 1.0, 0.0
            elif a is not None:
                mx_v, mn_v = float(a), float(b)
                params[k, 3:6] = (1.0, mn_v, max(mx_v - mn_v, _EPS))
            else:
                dmax, dmin = mx, mn
                if kind == "scale_range":
                    mx_v, mn_v = float(dmax), float(dmin)
                else:
                    mx_v, mn_v = (65535 if dmax > 255 else 255), 0
                params[k, 3:6] = (1.0, mn_v, max(mx_v - mn_v, _EPS))
            if not apply_norm:
                params[k, 3] = 0.0
            ch["min_val_to_div"], ch["max_val_to_div"] = mn_v, mx_v
        else:
            mean = pci[str(k)].get("mean") if pci is not None else None
            std = pci[str(k)].get("std") if pci is not None else None
            if pci is None:
                m_l, s_l = _per_channel(norm_module, "mean", c), _per_channel(norm_module, "std", c)
                mean = m_l[k] if m_l is not None else None
                std = s_l[k] if s_l is not None else None
            if is_bin:
                mean_v, std_v = 0.0, 1.0
            else:
                mean_v = float(np.float32(s1 / nvox)) if mean is None else mean
                std_v = float(np.float32(np.sqrt(max(s2 / nvox - (s1 / nvox) ** 2, 0.0)))) if std is None else std
                if apply_norm:
                    params[k, 3:6] = (2.0, mean_v, max(std_v, _EPS))
            ch["mean"], ch["std"] = float(mean_v), float(std_v)
        info["per_channel_info"][str(k)] = ch
    if not apply_norm and dev.dtype == torch.float32:
        out = dev
    else:
        out = torch.empty(dev.shape, dtype=torch.float32, device=dev.device)
        src = dev.contiguous()
        ops._launch("b200_image_norm_apply", ops._ptr(src), _lib.torch_dtype_code(src.dtype), nvox, c,
                    params.ctypes.data_as(C.POINTER(C.c_float)), ops._ptr(out), _lib.stream_ptr())
    return _stitch.like_input(out, img), info


def undo_image_norm(data, norm_info: Dict):
    """Inverse of :func:`normalize_image` with the reference's arithmetic (float64 products, truncating integer casts)."""
    assert "type" in norm_info, "'type' key should be in 'norm_info' dict. Ensure you input the same normalization dict used to normalize the data previously"
    assert "per_channel_info" in norm_info, "'per_channel_info' key should be in 'norm_info' dict. Ensure you input the same normalization dict used to normalize the data previously"
    dev = _stitch.to_device(data)
    if dev.dtype != torch.float32:
        raise NotImplementedError("undo_image_norm takes the float32 prediction of the engine")
    c = dev.shape[-1]
    pci = norm_info["per_channel_info"]
    assert len(pci) == c, "The number of channels in the input data should be the same as the number of channels in 'per_channel_info' in 'norm_info'"
    params = np.zeros((c, 3), np.float64)
    for k in range(c):
        if norm_info["type"] in ("div", "scale_range"):
            assert "max_val_to_div" in pci[str(k)] and "min_val_to_div" in pci[str(k)], f"'max_val_to_div' / 'min_val_to_div' missing for channel {k}"
            params[k] = (1.0, pci[str(k)]["max_val_to_div"], pci[str(k)]["min_val_to_div"])
        else:
            assert "mean" in pci[str(k)] and "std" in pci[str(k)], f"'mean' / 'std' missing for channel {k}"
            params[k] = (2.0, pci[str(k)]["std"], pci[str(k)]["mean"])
    orig = norm_info["orig_dtype"]
    if orig not in _TORCH_DTYPE:
        raise NotImplementedError(f"orig_dtype {orig!r} is not supported (uint8, uint16, float32)")
    out = torch.empty(dev.shape, dtype=_TORCH_DTYPE[orig], device=dev.device)
    src = dev.contiguous()
    ops._launch("b200_image_denorm_apply", ops._ptr(src), dev.numel() // c, c, params.ctypes.data_as(C.POINTER(C.c_double)), ops._ptr(out),
                _lib.torch_dtype_code(out.dtype), _lib.stream_ptr())
    return _stitch.like_input(out, data)


def binarize_prediction(pred, n_classes: int, threshold: float = 0.5):
    """Binarisation behind the merge (``semantic_seg.py:418-425, 524-531``): ``pred > threshold`` as uint8 for binary problems
    (the by-chunks path's fixed 0.5; pass the Otsu threshold of the whole-image path explicitly), else the channel arg-max as
    uint8 (uint16 from 255 classes) with a trailing unit channel."""
    dev = _stitch.to_device(pred)
    if dev.dtype != torch.float32:
        raise NotImplementedError("binarize_prediction takes the float32 prediction of the engine")
    dev = dev.contiguous()
    if n_classes <= 2:
        out = torch.empty(dev.shape, dtype=torch.uint8, device=dev.device)
        ops._launch("b200_binarize", ops._ptr(dev), dev.numel(), float(threshold), ops._ptr(out), _lib.stream_ptr())
    else:
        dt = torch.uint8 if n_classes < 255 else torch.uint16
        out = torch.empty(tuple(dev.shape[:-1]) + (1,), dtype=dt, device=dev.device)
        ops._launch("b200_argmax_channels", ops._ptr(dev), dev.numel() // dev.shape[-1], dev.shape[-1], ops._ptr(out),
                    _lib.torch_dtype_code(dt), _lib.stream_ptr())
    return _stitch.like_input(out, pred)

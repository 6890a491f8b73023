"""3D sliding-window crop / merge with BiaPy's signatures, executed on the GPU.

Drop-in for the hot functions of ``biapy/data/data_3D_manipulation.py``:

* :func:`crop_3D_data_with_overlap`   (reference ``:353-636``)
* :func:`merge_3D_data_with_overlap`  (reference ``:690-859``)

Argument names, defaults, return conventions, patch order, coordinates and error types follow the reference.
Inputs may be numpy arrays (copied to the GPU and back, like a user of the reference would call them) or
CUDA tensors (stay on the device).  There is no CPU implementation.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import _stitch
from .dataset import PatchCoords

__all__ = ["crop_3D_data_with_overlap", "merge_3D_data_with_overlap"]


def _check_overlap(overlap):
    if (overlap[0] >= 1 or overlap[0] < 0) or (overlap[1] >= 1 or overlap[1] < 0) or (overlap[2] >= 1 or overlap[2] < 0):
        raise ValueError("'overlap' values must be floats between range [0, 1)")


def crop_3D_data_with_overlap(data, vol_shape: Tuple[int, ...], data_mask=None,
                              overlap: Tuple[float, ...] = (0, 0, 0), padding: Tuple[int, ...] = (0, 0, 0),
                              verbose: bool = True, median_padding: bool = False, load_data: bool = True,
                              pad_type: str = "reflect"):
    """Crop a ``(z, y, x, c)`` volume into overlapping ``vol_shape`` patches (reference ``:353-363``).

    Returns ``(patches[, mask_patches], List[PatchCoords])``, or only the coordinates when ``load_data`` is
    False (that mode needs no GPU: it only runs the host planner).
    """
    if verbose:
        print("### 3D-OV-CROP ###")
        print("Cropping {} images into {} with overlapping . . .".format(tuple(data.shape), vol_shape))
        print("Minimum overlap selected: {}".format(overlap))
        print("Padding: {}".format(padding))
    if data.ndim != 4:
        raise ValueError("data expected to be 4 dimensional, given {}".format(tuple(data.shape)))
    if data_mask is not None:
        if data_mask.ndim != 4:
            raise ValueError("data_mask expected to be 4 dimensional, given {}".format(tuple(data_mask.shape)))
        if tuple(data.shape[:-1]) != tuple(data_mask.shape[:-1]):
            raise ValueError("data and data_mask shapes mismatch: {} vs {}".format(tuple(data.shape[:-1]), tuple(data_mask.shape[:-1])))
    if len(vol_shape) != 4:
        raise ValueError("vol_shape expected to be of length 4, given {}".format(vol_shape))
    for i, p in enumerate(padding):
        if p >= vol_shape[i] // 2:
            raise ValueError(
                "'Padding' can not be greater than half of 'vol_shape'. Max value for the given input shape {} is {}".format(
                    vol_shape, ((vol_shape[0] // 2) - 1, (vol_shape[1] // 2) - 1, (vol_shape[2] // 2) - 1)))
    for i in range(3):
        if vol_shape[i] > data.shape[i]:
            raise ValueError(
                "'vol_shape[{}]' {} greater than {} (you can reduce 'DATA.PATCH_SIZE' or use 'DATA.REFLECT_TO_COMPLETE_SHAPE')".format(
                    i, vol_shape[i], data.shape[i]))
    _check_overlap(overlap)
    if median_padding:
        # reference :524-535 has an indexing typo on the y axis (SURVEY 8a addendum); not on any BASELINE path
        raise NotImplementedError("median_padding=True is not supported by biapy_b200")

    axes = [_stitch.Axis(data.shape[i], vol_shape[i], padding[i], overlap[i]) for i in range(3)]
    starts = [a.starts(0) for a in axes]
    if verbose:
        real = [(a.core - a.ov_px - a.step + 0) for a in axes]  # per-block overlap is folded into `step`
        print("{} patches per (z,y,x) axis".format(tuple(a.n for a in axes)))

    crop_coords: List[PatchCoords] = []
    for z in starts[0]:
        for y in starts[1]:
            for x in starts[2]:
                crop_coords.append(PatchCoords(z_start=int(z), z_end=int(z) + vol_shape[0], y_start=int(y),
                                               y_end=int(y) + vol_shape[1], x_start=int(x), x_end=int(x) + vol_shape[2]))
    if not load_data:
        if verbose:
            print("### END 3D-OV-CROP ###")
        return crop_coords

    dev = _stitch.to_device(data)
    cropped = _stitch.like_input(_stitch.crop_device(dev, vol_shape[:3], starts, padding, pad_type), data)
    if data_mask is not None:
        mdev = _stitch.to_device(data_mask)
        cropped_mask = _stitch.like_input(_stitch.crop_device(mdev, vol_shape[:3], starts, padding, pad_type), data_mask)
    if verbose:
        print("**** New data shape is: {}".format(tuple(cropped.shape)))
        print("### END 3D-OV-CROP ###")
    if data_mask is not None:
        return cropped, cropped_mask, crop_coords
    return cropped, crop_coords


def merge_3D_data_with_overlap(data, orig_vol_shape: Tuple, data_mask=None, overlap: Tuple[float, ...] = (0, 0, 0),
                               padding: Tuple[int, ...] = (0, 0, 0), verbose: bool = True):
    """Merge ``(n, z, y, x, c)`` patches into a ``orig_vol_shape`` volume with the spline-weighted
    overlap-add of the reference (``:690-697``, ``:822-849``).  Returns an array of ``data``'s dtype, or a
    ``(data, mask)`` pair."""
    assert data.ndim == 5, f"data expected to be 5 dimensional, given {tuple(data.shape)}"
    assert len(orig_vol_shape) == 4, f"orig_vol_shape expected to be 4 dimensional, given {orig_vol_shape}"
    if data_mask is not None:
        if tuple(data.shape[:-1]) != tuple(data_mask.shape[:-1]):
            raise ValueError("data and data_mask shapes mismatch: {} vs {}".format(tuple(data.shape[:-1]), tuple(data_mask.shape[:-1])))
    _check_overlap(overlap)
    if verbose:
        print("### MERGE-3D-OV-CROP ###")
        print("Merging {} images into {} with smooth blending . . .".format(tuple(data.shape), orig_vol_shape))
        print("Minimum overlap selected: {}".format(overlap))
        print("Padding: {}".format(padding))

    axes = [_stitch.Axis(orig_vol_shape[i], data.shape[1 + i], padding[i], overlap[i]) for i in range(3)]
    starts = [a.starts(1) for a in axes]
    wins = [a.window() for a in axes]

    def run(arr):
        dt = str(arr.dtype).replace("torch.", "")
        if dt not in ("float32", "float16", "bfloat16"):
            raise TypeError(f"merge_3D_data_with_overlap: unsupported dtype {arr.dtype} (float32/float16/bfloat16)")
        dev = _stitch.to_device(arr)
        out = _stitch.merge_device(dev, orig_vol_shape[:3], starts, wins, padding)
        return _stitch.like_input(out, arr)

    merged = run(data)
    if verbose:
        print("**** New data shape is: {}".format(tuple(merged.shape)))
        print("### END MERGE-3D-OV-CROP ###")
    if data_mask is not None:
        return merged, run(data_mask)
    return merged

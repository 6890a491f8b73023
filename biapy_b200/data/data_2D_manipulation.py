"""2D sliding-window crop / merge with BiaPy's signatures, executed on the GPU.

Drop-in for ``crop_data_with_overlap`` (``biapy/data/data_2D_manipulation.py:54-316``) and
``merge_data_with_overlap`` (``:366-533``).  A stack ``(n_img, y, x, c)`` is treated as a volume whose z axis
is the image index with a z-patch of one slice, so the 3D kernels serve both.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import _stitch
from .dataset import PatchCoords

__all__ = ["crop_data_with_overlap", "merge_data_with_overlap"]


def _check_overlap(overlap):
    if (overlap[0] >= 1 or overlap[0] < 0) or (overlap[1] >= 1 or overlap[1] < 0):
        raise ValueError("'overlap' values must be floats between range [0, 1)")


def crop_data_with_overlap(data, crop_shape: Tuple[int, ...], data_mask=None, overlap: Tuple[float, ...] = (0, 0),
                           padding: Tuple[int, ...] = (0, 0), verbose: bool = True, load_data: bool = True,
                           pad_type: str = "reflect"):
    """Crop ``(n_img, y, x, c)`` images into ``crop_shape = (y, x, c)`` patches (reference ``:54-63``)."""
    if data.ndim != 4:
        raise ValueError("data expected to be 4 dimensional, given {}".format(tuple(data.shape)))
    if data_mask is not None:
        if data_mask.ndim != 4:
            raise ValueError("data mask expected to be 4 dimensional, given {}".format(tuple(data_mask.shape)))
        if tuple(data.shape[:-1]) != tuple(data_mask.shape[:-1]):
            raise ValueError("data and data_mask shapes mismatch: {} vs {}".format(tuple(data.shape[:-1]), tuple(data_mask.shape[:-1])))
    for i, p in enumerate(padding):
        if p >= crop_shape[i] // 2:
            raise ValueError("'Padding' can not be greater than the half of 'crop_shape'. Max value for this {} input shape is {}".format(
                tuple(data.shape), [(crop_shape[0] // 2) - 1, (crop_shape[1] // 2) - 1]))
    if len(crop_shape) != 3:
        raise ValueError("crop_shape expected to be of length 3, given {}".format(crop_shape))
    for i in range(2):
        if crop_shape[i] > data.shape[1 + i]:
            raise ValueError(
                "'crop_shape[{}]' {} greater than {} (you can reduce 'DATA.PATCH_SIZE' or use 'DATA.REFLECT_TO_COMPLETE_SHAPE')".format(
                    i, crop_shape[i], data.shape[1 + i]))
    _check_overlap(overlap)
    if verbose:
        print("### OV-CROP ###")
        print("Cropping {} images into {} with overlapping. . .".format(tuple(data.shape), crop_shape))
        print("Minimum overlap selected: {}".format(overlap))
        print("Padding: {}".format(padding))

    axes = [_stitch.Axis(data.shape[1 + i], crop_shape[i], padding[i], overlap[i]) for i in range(2)]
    sy, sx = axes[0].starts(0), axes[1].starts(0)
    if verbose:
        print("{} patches per (y,x) axis".format((axes[0].n, axes[1].n)))
    crop_coords: List[PatchCoords] = []
    for _ in range(data.shape[0]):
        for y in sy:
            for x in sx:
                crop_coords.append(PatchCoords(y_start=int(y), y_end=int(y) + crop_shape[0], x_start=int(x),
                                               x_end=int(x) + crop_shape[1]))
    if not load_data:
        if verbose:
            print("### END OV-CROP ###")
        return crop_coords

    sz, _ = _stitch.identity_axis(data.shape[0])

    def run(arr):
        dev = _stitch.to_device(arr)
        out = _stitch.crop_device(dev, (1, crop_shape[0], crop_shape[1]), [sz, sy, sx], (0, padding[0], padding[1]), pad_type)
        return _stitch.like_input(out[:, 0], arr)

    cropped = run(data)
    if verbose:
        print("**** New data shape is: {}".format(tuple(cropped.shape)))
        print("### END OV-CROP ###")
    if data_mask is not None:
        return cropped, run(data_mask), crop_coords
    return cropped, crop_coords


def merge_data_with_overlap(data, original_shape: Tuple[int, ...], data_mask=None, overlap: Tuple[float, ...] = (0, 0),
                            padding: Tuple[int, ...] = (0, 0), verbose: bool = True):
    """Merge ``(n_patches, y, x, c)`` patches into ``original_shape = (n_img, y, x, c)`` (reference ``:366-373``)."""
    if data_mask is not None:
        if tuple(data.shape[:-1]) != tuple(data_mask.shape[:-1]):
            raise ValueError("data and data_mask shapes mismatch: {} vs {}".format(tuple(data.shape[:-1]), tuple(data_mask.shape[:-1])))
    for i, p in enumerate(padding):
        if p >= data.shape[i + 1] // 2:
            raise ValueError(f"'Padding' cannot be greater than half of 'data' shape. Max value for this {tuple(data.shape)} input shape is "
                             f"{(data.shape[1] // 2) - 1, (data.shape[2] // 2) - 1}")
    _check_overlap(overlap)
    if verbose:
        print("### MERGE-OV-CROP ###")
        print(f"Merging {tuple(data.shape)} images into {original_shape} with smooth blending . . .")
        print(f"Overlap selected: {overlap}")
        print(f"Padding: {padding}")
    axes = [_stitch.Axis(original_shape[1 + i], data.shape[1 + i], padding[i], overlap[i]) for i in range(2)]
    sz, wz = _stitch.identity_axis(original_shape[0])
    starts = [sz, axes[0].starts(1), axes[1].starts(1)]
    wins = [wz, axes[0].window(), axes[1].window()]

    def run(arr):
        dt = str(arr.dtype).replace("torch.", "")
        if dt not in ("float32", "float16", "bfloat16"):
            raise TypeError(f"merge_data_with_overlap: unsupported dtype {arr.dtype} (float32/float16/bfloat16)")
        dev = _stitch.to_device(arr)
        out = _stitch.merge_device(dev[:, None], (original_shape[0], original_shape[1], original_shape[2]), starts, wins,
                                   (0, padding[0], padding[1]))
        return _stitch.like_input(out, arr)

    merged = run(data)
    if verbose:
        print(f"**** New data shape is: {tuple(merged.shape)}")
        print("### END MERGE-OV-CROP ###")
    if data_mask is not None:
        return merged, run(data_mask)
    return merged

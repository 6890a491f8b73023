"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the image normalisation at the ends of the path.

Follows ``biapy/data/norm.py``: normalize_image :44-220 (per-channel loop :188-218), percentile_clip :408-473 (value bounds
and np.percentile of the channel), norm_range01 :496-580, zero_mean_unit_variance_normalization :582-639, undo_image_norm :641-683, undo_norm_range01
:685-713, undo_zero_mean_unit_variance_normalization :715-780; and the binarisation of ``biapy/engine/semantic_seg.py:418-425,
524-531``.  Pinned against the reference's own functions by ``tests/test_oracle_golden.py`` / ``tests/golden/norm_*.npz``.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

EPS = 1e-6
_NP = {"uint8": np.uint8, "uint16": np.uint16, "float32": np.float32}


def is_binary(a: np.ndarray) -> bool:
    return bool(np.all((a == 0) | (a == 1)))


def _bounds(norm_module: Dict, key: str, c: int):
    v = norm_module[key]
    return [float(v[0])] * c if len(v) == 1 else list(v)


def normalize_image(img: np.ndarray, norm_module: Dict, apply_norm: bool = True):
    """norm.py:44-220 for value-given clip bounds and out_dtype float32."""
    orig = str(img.dtype)
    c = img.shape[-1]
    kind = norm_module["type"]
    pci = norm_module.get("per_channel_info")
    info = {"type": kind, "percentile_clip": norm_module["percentile_clip"], "orig_dtype": orig,
            "out_dtype": norm_module["out_dtype"], "per_channel_info": {}}
    img = img.astype(np.float32)
    for k in range(c):
        ch = {}
        d = img[..., k]
        if norm_module["percentile_clip"]:
            if pci is not None:
                lo, hi = pci[str(k)]["lower_bound_val"], pci[str(k)]["upper_bound_val"]
            else:
                # norm.py:445-466: a percentile of the (float32) channel unless it is None / -1, else the value given
                pl, pu = norm_module.get("per_lower_bound"), norm_module.get("per_upper_bound")
                lo = (float(np.percentile(d, pl)) if pl is not None and pl != -1
                      else _bounds(norm_module, "lower_bound_val", c)[k])
                hi = (float(np.percentile(d, pu)) if pu is not None and pu != -1
                      else _bounds(norm_module, "upper_bound_val", c)[k])
            if is_binary(d):
                lo, hi = 0.0, 1.0
            elif apply_norm:
                d = np.clip(d, lo, hi)
            ch["lower_bound_val"], ch["upper_bound_val"] = lo, hi
        if kind in ("div", "scale_range"):
            if is_binary(d):
                mx, mn = 1.0, 0.0
            else:
                if pci is not None and pci[str(k)].get("max_val_to_div") is not None:
                    mx, mn = float(pci[str(k)]["max_val_to_div"]), float(pci[str(k)]["min_val_to_div"])
                elif kind == "scale_range":
                    mx, mn = float(d.max()), float(d.min())
                else:
                    mx, mn = (65535 if d.max() > 255 else 255), 0
                if apply_norm:
                    d = (d - mn) / (max(mx - mn, EPS))
            ch["min_val_to_div"], ch["max_val_to_div"] = mn, mx
        else:
            mean = std = None
            if pci is not None:
                mean, std = pci[str(k)].get("mean"), pci[str(k)].get("std")
            else:
                if "mean" in norm_module and norm_module["mean"][0] != -1:
                    mean = _bounds(norm_module, "mean", c)[k]
                if "std" in norm_module and norm_module["std"][0] != -1:
                    std = _bounds(norm_module, "std", c)[k]
            if is_binary(d):
                m, s = 0.0, 1.0
            else:
                m = d.mean() if mean is None else mean
                s = d.std() if std is None else std
                if apply_norm:
                    d = (d - m) / (max(s, EPS))
            ch["mean"], ch["std"] = float(m), float(s)
        img[..., k] = d
        info["per_channel_info"][str(k)] = ch
    return img.astype(_NP[norm_module["out_dtype"]]), info


def undo_image_norm(data: np.ndarray, info: Dict) -> np.ndarray:
    """norm.py:641-780."""
    c = data.shape[-1]
    pci = info["per_channel_info"]
    if info["type"] in ("div", "scale_range"):
        data = np.clip(data, 0, 1)
        data = (data * [pci[str(k)]["max_val_to_div"] for k in range(c)]) + [pci[str(k)]["min_val_to_div"] for k in range(c)]
    else:
        data = (data * [pci[str(k)]["std"] for k in range(c)]) + [pci[str(k)]["mean"] for k in range(c)]
        if "float" not in str(info["orig_dtype"]):
            ii = np.iinfo(_NP[info["orig_dtype"]])
            data = np.clip(np.round(data), ii.min, ii.max)
    return data.astype(_NP[info["orig_dtype"]])


def threshold_otsu(image: np.ndarray, nbins: int = 256):
    """``skimage.filters.threshold_otsu(image)`` as ``after_merge_patches`` / ``after_full_image`` call it
    (``biapy/engine/semantic_seg.py:429, 455``) on the float32 merged prediction.

    PARITY UNPINNED against scikit-image itself: the package (``scikit-image>=0.21.0``, reference ``pyproject.toml:26``) is neither
    vendored in /root/reference nor installed in this image, so this restates its published algorithm
    (``skimage/filters/thresholding.py``: ``threshold_otsu`` -> ``_validate_image_histogram`` -> ``skimage.exposure.histogram``):
      1. one intensity value in the image -> return it;
      2. ``counts, edges = np.histogram(image.reshape(-1), bins=nbins)`` over the image's own [min, max] (float images);
         ``bin_centers = (edges[:-1] + edges[1:]) / 2``; counts cast to float32;
      3. class weights / means by cumulative sums from both ends, between-class variance
         ``w1[:-1] * w2[1:] * (m1[:-1] - m2[1:]) ** 2``, threshold = centre of its arg-max bin.
    Step 2 is numpy's own ``np.histogram`` (installed, exact); tests/golden/otsu_cases.npz holds what this function returns so a
    change of behaviour in a later numpy shows up."""
    first = image.reshape(-1)[0]
    if np.all(image == first):
        return first
    counts, edges = np.histogram(image.reshape(-1), bins=nbins)
    centers = (edges[:-1] + edges[1:]) / 2.0
    counts = counts.astype("float32", copy=False)
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * centers) / weight1
        mean2 = (np.cumsum((counts * centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return centers[np.argmax(variance12)]


def otsu_cases():
    """Seeded float32 predictions for the Otsu tests: sigmoid-like bimodal volume, flat noise, heavy ties, a constant image."""
    rng = np.random.default_rng(11)
    a = 1.0 / (1.0 + np.exp(-(rng.standard_normal((24, 40, 36, 1)) * 3.0 + np.where(rng.random((24, 40, 36, 1)) < 0.3, 4.0, -4.0))))
    b = rng.random((50, 70, 1))
    c = np.round(rng.random((16, 16, 16, 1)) * 7.0) / 7.0
    d = np.full((8, 8, 8, 1), 0.25)
    e = rng.standard_normal((33, 65, 1)) * 100.0 - 40.0
    return {k: v.astype(np.float32) for k, v in dict(bimodal=a, flat=b, ties=c, constant=d, wide=e).items()}


def binarize(pred: np.ndarray, n_classes: int, threshold: Optional[float] = 0.5) -> np.ndarray:
    """semantic_seg.py:418-425 (`threshold` None: Otsu of the whole prediction, as the reference; a number: that threshold) /
    :524-531 (by chunks: 0.5)."""
    if n_classes <= 2:
        if threshold is None:
            threshold = threshold_otsu(pred)
        return (pred > threshold).astype(np.uint8)
    return np.expand_dims(np.argmax(pred, -1), -1).astype(np.uint8 if n_classes < 255 else np.uint16)


def golden_cases(golden_dir: str):
    """(image, norm_module, golden normalised, golden undone, golden norm_info, key) from tests/golden/norm_cases.npz; images
    and modules are regenerated from oracle/make_golden.py's seeded definitions (numpy only)."""
    import json
    import os
    from .make_golden import NORM_MODULES, norm_images
    z = np.load(os.path.join(golden_dir, "norm_cases.npz"))
    imgs = norm_images()
    for m in json.loads(str(z["meta"])):
        key = f"{m['image']}{m['module']}"
        yield imgs[m["image"]], NORM_MODULES[m["module"]], z[key + "_y"], z[key + "_u"], m["info"], key

"""Image normalisation at the two ends of the inference path, on the device: mirror of ``normalize_image`` /
``undo_image_norm`` (``biapy/data/norm.py:44-220, 408-780``) with the reference's ``norm_module`` / ``norm_info`` dictionaries.

The statistics (min / max / mean / std / "is this channel binary") come from one pass of ``b200_image_stats``; the
normalisation itself is one float32 stream (``b200_image_norm_apply``: clip -> subtract -> divide in the reference's operation
order, so given the same statistics the result is bit-identical to numpy's).  Differences kept deliberately small and loud:

* percentile bounds computed from the data (``per_lower_bound`` / ``per_upper_bound``) are exact order statistics from a radix
  select on the device (``b200_select_hist``) combined on the host the way numpy (numpy input: ``np.percentile``, linear,
  float32) or the reference's ``torch_percentile`` (tensor input: ``kthvalue``) combine them; NaN-holding channels are not
  special-cased (numpy would answer NaN);
* ``out_dtype`` must be ``float32`` (what the engine consumes).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence, Tuple

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


# ------------------------------------------------------------------------------------------ percentiles from the data
# percentile_clip (norm.py:445-466) asks numpy for `np.percentile(channel, q)` (numpy input) or takes one order statistic with
# `kthvalue` (torch input, `torch_percentile` :475-497).  Both need exact order statistics of a whole channel: a radix select on
# the device (b200_select_hist: one histogram pass per 11-bit digit of an order-preserving key), then the same scalar
# arithmetic on the host that numpy / the reference do with the two neighbouring values.
_SELECT_PLAN = {torch.uint8: ((0, 11),), torch.uint16: ((11, 11), (0, 11)), torch.float32: ((21, 11), (10, 11), (0, 10))}


def _key_to_value(key: int, dtype: torch.dtype) -> np.float32:
    """Inverse of the device key mapping, as the float32 value the reference sees (it casts integer images to float32 first)."""
    if dtype != torch.float32:
        return np.float32(key)
    bits = (key ^ 0x80000000) if (key & 0x80000000) else (~key & 0xFFFFFFFF)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def select_keys(hist_fn: Callable[[int, int, int, int], np.ndarray], plan, ranks: Sequence[int]) -> Dict[int, int]:
    """Keys of the elements at the (0-based, ascending-order) `ranks`.  `hist_fn(shift, bits, prefix, has_prefix)` returns the
    digit histogram of the keys below `prefix`; ranks that share their leading digits share the passes."""
    out: Dict[int, int] = {}

    def solve(level: int, prefix: int, has: int, rel: List[Tuple[int, int]]):
        shift, bits = plan[level]
        hist = np.asarray(hist_fn(shift, bits, prefix, has), dtype=np.int64)
        cum = np.cumsum(hist)
        groups: Dict[int, List[Tuple[int, int]]] = {}
        for rank, r in rel:
            d = int(np.searchsorted(cum, r, side="right"))
            assert d < len(hist), "rank beyond the number of elements"
            groups.setdefault(d, []).append((rank, r - (int(cum[d - 1]) if d else 0)))
        for d, sub in groups.items():
            key = ((prefix << bits) | d) if has else d
            if level + 1 == len(plan):
                for rank, _ in sub:
                    out[rank] = key
            else:
                solve(level + 1, key, 1, sub)

    uniq = sorted(set(int(r) for r in ranks))
    solve(0, 0, 0, [(r, r) for r in uniq])
    return out


def _numpy_percentile_ranks(n: int, q: float):
    """The two neighbouring ranks and the weight of ``np.percentile(float32 array of n values, q)`` (method 'linear'): numpy
    divides q by float32(100) and keeps the virtual index (n - 1) * q in the array's dtype, float32 (numpy
    ``_function_base_impl.percentile`` / ``_quantile``)."""
    quant = np.true_divide(q, np.float32(100))
    vi = np.float32((n - 1) * quant)
    prev = int(np.floor(vi))
    gamma = np.float32(vi - np.floor(vi))
    if vi >= n - 1:
        return n - 1, n - 1, gamma
    if vi < 0:
        return 0, 0, gamma
    return prev, prev + 1, gamma


def _numpy_lerp(a: np.float32, b: np.float32, t: np.float32) -> float:
    """numpy's ``_lerp`` on float32 scalars: a + (b - a) * t, or b - (b - a) * (1 - t) from the midpoint on."""
    diff = np.float32(b - a)
    if t >= 0.5:
        return float(np.float32(b - np.float32(diff * np.float32(1 - t))))
    return float(np.float32(a + np.float32(diff * t)))


def channel_percentiles(hist_fn, plan, dtype: torch.dtype, n: int, qs: Sequence[float], torch_rule: bool) -> List[float]:
    """``[percentile(channel, q) for q in qs]`` with the reference's rule for the input type: ``np.percentile`` (linear
    interpolation in float32) for numpy images, ``kthvalue(1 + round(0.01 * q * (n - 1)))`` for tensors."""
    if n >= 2 ** 32:
        raise NotImplementedError("percentiles of a channel with 2^32 or more voxels: the select histogram counts in 32 bits")
    need, plans = [], []
    for q in qs:
        if torch_rule:
            k = 1 + round(0.01 * float(q) * (n - 1))
            plans.append((k - 1, k - 1, np.float32(0)))
        else:
            plans.append(_numpy_percentile_ranks(n, q))
        need += [plans[-1][0], plans[-1][1]]
    keys = select_keys(hist_fn, plan, need)
    res = []
    for lo, hi, gamma in plans:
        a, b = _key_to_value(keys[lo], dtype), _key_to_value(keys[hi], dtype)
        res.append(float(a) if lo == hi else _numpy_lerp(a, b, gamma))
    return res


def _device_hist_fn(img: torch.Tensor, ch: int):
    c = img.shape[-1]
    nvox = img.numel() // c
    buf = torch.empty(2048, dtype=torch.int32, device=img.device)

    def hist_fn(shift: int, bits: int, prefix: int, has_prefix: int) -> np.ndarray:
        ops._launch("b200_select_hist", ops._ptr(img), _lib.torch_dtype_code(img.dtype), nvox, c, ch, shift, bits, prefix, has_prefix,
                    ops._ptr(buf), _lib.stream_ptr())
        return buf[:1 << bits].cpu().numpy().view(np.uint32)
    return hist_fn


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
        if pci is None:
            # bounds from the data (norm.py:445-466): a percentile of -1 / None means "use the value given"
            pl, pu = norm_module.get("per_lower_bound"), norm_module.get("per_upper_bound")
            use_l, use_u = pl is not None and pl != -1, pu is not None and pu != -1
            if use_l:
                assert pl > 0, "Value in 'per_lower_bound' should be less than 100"
            if use_u:
                assert pu < 100, "Value in 'per_upper_bound' should be less than 100"
            if use_l or use_u:
                lo, hi = list(lo) if lo is not None else [None] * c, list(hi) if hi is not None else [None] * c
                plan, cont = _SELECT_PLAN[dev.dtype], dev.contiguous()
                for k in range(c):
                    qs = ([pl] if use_l else []) + ([pu] if use_u else [])
                    vals = channel_percentiles(_device_hist_fn(cont, k), plan, dev.dtype, cont.numel() // c, qs,
                                               torch_rule=isinstance(img, torch.Tensor))
                    if use_l:
                        lo[k] = vals[0]
                    if use_u:
                        hi[k] = vals[-1]
        if lo is None or hi is None or any(v is None for v in lo + hi):
            raise AssertionError("If 'per_lower_bound' / 'per_upper_bound' is not provided, 'lower_bound_val' / 'upper_bound_val' "
                                 "should be provided")
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


def otsu_from_counts(counts: np.ndarray, bin_edges: np.ndarray) -> np.floating:
    """The arithmetic of ``skimage.filters.threshold_otsu`` (scikit-image >= 0.21: ``filters/thresholding.py``) on a histogram:
    counts as float32 (``_validate_image_histogram``), bin centres = means of neighbouring edges (``exposure.histogram``), class
    weights / means by cumulative sums from both ends, threshold = centre of the bin that maximises the between-class variance."""
    counts = np.asarray(counts).astype("float32", copy=False)
    bin_centers = (bin_edges[:-1] + bin_edges[1:]) / 2.0
    weight1 = np.cumsum(counts)
    weight2 = np.cumsum(counts[::-1])[::-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        mean1 = np.cumsum(counts * bin_centers) / weight1
        mean2 = (np.cumsum((counts * bin_centers)[::-1]) / weight2[::-1])[::-1]
    variance12 = weight1[:-1] * weight2[1:] * (mean1[:-1] - mean2[1:]) ** 2
    return bin_centers[np.argmax(variance12)]


def threshold_otsu(pred, nbins: int = 256) -> np.floating:
    """``skimage.filters.threshold_otsu(pred)`` for the float32 prediction of the engine (``semantic_seg.py:429, 455``).  The two
    passes over the volume run on the device -- min / max (``b200_image_stats``) and numpy's equal-bin histogram over
    ``np.linspace(min, max, nbins + 1)`` (``b200_edge_hist``) --, the 256-element arithmetic is skimage's own sequence of numpy
    calls on the host, so the threshold is the one skimage returns for the same array."""
    dev = _stitch.to_device(pred)
    if dev.dtype != torch.float32:
        raise NotImplementedError("threshold_otsu takes the float32 prediction of the engine")
    flat = dev.contiguous().reshape(-1, 1)
    st = image_stats(flat)
    mn, mx = np.float32(st[0, 0]), np.float32(st[0, 1])
    if mn == mx:                                   # one intensity value: skimage returns it (first_pixel)
        return mn
    edges = np.linspace(mn, mx, nbins + 1, endpoint=True, dtype=np.result_type(mn, mx, np.float32))
    edges_d = torch.from_numpy(edges).to(flat.device)
    counts = torch.empty(nbins, dtype=torch.int64, device=flat.device)
    ops._launch("b200_edge_hist", ops._ptr(flat), flat.numel(), ops._ptr(edges_d), int(nbins), ops._ptr(counts), _lib.stream_ptr())
    return otsu_from_counts(counts.cpu().numpy(), edges)


def binarize_prediction(pred, n_classes: int, threshold: Optional[float] = 0.5):
    """Binarisation behind the merge (``semantic_seg.py:418-425, 524-531``): ``pred > threshold`` as uint8 for binary problems
    (`threshold` = None: the Otsu threshold of the whole prediction, as ``after_merge_patches`` / ``after_full_image`` do; 0.5: the
    by-chunks path), else the channel arg-max as uint8 (uint16 from 255 classes) with a trailing unit channel."""
    dev = _stitch.to_device(pred)
    if dev.dtype != torch.float32:
        raise NotImplementedError("binarize_prediction takes the float32 prediction of the engine")
    dev = dev.contiguous()
    if n_classes <= 2 and threshold is None:
        threshold = float(threshold_otsu(dev))
    if n_classes <= 2:
        out = torch.empty(dev.shape, dtype=torch.uint8, device=dev.device)
        ops._launch("b200_binarize", ops._ptr(dev), dev.numel(), float(threshold), ops._ptr(out), _lib.stream_ptr())
    else:
        dt = torch.uint8 if n_classes < 255 else torch.uint16
        out = torch.empty(tuple(dev.shape[:-1]) + (1,), dtype=dt, device=dev.device)
        ops._launch("b200_argmax_channels", ops._ptr(dev), dev.numel() // dev.shape[-1], dev.shape[-1], ops._ptr(out),
                    _lib.torch_dtype_code(dt), _lib.stream_ptr())
    return _stitch.like_input(out, pred)

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for BiaPy's sliding-window crop / overlap-add stitching.

A numpy restatement of the reference algorithm (BiaPy 3.7.0 @ 29539acd), written axis-generically so one
code path covers the 3D functions and their 2D twins.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the product
(``biapy_b200``) never does.

Reference sites restated here:
  * grid arithmetic      biapy/data/data_3D_manipulation.py:536-563 (crop), :777-811 (merge)
                         biapy/data/data_2D_manipulation.py:226-243 (crop), :464-484 (merge)
  * last-tile predicate  data_3D_manipulation.py:596-598 (crop), :826-828 (merge)
  * padding              data_3D_manipulation.py:505-516 (np.pad, 'zeros' -> 'constant')
  * spline window        data_3D_manipulation.py:664-688, data_2D_manipulation.py:336-364
  * accumulate/normalise data_3D_manipulation.py:838-849, data_2D_manipulation.py:504-517

Pinned against: the reference itself (oracle/ref_loader.py, tests/test_oracle_vs_reference.py, container
only), the reference's docstring known-answers (2600/390 patches 3D; 1980/3960/7920/3960 2D) and the
golden fixtures in tests/golden/ generated from the reference by oracle/make_golden.py.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np


class AxisPlan:
    """Grid of patches along one axis, exactly as the reference computes it (Python float/int semantics)."""

    __slots__ = ("dim", "patch", "pad", "step", "n", "last", "core", "ov_px")

    def __init__(self, dim: int, patch: int, pad: int, overlap: float):
        keep = 1 if overlap == 0 else 1 - overlap            # data_3D_manipulation.py:537-539
        core = patch - 2 * pad
        step = int(core * keep)                               # :542  (IEEE double product, truncation)
        n = math.ceil(dim / step)                             # :543  (ZeroDivisionError if step == 0, as the reference)
        last = 0 if n == 1 else ((n - 1) * step + patch) - (dim + 2 * pad)   # :544
        per_block = last // (n - 1) if n > 1 else 0          # :545
        step -= per_block                                     # :546
        last -= per_block * (n - 1)                           # :547
        self.dim, self.patch, self.pad, self.core = dim, patch, pad, core
        self.step, self.n, self.last = step, n, last
        self.ov_px = core - step                              # :814-816 (merge only)

    def crop_start(self, i: int) -> int:
        """Start of patch i in the *padded* frame (data_3D_manipulation.py:596-606)."""
        d = 0 if (i * self.step + self.patch) < (self.dim + 2 * self.pad) else self.last
        return i * self.step - d

    def merge_start(self, i: int) -> int:
        """Start of the un-padded core of patch i in the *original* frame (data_3D_manipulation.py:826-835)."""
        d = 0 if (i * self.step + self.core) < self.dim else self.last
        return i * self.step - d


def _check_overlap(overlap: Sequence[float]):
    for o in overlap:
        if o >= 1 or o < 0:
            raise ValueError("'overlap' values must be floats between range [0, 1)")


def pad_source_index(n_out: int, pad: int, dim: int, mode: str) -> np.ndarray:
    """For each index of the padded axis return the source index in the original axis (or -1 = constant 0)."""
    j = np.arange(n_out, dtype=np.int64) - pad
    if mode in ("constant", "zeros"):
        return np.where((j >= 0) & (j < dim), j, -1)
    if mode == "edge":
        return np.clip(j, 0, dim - 1)
    if mode == "reflect":          # d c b | a b c d | c b a
        if dim == 1:
            return np.zeros_like(j)
        period = 2 * (dim - 1)
        m = np.mod(j, period)
        return np.where(m < dim, m, period - m)
    if mode == "symmetric":        # c b a | a b c d | d c b
        period = 2 * dim
        m = np.mod(j, period)
        return np.where(m < dim, m, period - 1 - m)
    if mode == "wrap":
        return np.mod(j, dim)
    raise ValueError(f"unsupported pad_type {mode!r}")


def crop_grid(shape_sp: Sequence[int], patch_sp: Sequence[int], overlap, padding) -> Tuple[List[AxisPlan], np.ndarray]:
    """Return per-axis plans and an (n_patches, n_axes) int64 array of crop starts (padded frame), patch order = C order."""
    plans = [AxisPlan(int(d), int(p), int(q), o) for d, p, q, o in zip(shape_sp, patch_sp, padding, overlap)]
    per_axis = [np.array([pl.crop_start(i) for i in range(pl.n)], dtype=np.int64) for pl in plans]
    mesh = np.meshgrid(*per_axis, indexing="ij")
    return plans, np.stack([m.reshape(-1) for m in mesh], axis=1)


def merge_grid(orig_sp: Sequence[int], patch_sp: Sequence[int], overlap, padding) -> Tuple[List[AxisPlan], np.ndarray]:
    plans = [AxisPlan(int(d), int(p), int(q), o) for d, p, q, o in zip(orig_sp, patch_sp, padding, overlap)]
    per_axis = [np.array([pl.merge_start(i) for i in range(pl.n)], dtype=np.int64) for pl in plans]
    mesh = np.meshgrid(*per_axis, indexing="ij")
    return plans, np.stack([m.reshape(-1) for m in mesh], axis=1)


def crop_nd(data: np.ndarray, patch_shape: Sequence[int], overlap, padding, pad_type: str = "reflect"):
    """data (*spatial, C) -> patches (n, *patch_spatial, C), starts (n, n_axes).  Gather formulation of
    np.pad + strided copies (data_3D_manipulation.py:505-516, 591-623)."""
    nd = data.ndim - 1
    _check_overlap(overlap)
    plans, starts = crop_grid(data.shape[:nd], patch_shape[:nd], overlap, padding)
    mode = "constant" if pad_type == "zeros" else pad_type
    src = [pad_source_index(pl.dim + 2 * pl.pad, pl.pad, pl.dim, mode) for pl in plans]
    out = np.zeros((starts.shape[0],) + tuple(patch_shape[:nd]) + (data.shape[-1],), dtype=data.dtype)
    for c, st in enumerate(starts):
        idx = [src[a][st[a]: st[a] + plans[a].patch] for a in range(nd)]
        block = data
        valid = np.ones(tuple(len(i) for i in idx), dtype=bool)
        for a in range(nd):
            ia = idx[a]
            block = np.take(block, np.maximum(ia, 0), axis=a)
            shp = [1] * nd
            shp[a] = len(ia)
            valid &= (ia >= 0).reshape(shp)
        out[c] = np.where(valid[..., None], block, np.zeros((), dtype=data.dtype))
    return out, starts, plans


def spline_window_1d(size: int, ov_px: int, power: int = 2) -> np.ndarray:
    """data_3D_manipulation.py:664-672 (identical in data_2D_manipulation.py:336-345)."""
    w = np.ones(size, dtype=np.float32)
    if ov_px > 0:
        ov = min(ov_px, size // 2)
        x = np.linspace(0, 1, ov + 2)[1:-1]
        t = (x ** power) / (x ** power + (1 - x) ** power + 1e-8)
        w[:ov] = t
        w[-ov:] = t[::-1]
    return w


def spline_window_nd(core_shape: Sequence[int], ov_px: Sequence[int]) -> np.ndarray:
    """Outer product in float32, left to right (data_3D_manipulation.py:675-688)."""
    nd = len(core_shape)
    w = None
    for a in range(nd):
        shp = [1] * nd
        shp[a] = core_shape[a]
        wa = spline_window_1d(core_shape[a], ov_px[a]).reshape(shp)
        w = wa if w is None else w * wa
    return w[..., None].astype(np.float32)


def merge_nd(patches: np.ndarray, orig_shape: Sequence[int], overlap, padding) -> np.ndarray:
    """patches (n, *patch_spatial, C), orig_shape (*spatial, C) -> merged (*spatial, C) in patches.dtype.
    Scatter formulation, patch order = C order over the grid (data_3D_manipulation.py:822-849)."""
    nd = patches.ndim - 2
    _check_overlap(overlap)
    plans, starts = merge_grid(orig_shape[:nd], patches.shape[1:1 + nd], overlap, padding)
    core = tuple(slice(pl.pad, pl.patch - pl.pad) for pl in plans)
    data = patches[(slice(None),) + core]
    acc = np.zeros(tuple(orig_shape), dtype=np.float32)
    wsum = np.zeros(tuple(orig_shape[:nd]) + (1,), dtype=np.float32)
    win = spline_window_nd([pl.core for pl in plans], [pl.ov_px for pl in plans])
    for c, st in enumerate(starts):
        sl = tuple(slice(int(st[a]), int(st[a]) + plans[a].core) for a in range(nd))
        acc[sl] += data[c] * win
        wsum[sl] += win
    return np.true_divide(acc, wsum + 1e-18).astype(patches.dtype)


# ---------------------------------------------------------------------------------------------------------
# Reference-shaped wrappers (3D: data (z,y,x,c); 2D: data (n_img,y,x,c)) used by the parity tests
# ---------------------------------------------------------------------------------------------------------
def crop_3d(data, vol_shape, overlap=(0, 0, 0), padding=(0, 0, 0), pad_type="reflect"):
    out, starts, plans = crop_nd(data, vol_shape, overlap, padding, pad_type)
    return out, starts


def merge_3d(data, orig_vol_shape, overlap=(0, 0, 0), padding=(0, 0, 0)):
    return merge_nd(data, orig_vol_shape, overlap, padding)


def crop_2d(data, crop_shape, overlap=(0, 0), padding=(0, 0), pad_type="reflect"):
    """2D twin: every image of the stack is cropped independently, image-major order
    (data_2D_manipulation.py:266-300)."""
    outs, sts = [], []
    for z in range(data.shape[0]):
        o, s, _ = crop_nd(data[z], crop_shape, overlap, padding, pad_type)
        outs.append(o)
        sts.append(s)
    return np.concatenate(outs, 0), np.concatenate(sts, 0)


def merge_2d(data, original_shape, overlap=(0, 0), padding=(0, 0)):
    n_img = original_shape[0]
    per = data.shape[0] // n_img
    out = [merge_nd(data[z * per:(z + 1) * per], original_shape[1:], overlap, padding) for z in range(n_img)]
    return np.stack(out, 0)

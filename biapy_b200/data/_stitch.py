"""Device implementation shared by the 2D / 3D crop and merge mirrors.

The integer bookkeeping runs in the C planner (``b200_plan_axis`` / ``b200_axis_start``, host code, bit-exact
with the reference's Python float/int arithmetic); the data movement runs in two CUDA gather kernels
(``b200_crop_gather``, ``b200_overlap_add``).  2D stacks are handled as volumes whose z axis is the image
index with a z-patch of 1, so one kernel pair serves both families.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from .. import _lib


class Axis:
    """One axis of the patch grid (a thin view over the C planner's ``b200_axis_plan``)."""

    def __init__(self, dim: int, patch: int, pad: int, overlap: float):
        self.c = _lib.AxisPlan()
        st = _lib.lib().b200_plan_axis(int(dim), int(patch), int(pad), float(overlap), C.byref(self.c))
        if st != 0:
            msg = _lib.lib().b200_last_error().decode()
            if "division by zero" in msg:
                raise ZeroDivisionError("division by zero")          # what the reference raises (math.ceil(dim / 0))
            raise ValueError(msg)

    step = property(lambda s: s.c.step)
    n = property(lambda s: s.c.n)
    last = property(lambda s: s.c.last)
    core = property(lambda s: s.c.core)
    ov_px = property(lambda s: s.c.ov_px)
    patch = property(lambda s: s.c.patch)
    pad = property(lambda s: s.c.pad)
    dim = property(lambda s: s.c.dim)

    def starts(self, frame: int) -> np.ndarray:
        f = _lib.lib().b200_axis_start
        return np.array([f(C.byref(self.c), i, frame) for i in range(self.c.n)], dtype=np.int64)

    def window(self) -> np.ndarray:
        out = np.empty(self.core, dtype=np.float32)
        _lib.call("b200_spline_window_1d", self.core, self.ov_px, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out


def identity_axis(n: int) -> Tuple[np.ndarray, np.ndarray]:
    """z axis of a 2D stack: one 'patch' of extent 1 per image, no taper."""
    return np.arange(n, dtype=np.int64), np.ones(1, dtype=np.float32)


_DEV_TABLES: dict = {}


def _dev(a: np.ndarray, device):
    """Small host table (patch starts, spline windows) on the device; cached by content, so that the 54 crop / merge calls of a
    volume (and every call of a benchmark loop) do not each pay six pageable host-to-device copies."""
    import torch
    a = np.ascontiguousarray(a)
    key = (str(device), a.dtype.str, a.shape, a.tobytes())
    t = _DEV_TABLES.get(key)
    if t is None:
        if len(_DEV_TABLES) > 256:
            _DEV_TABLES.clear()
        t = _DEV_TABLES[key] = torch.from_numpy(a).to(device)
    return t


class VolumeShard:
    """The planes ``[z0, z0 + data.shape[0])`` of a ``depth``-plane volume ``(Z, Y, X, C)``: what one rank of the sharded
    sliding-window inference reads from disk / uploads (``planes_needed`` tells which)."""

    def __init__(self, data, z0: int, depth: int):
        self.data, self.z0, self.depth = data, int(z0), int(depth)
        if not (0 <= self.z0 and self.z0 + data.shape[0] <= self.depth):
            raise ValueError(f"shard planes [{self.z0}, {self.z0 + data.shape[0]}) outside a volume of {self.depth} planes")

    @property
    def shape(self):
        return (self.depth,) + tuple(self.data.shape[1:])


def _pad_src(j: np.ndarray, dim: int, mode: str) -> np.ndarray:
    """numpy.pad source coordinate for (possibly out-of-range) coordinates j; -1 = constant fill.  Host twin of `pad_src`."""
    j = np.asarray(j, dtype=np.int64)
    inside = (j >= 0) & (j < dim)
    if mode in ("zeros", "constant"):
        return np.where(inside, j, -1)
    if mode == "edge":
        return np.clip(j, 0, dim - 1)
    if mode == "reflect":
        if dim == 1:
            return np.zeros_like(j)
        m = np.mod(j, 2 * (dim - 1))
        return np.where(inside, j, np.where(m < dim, m, 2 * (dim - 1) - m))
    if mode == "symmetric":
        m = np.mod(j, 2 * dim)
        return np.where(inside, j, np.where(m < dim, m, 2 * dim - 1 - m))
    return np.where(inside, j, np.mod(j, dim))          # wrap


def planes_needed(depth: int, patch_z: int, pad_z: int, starts_z_crop: np.ndarray, n_yx: int, patch_range, pad_mode: str):
    """[z0, z1) of the volume planes the patches `patch_range` = (first, end) read, padding mode included."""
    first, end = patch_range
    if end <= first:
        return 0, 1
    rows = range(first // n_yx, (end - 1) // n_yx + 1)
    z = np.concatenate([_pad_src(int(starts_z_crop[iz]) - pad_z + np.arange(patch_z), depth, pad_mode) for iz in rows])
    z = z[z >= 0]
    return (int(z.min()), int(z.max()) + 1) if z.size else (0, 1)


def crop_device(vol, patch: Sequence[int], starts: Sequence[np.ndarray], pads: Sequence[int], pad_mode: str, patch_range=None):
    """vol: CUDA tensor (D,H,W,C) dense, or a :class:`VolumeShard` of one.  Returns (n_patches, pd, ph, pw, C) on the same
    device; with `patch_range = (first, end)` only those patches of the grid (C order over (z,y,x)), as (end - first, ...)."""
    import torch
    src_z0 = 0
    depth = None
    if isinstance(vol, VolumeShard):
        src_z0, depth, vol = vol.z0, vol.depth, vol.data
    _lib.require_cuda(vol, "crop input")
    vol = vol.contiguous()
    src_nz, H, W, Cc = vol.shape
    D = src_nz if depth is None else depth
    pd, ph, pw = (int(p) for p in patch)
    if pad_mode not in _lib.PAD_MODE:
        raise ValueError(f"unsupported pad_type {pad_mode!r}")
    n = len(starts[0]) * len(starts[1]) * len(starts[2])
    first, end = (0, n) if patch_range is None else (int(patch_range[0]), int(patch_range[1]))
    out = torch.empty((end - first, pd, ph, pw, Cc), dtype=vol.dtype, device=vol.device)
    if end == first:
        return out
    tabs = [_dev(s, vol.device) for s in starts]
    _lib.call("b200_crop_gather_range", vol.data_ptr(), _lib.torch_dtype_code(vol.dtype), D, H, W, Cc, out.data_ptr(),
              pd, ph, pw, tabs[0].data_ptr(), len(starts[0]), tabs[1].data_ptr(), len(starts[1]),
              tabs[2].data_ptr(), len(starts[2]), int(pads[0]), int(pads[1]), int(pads[2]),
              _lib.PAD_MODE[pad_mode], first, end - first, src_z0, src_nz, _lib.stream_ptr())
    return out


def merge_device(patches, out_shape: Sequence[int], starts: Sequence[np.ndarray], windows: Sequence[np.ndarray],
                 pads: Sequence[int], out_dtype=None, z_range=None, out=None):
    """patches: CUDA tensor (n, pz, py, px, C) dense incl. padding border.  Returns (D,H,W,C), or only the output planes
    `z_range = (z0, z1)` as (z1 - z0, H, W, C) -- then only the patch elements covering that slab are read."""
    import torch
    _lib.require_cuda(patches, "merge input")
    patches = patches.contiguous()
    n, pz, py, px, Cc = patches.shape
    D, H, W = (int(v) for v in out_shape[:3])
    assert n == len(starts[0]) * len(starts[1]) * len(starts[2]), (n, [len(s) for s in starts])
    out_dtype = out_dtype or patches.dtype
    z0, z1 = (0, D) if z_range is None else (int(z_range[0]), int(z_range[1]))
    if out is None:
        out = torch.empty((z1 - z0, H, W, Cc), dtype=out_dtype, device=patches.device)
    assert tuple(out.shape) == (z1 - z0, H, W, Cc) and out.is_contiguous() and out.dtype == out_dtype
    tabs = [_dev(s, patches.device) for s in starts]
    wins = [_dev(w, patches.device) for w in windows]
    _lib.call("b200_overlap_add_slab", patches.data_ptr(), _lib.torch_dtype_code(patches.dtype), out.data_ptr(),
              _lib.torch_dtype_code(out_dtype), D, H, W, Cc, pz, py, px, int(pads[0]), int(pads[1]), int(pads[2]),
              tabs[0].data_ptr(), len(starts[0]), tabs[1].data_ptr(), len(starts[1]), tabs[2].data_ptr(),
              len(starts[2]), wins[0].data_ptr(), wins[1].data_ptr(), wins[2].data_ptr(), z0, z1 - z0, _lib.stream_ptr())
    return out


_FLOAT_OK = ("float32", "float16")


def to_device(a, device=None):
    """numpy / torch -> CUDA tensor (no dtype change).  Fails loudly without a GPU."""
    import torch
    if isinstance(a, torch.Tensor):
        _lib.require_cuda(a)
        return a
    if not torch.cuda.is_available():
        raise _lib.B200Error("no CUDA device: biapy_b200 has no CPU path for crop/merge")
    return torch.from_numpy(np.ascontiguousarray(a)).to(device or "cuda")


def like_input(result, template):
    """Return `result` (CUDA tensor) as numpy if the caller passed numpy, else as a tensor."""
    if isinstance(template, np.ndarray):
        return result.cpu().numpy()
    return result

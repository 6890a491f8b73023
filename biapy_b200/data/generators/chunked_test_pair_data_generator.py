"""By-chunks tile generator on the B200: mirror of ``biapy.data.generators.chunked_test_pair_data_generator``
(``biapy/data/generators/chunked_test_pair_data_generator.py``) for volumes that are resident in HBM (numpy arrays are uploaded
once) AND for lazy array-likes -- anything with ``.shape`` and slice reads, i.e. a ``zarr.Array``, an ``h5py.Dataset`` or a
``numpy.memmap`` -- from which every tile is read with its halo exactly like the reference's
``extract_patch_from_efficient_file`` does (``biapy/data/data_3D_manipulation.py:179-351``): volumes larger than HBM stream through.

Same grid attributes (``step_z``, ``vols_per_z``, ``z_vol_start`` ..., ``total_vols``, ``tile_ids``, ``patches_of_tile``),
same methods (``_patch_coords``, ``tile_coords``, ``rank_workload``, ``extract_and_prepare_sample``,
``insert_patch_in_file``) and the same tile dealing as the reference's ``__iter__`` (``DistributedSampler(shuffle=False)``
over the sorted tile ids, ``:612-624``).  The integer bookkeeping runs in the C planner (``b200_chunk_grid_plan`` /
``b200_chunk_patch_coords``), the data movement in two CUDA kernels (``b200_chunk_extract`` / ``b200_chunk_insert``) that
handle a whole batch of tiles per launch.

Out of scope here: opening Zarr / H5 files (the caller hands over the opened dataset; neither package is needed by this module),
axis orders other than ``ZYXC``, ROI masks, sample filtering, normalisation modules -- they raise ``NotImplementedError``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ... import _lib
from ..dataset import PatchCoords


def _coords(v: Sequence[int]) -> PatchCoords:
    return PatchCoords(z_start=int(v[0]), z_end=int(v[1]), y_start=int(v[2]), y_end=int(v[3]), x_start=int(v[4]), x_end=int(v[5]))


class chunked_test_pair_data_generator:
    def __init__(self, sample_to_process: Dict, norm_module: Optional[Dict] = None, input_axes: str = "ZYXC",
                 mask_input_axes: str = "ZYXC", crop_shape: Tuple[int, ...] = (), padding: Tuple[int, ...] = (0, 0, 0),
                 out_dir: Optional[str] = None, dtype_str: str = "float32", convert_to_rgb: bool = False, z_start: int = -1,
                 z_end: int = -1, roi_mask_path: Optional[str] = None, patches_per_tile: Tuple[int, ...] = (1, 1, 1), **unsupported):
        import torch
        if input_axes != "ZYXC" or mask_input_axes != "ZYXC":
            raise NotImplementedError("the B200 by-chunks generator takes ZYXC volumes resident in memory")
        if roi_mask_path or convert_to_rgb or any(v for v in unsupported.values()):
            raise NotImplementedError(f"not implemented by the B200 by-chunks generator: roi_mask / convert_to_rgb / {sorted(unsupported)}")
        X = sample_to_process["X"]
        self.lazy = None
        if isinstance(X, np.ndarray) and not isinstance(X, np.memmap):
            if not torch.cuda.is_available():
                raise _lib.B200Error("no CUDA device: biapy_b200 has no CPU path for by-chunks extraction")
            X = torch.from_numpy(np.ascontiguousarray(X)).cuda()
        if not isinstance(X, torch.Tensor):
            # lazy array-like (zarr.Array / h5py.Dataset / numpy.memmap): tiles are read one by one with their halo
            if not (hasattr(X, "shape") and hasattr(X, "__getitem__")) or len(X.shape) != 4:
                raise NotImplementedError("by-chunks input must be a numpy array, a CUDA tensor or a (Z, Y, X, C) array-like with "
                                          "slice reads (zarr.Array, h5py.Dataset, numpy.memmap)")
            if not torch.cuda.is_available():
                raise _lib.B200Error("no CUDA device: biapy_b200 has no CPU path for by-chunks extraction")
            self.lazy = X
            self.device = torch.device("cuda", torch.cuda.current_device())
            self.bytes_read = 0
        else:
            _lib.require_cuda(X, "by-chunks volume")
            if X.dim() != 4:
                raise ValueError(f"expected a (Z, Y, X, C) volume, got shape {tuple(X.shape)}")
            self.device = X.device
        self.sample_to_process = sample_to_process
        self.X_parallel_data = None if self.lazy is not None else X.contiguous()
        self.filename = sample_to_process.get("X_filename", "")
        self.norm_module = norm_module
        self.input_axes, self.mask_input_axes, self.out_data_order = input_axes, mask_input_axes, input_axes
        self.dtype_str, self.out_dir = dtype_str, out_dir
        self.crop_shape = tuple(int(v) for v in crop_shape[:3]) + (int(X.shape[3]),)
        self.padding = tuple(int(v) for v in padding)
        self.z_dim, self.y_dim, self.x_dim = (int(v) for v in X.shape[:3])

        self.c = _lib.ChunkGrid()
        arr = lambda v: (C.c_int64 * 3)(*[int(t) for t in v[:3]])
        st = _lib.lib().b200_chunk_grid_plan(arr(tuple(X.shape)), arr(self.crop_shape), arr(self.padding), int(z_start), int(z_end), C.byref(self.c))
        if st != 0:
            raise ValueError(_lib.lib().b200_last_error().decode())           # the reference raises ValueError (:247-274)
        self.step_z, self.step_y, self.step_x = (int(v) for v in self.c.step)
        self.vols_per_z, self.vols_per_y, self.vols_per_x = (int(v) for v in self.c.vols)
        self.z_vol_start, self.z_vol_end = int(self.c.z_vol_start), int(self.c.z_vol_end)
        self.vols_per_z_effective = self.z_vol_end - self.z_vol_start
        self.total_vols = int(self.c.total)
        self.roi_mask = None
        self.vol_ids = list(range(self.total_vols))
        self.len = len(self.vol_ids)

        # tiles = groups of consecutive patches of the global grid (:331-360)
        self.patches_per_tile = tuple(max(1, int(x)) for x in patches_per_tile)
        self.tile_step = (self.step_z * self.patches_per_tile[0], self.step_y * self.patches_per_tile[1],
                          self.step_x * self.patches_per_tile[2])
        self.tiles_per_z = math.ceil(self.vols_per_z / self.patches_per_tile[0])
        self.tiles_per_y = math.ceil(self.vols_per_y / self.patches_per_tile[1])
        self.tiles_per_x = math.ceil(self.vols_per_x / self.patches_per_tile[2])
        self.patches_of_tile: Dict[int, List[int]] = {}
        for vol_id in self.vol_ids:
            x = vol_id % self.vols_per_x
            y = (vol_id // self.vols_per_x) % self.vols_per_y
            z = vol_id // (self.vols_per_x * self.vols_per_y) + self.z_vol_start
            tile_id = ((z // self.patches_per_tile[0]) * self.tiles_per_y + y // self.patches_per_tile[1]) * self.tiles_per_x \
                + x // self.patches_per_tile[2]
            self.patches_of_tile.setdefault(tile_id, []).append(vol_id)
        self.tile_ids = sorted(self.patches_of_tile.keys())
        self.out_data = None

    def __len__(self):
        return self.len

    # ------------------------------------------------------------------------------------------- bookkeeping
    def _raw(self, vol_id: int) -> List[int]:
        out = (C.c_int64 * 27)()
        _lib.call("b200_chunk_patch_coords", C.byref(self.c), int(vol_id), out)
        return list(out)

    def _patch_coords(self, vol_id: int) -> Tuple[int, int, int, PatchCoords, PatchCoords]:
        r = self._raw(vol_id)
        return r[0], r[1], r[2], _coords(r[3:9]), _coords(r[9:15])

    def tile_coords(self, tile_id: int) -> PatchCoords:
        x = tile_id % self.tiles_per_x
        y = (tile_id // self.tiles_per_x) % self.tiles_per_y
        z = tile_id // (self.tiles_per_x * self.tiles_per_y)
        z0, y0, x0 = z * self.tile_step[0], y * self.tile_step[1], x * self.tile_step[2]
        return _coords([z0, min(z0 + self.tile_step[0], self.z_dim), y0, min(y0 + self.tile_step[1], self.y_dim),
                        x0, min(x0 + self.tile_step[2], self.x_dim)])

    def rank_workload(self, num_workers: int, world_size: int, rank: int) -> Tuple[int, int]:
        workers = max(1, int(num_workers))
        replicas = workers * max(1, int(world_size))
        total = len(self.tile_ids)
        if total == 0:
            return 0, 0
        padded = math.ceil(total / replicas) * replicas
        order = [i % total for i in range(padded)]
        mine = set()
        for worker in range(workers):
            mine.update(order[rank * workers + worker:: replicas])
        return sum(len(self.patches_of_tile[self.tile_ids[i]]) for i in mine), len(mine)

    def rank_patches(self, world_size: int = 1, rank: int = 0, num_workers: int = 1, worker_id: int = 0, drop_repeats: bool = False) -> List[int]:
        """Patch ids in the order rank `rank` visits them in the reference's ``__iter__`` (``:612-624``): tiles are dealt by
        ``DistributedSampler(tile_ids, num_replicas, rank, shuffle=False)`` -- the index list is padded by wrapping around so
        that every replica gets ``ceil(n / replicas)`` tiles -- and the patches of a tile stay together."""
        from ...engine.dist import deal_tiles
        out: List[int] = []
        for i in deal_tiles(len(self.tile_ids), rank, world_size, num_workers, worker_id, drop_repeats):
            out += self.patches_of_tile[self.tile_ids[i]]
        return out

    # ------------------------------------------------------------------------------------------ device data path
    def extract_batch(self, vol_ids: Sequence[int]):
        """-> (patches CUDA tensor (n, *crop_shape), added_pad list, real coords list) for a batch of tiles, one launch."""
        import torch
        raws = [self._raw(v) for v in vol_ids]
        pads = [[[r[21], r[22]], [r[23], r[24]], [r[25], r[26]], [0, 0]] for r in raws]
        if self.lazy is not None:
            return self._extract_lazy(raws), pads, [_coords(r[9:15]) for r in raws]
        desc = np.array([[r[3], r[4] - r[3], r[15], r[5], r[6] - r[5], r[17], r[7], r[8] - r[7], r[19]] for r in raws], dtype=np.int64)
        X = self.X_parallel_data
        out = torch.empty((len(raws),) + self.crop_shape, dtype=X.dtype, device=X.device)
        d = torch.from_numpy(desc).to(X.device)
        _lib.call("b200_chunk_extract", X.data_ptr(), _lib.torch_dtype_code(X.dtype), self.z_dim, self.y_dim, self.x_dim, X.shape[3],
                  out.data_ptr(), len(raws), self.crop_shape[0], self.crop_shape[1], self.crop_shape[2], d.data_ptr(), _lib.stream_ptr())
        return out, pads, [_coords(r[9:15]) for r in raws]

    def _extract_lazy(self, raws):
        """Tiles of a lazy volume: read the effective region of every tile (halo included, clipped at the volume border: what
        ``extract_patch_from_efficient_file`` reads), upload it, and let the same extraction kernel reflect-pad it to the crop
        shape -- the box plays the role of the volume, so its descriptor starts at 0."""
        import torch
        first = np.array(self.lazy[0:1, 0:1, 0:1])                # a copy: memmaps opened read-only are not writable
        tdt = torch.from_numpy(first).dtype
        out = torch.empty((len(raws),) + self.crop_shape, dtype=tdt, device=self.device)
        for j, r in enumerate(raws):
            box = np.array(self.lazy[r[3]:r[4], r[5]:r[6], r[7]:r[8]], order="C")
            self.bytes_read += box.nbytes
            bd = torch.from_numpy(box).to(self.device, non_blocking=True)
            desc = torch.tensor([[0, r[4] - r[3], r[15], 0, r[6] - r[5], r[17], 0, r[8] - r[7], r[19]]], dtype=torch.int64, device=self.device)
            _lib.call("b200_chunk_extract", bd.data_ptr(), _lib.torch_dtype_code(bd.dtype), box.shape[0], box.shape[1], box.shape[2],
                      box.shape[3], out[j].data_ptr(), 1, self.crop_shape[0], self.crop_shape[1], self.crop_shape[2], desc.data_ptr(),
                      _lib.stream_ptr())
        return out

    def write_batch(self, pred, pads: Sequence, coords: Sequence[PatchCoords], out):
        """Lazy output (zarr.Array / h5py.Dataset / numpy.memmap, (Z, Y, X, C_out)): strip the padding of every predicted tile
        and assign it to its region -- ``insert_patch_in_efficient_file`` (``data_3D_manipulation.py:265-351``); the storage
        library does the chunking."""
        host = pred.cpu().numpy()
        for j, (c, p) in enumerate(zip(coords, pads)):
            pz, py, px = host.shape[1:4]
            tile = host[j, p[0][0]:pz - p[0][1], p[1][0]:py - p[1][1], p[2][0]:px - p[2][1]]
            if tile.shape[:3] != (c.z_end - c.z_start, c.y_end - c.y_start, c.x_end - c.x_start):
                raise ValueError("could not broadcast the stripped prediction into the output region")
            out[c.z_start:c.z_end, c.y_start:c.y_end, c.x_start:c.x_end] = tile

    def extract_and_prepare_sample(self, z: int, y: int, x: int, patch_coords: PatchCoords, extract: str = "image"):
        """Reference signature (``:489-575``): one tile -> (numpy patch of ``crop_shape``, pad_to_add)."""
        if extract != "image":
            raise NotImplementedError("mask extraction is not implemented by the B200 by-chunks generator")
        vol_id = ((z - self.z_vol_start) * self.vols_per_y + y) * self.vols_per_x + x
        data, pads, _ = self.extract_batch([vol_id])
        return data[0].cpu().numpy(), pads[0]

    def _ensure_out(self, channels: int, dtype):
        import torch
        if self.out_data is None:
            self.out_data = torch.zeros((self.z_dim, self.y_dim, self.x_dim, channels), dtype=dtype, device=self.device)
        return self.out_data

    def insert_batch(self, pred, pads: Sequence, coords: Sequence[PatchCoords], mode: str = "replace", out=None):
        """Strip the padding of every predicted tile (``base_workflow.py:2606-2612``) and write it at its coordinates
        (``insert_patch_in_efficient_file``), one launch for the batch.  pred: CUDA tensor (n, pz, py, px, C)."""
        import torch
        assert mode in ("add", "replace")
        _lib.require_cuda(pred, "by-chunks prediction")
        pred = pred.contiguous()
        n, pz, py, px, ch = pred.shape
        out = self._ensure_out(ch, torch.float32) if out is None else out
        desc = np.array([[c.z_start, c.z_end - c.z_start, p[0][0], c.y_start, c.y_end - c.y_start, p[1][0],
                          c.x_start, c.x_end - c.x_start, p[2][0]] for c, p in zip(coords, pads)], dtype=np.int64)
        for row, p in zip(desc, pads):       # what numpy would raise on: the stripped prediction must fill the region exactly
            if (pz - p[0][0] - p[0][1], py - p[1][0] - p[1][1], px - p[2][0] - p[2][1]) != (row[1], row[4], row[7]):
                raise ValueError("could not broadcast the stripped prediction into the output region")
        d = torch.from_numpy(desc).to(pred.device)
        _lib.call("b200_chunk_insert", pred.data_ptr(), _lib.torch_dtype_code(pred.dtype), n, pz, py, px, ch, out.data_ptr(),
                  _lib.torch_dtype_code(out.dtype), out.shape[0], out.shape[1], out.shape[2], d.data_ptr(), 1 if mode == "add" else 0,
                  _lib.stream_ptr())
        return out

    def insert_patch_in_file(self, patch, patch_coords: PatchCoords):
        """Reference signature (``:802-832``): `patch` is the already stripped prediction (numpy or CUDA, (z, y, x, C))."""
        import torch
        if isinstance(patch, np.ndarray):
            patch = torch.from_numpy(np.ascontiguousarray(patch)).to(self.device)
        zero = [[0, 0], [0, 0], [0, 0]]
        self.insert_batch(patch[None], [zero], [patch_coords], out=self._ensure_out(patch.shape[-1], patch.dtype))

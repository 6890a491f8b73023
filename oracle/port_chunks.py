"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, plain Python integers) of BiaPy's by-chunks tile grid.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product path
(``biapy_b200``) never does.  Pinned against the real reference class by ``oracle/make_golden.py`` ->
``tests/golden/chunks_*.npz`` / ``chunks.json`` (``tests/test_oracle_golden.py``).

Follows (reference = BiaPy 3.7.0, paths relative to ``/root/reference``):

* grid sizes            ``biapy/data/generators/chunked_test_pair_data_generator.py:276-295``
* tiles of patches      ``...:331-360``; ``tile_coords`` ``:380-405``; ``rank_workload`` ``:407-438``
* ``_patch_coords``     ``...:440-486``
* ``extract_and_prepare_sample`` (clip, reflect-pad to the crop shape, "real padding info")  ``...:489-575``
* tile dealing          ``...:612-624`` (``DistributedSampler(shuffle=False)`` over the sorted tile ids)
* strip + insert        ``biapy/engine/base_workflow.py:2606-2614`` and ``biapy/data/data_3D_manipulation.py:286-351``
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np


class ChunkGrid:
    def __init__(self, dims: Sequence[int], crop_shape: Sequence[int], padding: Sequence[int], z_start: int = -1,
                 z_end: int = -1, patches_per_tile: Sequence[int] = (1, 1, 1)):
        self.z_dim, self.y_dim, self.x_dim = (int(v) for v in dims[:3])
        self.crop_shape = tuple(int(v) for v in crop_shape)
        self.padding = tuple(int(v) for v in padding)
        for ax, name in enumerate("ZYX"):
            if self.crop_shape[ax] > (self.z_dim, self.y_dim, self.x_dim)[ax]:
                raise ValueError(f"{name} Axis problem: {self.crop_shape[ax]} greater than {(self.z_dim, self.y_dim, self.x_dim)[ax]}")
        for i, p in enumerate(self.padding):
            if p >= self.crop_shape[i] // 2:
                raise ValueError("'Padding' can not be greater than half of 'crop_shape'")
        self.step_z = self.crop_shape[0] - self.padding[0] * 2
        self.vols_per_z = math.ceil(self.z_dim / self.step_z)
        self.step_y = self.crop_shape[1] - self.padding[1] * 2
        self.vols_per_y = math.ceil(self.y_dim / self.step_y)
        self.step_x = self.crop_shape[2] - self.padding[2] * 2
        self.vols_per_x = math.ceil(self.x_dim / self.step_x)
        ez0 = 0 if z_start == -1 else z_start
        ez1 = self.z_dim if z_end == -1 else z_end
        self.z_vol_start = math.ceil(ez0 / self.step_z)
        self.z_vol_end = min(math.ceil(ez1 / self.step_z), self.vols_per_z)
        self.vols_per_z_effective = self.z_vol_end - self.z_vol_start
        self.total_vols = self.vols_per_z_effective * self.vols_per_y * self.vols_per_x
        self.vol_ids = list(range(self.total_vols))
        self.patches_per_tile = tuple(max(1, int(x)) for x in patches_per_tile)
        self.tile_step = (self.step_z * self.patches_per_tile[0], self.step_y * self.patches_per_tile[1],
                          self.step_x * self.patches_per_tile[2])
        self.tiles_per_z = math.ceil(self.vols_per_z / self.patches_per_tile[0])
        self.tiles_per_y = math.ceil(self.vols_per_y / self.patches_per_tile[1])
        self.tiles_per_x = math.ceil(self.vols_per_x / self.patches_per_tile[2])
        self.patches_of_tile: Dict[int, List[int]] = {}
        for vol_id in self.vol_ids:
            zl, y, x = self._unravel(vol_id)
            t = (((zl + self.z_vol_start) // self.patches_per_tile[0]) * self.tiles_per_y
                 + y // self.patches_per_tile[1]) * self.tiles_per_x + x // self.patches_per_tile[2]
            self.patches_of_tile.setdefault(t, []).append(vol_id)
        self.tile_ids = sorted(self.patches_of_tile.keys())

    def _unravel(self, vol_id: int) -> Tuple[int, int, int]:
        x = vol_id % self.vols_per_x
        y = (vol_id // self.vols_per_x) % self.vols_per_y
        z = vol_id // (self.vols_per_x * self.vols_per_y)
        return z, y, x

    def patch_coords(self, vol_id: int):
        """-> (z, y, x, extract[6], real[6]) with [z_start, z_end, y_start, y_end, x_start, x_end] lists."""
        zl, y, x = self._unravel(vol_id)
        z = zl + self.z_vol_start
        pos = (z, y, x)
        steps = (self.step_z, self.step_y, self.step_x)
        dims = (self.z_dim, self.y_dim, self.x_dim)
        ext, real = [], []
        for a in range(3):
            ext += [max(0, pos[a] * steps[a] - self.padding[a]), min((pos[a] + 1) * steps[a] + self.padding[a], dims[a])]
            real += [pos[a] * steps[a], min((pos[a] + 1) * steps[a], dims[a])]
        return z, y, x, ext, real

    def pad_to_add(self, pos: Sequence[int], ext: Sequence[int]):
        """-> (raw [[l, r]]*3 used for np.pad, "real padding info" [[l, r]]*3 returned to the caller)."""
        steps = (self.step_z, self.step_y, self.step_x)
        raw, info = [], []
        for a in range(3):
            lo = pos[a] * steps[a] - self.padding[a]
            left = abs(lo) if lo < 0 else 0
            right = self.crop_shape[a] - (ext[2 * a + 1] - ext[2 * a]) - left
            raw.append([left, right])
            info.append([max(left, self.padding[a]), max(right, self.padding[a])])
        return raw, info

    def extract(self, vol: np.ndarray, vol_id: int):
        """vol (Z, Y, X, C) -> (patch of crop_shape, real-padding info, real coords)."""
        z, y, x, ext, real = self.patch_coords(vol_id)
        data = vol[ext[0]:ext[1], ext[2]:ext[3], ext[4]:ext[5]]
        raw, info = self.pad_to_add((z, y, x), ext)
        data = np.pad(data, raw + [[0, 0]], "reflect")
        assert data.shape[:3] == self.crop_shape[:3], (data.shape, self.crop_shape)
        return data, info, real

    def tile_coords(self, tile_id: int):
        x = tile_id % self.tiles_per_x
        y = (tile_id // self.tiles_per_x) % self.tiles_per_y
        z = tile_id // (self.tiles_per_x * self.tiles_per_y)
        z0, y0, x0 = z * self.tile_step[0], y * self.tile_step[1], x * self.tile_step[2]
        return [z0, min(z0 + self.tile_step[0], self.z_dim), y0, min(y0 + self.tile_step[1], self.y_dim),
                x0, min(x0 + self.tile_step[2], self.x_dim)]

    def rank_workload(self, num_workers: int, world_size: int, rank: int) -> Tuple[int, int]:
        workers = max(1, int(num_workers))
        replicas = workers * max(1, int(world_size))
        total = len(self.tile_ids)
        if total == 0:
            return 0, 0
        padded = math.ceil(total / replicas) * replicas
        order = [i % total for i in range(padded)]
        mine = set()
        for w in range(workers):
            mine.update(order[rank * workers + w:: replicas])
        return sum(len(self.patches_of_tile[self.tile_ids[i]]) for i in mine), len(mine)

    def rank_patches(self, world_size: int, rank: int, workers: int = 1, worker: int = 0) -> List[int]:
        """vol ids in the order one (rank, worker) visits them: torch DistributedSampler(shuffle=False, drop_last=False)
        over the sorted tile ids (indices padded by wrapping, then strided), patches of a tile kept together."""
        replicas = workers * world_size
        r = rank * workers + worker
        n = len(self.tile_ids)
        num_samples = math.ceil(n / replicas)
        total = num_samples * replicas
        idx = list(range(n))
        pad = total - n
        if pad <= n:
            idx += idx[:pad]
        else:
            idx += (idx * math.ceil(pad / n))[:pad]
        out: List[int] = []
        for i in idx[r:total:replicas]:
            out += self.patches_of_tile[self.tile_ids[i]]
        return out


def strip_and_insert(out: np.ndarray, pred: np.ndarray, info, real, mode: str = "replace"):
    """base_workflow.py:2606-2614 + insert_patch_in_efficient_file(mode) for a ZYXC output array."""
    raw = pred[info[0][0]: pred.shape[0] - info[0][1], info[1][0]: pred.shape[1] - info[1][1],
               info[2][0]: pred.shape[2] - info[2][1]]
    sl = (slice(real[0], real[1]), slice(real[2], real[3]), slice(real[4], real[5]), slice(None))
    if mode == "replace":
        out[sl] = raw
    else:
        out[sl] += raw
    return out


def predict_by_chunks(vol: np.ndarray, grid: ChunkGrid, fn, out_channels: int, world_size: int = 1, rank: int = 0,
                      out: np.ndarray = None) -> np.ndarray:
    """The by-chunks loop for one rank with `fn(batch of 1 patch) -> prediction` as the model."""
    if out is None:
        out = np.zeros(vol.shape[:3] + (out_channels,), dtype=np.float32)
    seen = set()
    for vid in grid.rank_patches(world_size, rank):
        if vid in seen:                       # samples repeated to even out the ranks are predicted but not used
            continue
        seen.add(vid)
        patch, info, real = grid.extract(vol, vid)
        pred = fn(patch[None])[0]
        strip_and_insert(out, pred, info, real)
    return out

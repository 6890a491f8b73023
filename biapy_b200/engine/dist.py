"""Data-parallel plumbing of the hot path (one process per GPU, ``torch.distributed``):

* training: ONE all-reduce of the flat fp32 gradient buffer per step (NCCL over NVLink on the GPU box; the same code
  runs on ``gloo`` for the CPU tests) -- replaces DDP's bucketed hooks (``biapy/engine/base_workflow.py:951-958``);
* blended sliding-window inference: every rank predicts a contiguous range of the patch grid and OWNS one z slab of
  the output volume; the only traffic is the z-pieces of patch predictions that reach into another rank's slab (point-to-point
  over NVLink, contiguous chunks, no packing), after which each rank runs the bit-exact overlap-add on its slab alone;
* by-chunks inference: tiles are dealt round-robin to ranks exactly as the reference's generator deals them
  (``biapy/data/generators/chunked_test_pair_data_generator.py:613-618``).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allreduce_mean_(flat: torch.Tensor, group=None) -> float:
    """Sum-all-reduce `flat` in place; returns the factor (1/world) the caller folds into its optimiser kernel."""
    rank, world = world_info(group)
    if world > 1:
        dist.all_reduce(flat, group=group)
    return 1.0 / world


def deal_patch_range(n_patches: int, rank: int, world: int):
    """[first, end) of the patches rank `rank` predicts: contiguous in the grid's C order over (z, y, x), so a rank's patches sit
    in one or two z rows -- next to the output slab it owns."""
    return (n_patches * rank) // world, (n_patches * (rank + 1)) // world


def slab_range(depth: int, rank: int, world: int):
    """[z0, z1) of the output planes rank `rank` owns."""
    return (depth * rank) // world, (depth * (rank + 1)) // world


def plan_slab_exchange(starts_z, n_yx: int, core_z: int, pad_z: int, depth: int, world: int):
    """Who sends which z-piece of which patch prediction to whom.  `starts_z`: merge-frame z start of every z row of the grid
    (``Axis.starts(1)``), `n_yx`: patches per z row, `core_z` = patch depth - 2 * pad.  Returns a list of
    ``(patch, src_rank, dst_rank, a0, a1)``: planes ``[a0, a1)`` of the patch ARRAY (pad border included) cover output planes of
    `dst_rank`'s slab.  Deterministic and identical on every rank; pieces whose src == dst are left out (already in place)."""
    n = len(starts_z) * n_yx
    ops = []
    owners = [0] * n
    for r in range(world):
        lo, hi = deal_patch_range(n, r, world)
        for c in range(lo, hi):
            owners[c] = r
    for iz, s in enumerate(starts_z):
        s = int(s)
        for d in range(world):
            z0, z1 = slab_range(depth, d, world)
            l0, l1 = max(0, z0 - s), min(core_z, z1 - s)
            if l1 <= l0:
                continue
            for c in range(iz * n_yx, (iz + 1) * n_yx):
                if owners[c] != d:
                    ops.append((c, owners[c], d, l0 + pad_z, l1 + pad_z))
    return ops


def exchange_patch_slabs(pred_all: torch.Tensor, plan, rank: int, group=None) -> int:
    """Run the point-to-point plan on `pred_all` (n_patches, pz, py, px, C): a piece is a contiguous chunk on both sides, sent from
    the predicting rank's array straight into the same place of the slab owner's array.  Returns the bytes this rank received."""
    p2p, got = [], 0
    for c, src, dst, a0, a1 in plan:
        if src == rank:
            p2p.append(dist.P2POp(dist.isend, pred_all[c, a0:a1], dst if group is None else dist.get_global_rank(group, dst), group))
        elif dst == rank:
            p2p.append(dist.P2POp(dist.irecv, pred_all[c, a0:a1], src if group is None else dist.get_global_rank(group, src), group))
            got += pred_all[c, a0:a1].numel() * pred_all.element_size()
    if p2p:
        for w in dist.batch_isend_irecv(p2p):
            w.wait()
    return got


def deal_tiles(n_tiles: int, rank: int, world: int, num_workers: int = 1, worker_id: int = 0, drop_repeats: bool = False) -> List[int]:
    """Indices (into the sorted tile-id list) that one (rank, worker) visits in the reference's by-chunks ``__iter__``
    (``chunked_test_pair_data_generator.py:612-618``): ``DistributedSampler(tile_ids, num_replicas=workers*world,
    rank=rank*workers+worker, shuffle=False)`` -- the index list is padded by wrapping around so every replica gets
    ``ceil(n / replicas)`` tiles, then strided.  Repeated tiles are predicted but used once (``base_workflow.py:2582-2590``);
    `drop_repeats` leaves them out (every tile then has exactly one owner across the replicas)."""
    import math
    replicas = max(1, num_workers) * max(1, world)
    r = rank * max(1, num_workers) + worker_id
    if n_tiles == 0:
        return []
    total = math.ceil(n_tiles / replicas) * replicas
    idx = list(range(n_tiles))
    pad = total - n_tiles
    idx += idx[:pad] if pad <= n_tiles else (idx * math.ceil(pad / n_tiles))[:pad]
    if drop_repeats:      # the wrapped-around tail only evens out the replicas; its tiles already belong to an earlier position
        return [idx[p] for p in range(r, n_tiles, replicas)]
    return idx[r:total:replicas]

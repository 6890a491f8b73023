"""Data-parallel plumbing of the hot path (one process per GPU, ``torch.distributed``):

* training: ONE all-reduce of the flat fp32 gradient buffer per step (NCCL over NVLink on the GPU box; the same code
  runs on ``gloo`` for the CPU tests) -- replaces DDP's bucketed hooks (``biapy/engine/base_workflow.py:951-958``);
* inference: patches are dealt round-robin to ranks, as the reference's by-chunks generator deals tiles
  (``biapy/data/generators/chunked_test_pair_data_generator.py:613-618``), and the per-rank predictions are
  all-gathered once before the merge.
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


def deal_patches(n_patches: int, rank: int, world: int) -> List[int]:
    """Indices of the patches rank `rank` predicts (round-robin, every patch exactly once across ranks)."""
    return list(range(rank, n_patches, world))


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


def gather_patch_predictions(pred: torch.Tensor, n_patches: int, group=None) -> torch.Tensor:
    """`pred` (n_patches, ...) holds valid rows only for this rank's dealt patches; after the call every rank holds
    all rows.  One all_gather of ceil(n/world) rows per rank."""
    rank, world = world_info(group)
    if world == 1:
        return pred
    per = (n_patches + world - 1) // world
    mine = deal_patches(n_patches, rank, world)
    send = torch.zeros((per,) + tuple(pred.shape[1:]), dtype=pred.dtype, device=pred.device)
    if mine:
        send[: len(mine)] = pred[mine]
    gathered = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(gathered, send, group=group)
    for r in range(world):
        ids = deal_patches(n_patches, r, world)
        if ids:
            pred[ids] = gathered[r][: len(ids)]
    return pred

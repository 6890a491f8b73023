"""Device-resident sliding-window inference: crop -> batched forward -> head activation -> spline overlap-add.

Mirrors the core of ``Base_Workflow.process_test_sample`` (``biapy/engine/base_workflow.py:1944-1997``):
``crop_3D_data_with_overlap`` -> ``predict_batches_in_test`` (``:1696-1728``: ``model_call_func`` per
``TRAIN.BATCH_SIZE`` patches, head activations of ``apply_model_activations`` ``:1367-1470``) ->
``merge_3D_data_with_overlap``.  The reference moves every batch host->device->host and merges in numpy on rank 0;
here the volume is uploaded once, patches never leave HBM and the prediction buffer is written by the head kernel.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _lib, ops
from ..data import _stitch


def apply_head_activations(pred_cl: torch.Tensor, head_activations: Sequence[str], out: torch.Tensor) -> torch.Tensor:
    """Channels-last restatement of ``apply_model_activations`` for inference (training=False):
    ``ce_sigmoid`` -> sigmoid per channel, runs of ``ce_softmax`` -> softmax over the run, ``linear`` -> copy."""
    acts = [a.lower() for a in head_activations]
    C = pred_cl.shape[-1]
    assert len(acts) >= C, (acts, C)
    i = 0
    while i < C:
        a = acts[i]
        if a == "ce_softmax":
            j = i
            while j < C and acts[j] == "ce_softmax":
                j += 1
            if not (pred_cl.dtype == out.dtype):
                raise _lib.B200Error("softmax head activation needs matching dtypes")
            ops.softmax_channels(pred_cl, out, i, j)
            i = j
            continue
        name = {"ce_sigmoid": "sigmoid", "linear": "none"}.get(a, a)
        ops.scale_shift_act(pred_cl[..., i:i + 1], None, None, name, out[..., i:i + 1])
        i += 1
    return out


def shard_planes(vol_shape: Sequence[int], patch_shape: Sequence[int], overlap=(0, 0, 0), padding=(0, 0, 0),
                 pad_type: str = "reflect", rank: int = 0, world: int = 1):
    """[z0, z1): the planes of a (Z, Y, X, C) volume that rank `rank`'s patches read -- load those into a
    ``_stitch.VolumeShard`` and hand it to :func:`predict_volume` instead of the whole volume."""
    from . import dist as bd
    axes = [_stitch.Axis(vol_shape[i], patch_shape[i], padding[i], overlap[i]) for i in range(3)]
    n = axes[0].n * axes[1].n * axes[2].n
    return _stitch.planes_needed(int(vol_shape[0]), int(patch_shape[0]), int(padding[0]), axes[0].starts(0), axes[1].n * axes[2].n,
                                 bd.deal_patch_range(n, rank, world), pad_type)


class _GraphedBatch:
    """Forward + head activations of one full batch as a replayed CUDA graph (the ~90 launches per batch of the eager path leave
    gaps on the small deep levels).  Static input / output buffers; one instance per (batch shape, dtypes, activations), kept on
    the model and dropped with its packed-weight cache whenever the weights may change (`train()` / `load_state_dict()` / an
    optimiser step)."""

    def __init__(self, model, xb: torch.Tensor, acts, out_dtype):
        self.x = torch.empty_like(xb)
        self.x.copy_(xb)
        side = torch.cuda.Stream(device=xb.device)
        side.wait_stream(torch.cuda.current_stream(xb.device))
        with torch.cuda.stream(side):
            for _ in range(2):                                   # warm-up: packs into the eval-mode cache, allocator, attributes
                y = model(self.x.permute(0, 4, 1, 2, 3))
        torch.cuda.current_stream(xb.device).wait_stream(side)
        torch.cuda.synchronize(xb.device)
        self.out = torch.empty(tuple(xb.shape[:4]) + (y.shape[1],), dtype=out_dtype, device=xb.device)
        n0 = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            y = model(self.x.permute(0, 4, 1, 2, 3))
            apply_head_activations(y.permute(0, 2, 3, 4, 1), acts, self.out)
        self.launches = ops.LAUNCHES - n0

    def __call__(self, xb: torch.Tensor, dst: torch.Tensor):
        self.x.copy_(xb)
        self.graph.replay()
        ops.LAUNCHES += self.launches
        dst.copy_(self.out)


class _PipelinedUpload:
    """Host -> device copy of a pinned volume shard on a side stream, in plane order, cut where the batches of the sliding window
    first need more planes: the forward passes of batch i run while the planes of batch i + 1 ... arrive."""

    def __init__(self, shard: "_stitch.VolumeShard"):
        self.host = shard.data
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.dev = torch.empty(self.host.shape, dtype=self.host.dtype, device=self.device)
        self.shard = _stitch.VolumeShard(self.dev, shard.z0, shard.depth)
        self.z0 = shard.z0
        self.stream = torch.cuda.Stream(device=self.device)
        self.events = []

    def start(self, z_hi_per_batch):
        done = 0
        nz = self.host.shape[0]
        # the buffer may be a block the main stream just released: order the copies behind the work queued there, and tell the
        # allocator that the copy stream uses it
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self.dev.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            for i, z_hi in enumerate(z_hi_per_batch):
                hi = nz if i == len(z_hi_per_batch) - 1 else max(done, min(nz, int(z_hi) - self.z0))
                if hi > done:
                    self.dev[done:hi].copy_(self.host[done:hi], non_blocking=True)
                    done = hi
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.events.append(ev)

    def wait(self, batch_index: int):
        torch.cuda.current_stream(self.device).wait_event(self.events[batch_index])


def _graphed_batch(model, xb: torch.Tensor, acts, out_dtype) -> Optional[_GraphedBatch]:
    if os.environ.get("B200_INFER_GRAPH", "1") == "0" or model.training or not hasattr(model, "engine_dtype"):
        return None
    cache = model.__dict__.setdefault("_pack_cache", {})         # shares the life cycle of the eval-mode packed weights
    # the captured launches point at the packed weights of the moment: re-capture when any parameter was rewritten or re-homed
    stamp = tuple((p._version, p.data_ptr()) for p in model.parameters())
    key = ("infer_graph", tuple(xb.shape), xb.dtype, model.engine_dtype, tuple(acts), out_dtype, stamp)
    g = cache.get(key)
    if g is None:
        g = cache[key] = _GraphedBatch(model, xb, acts, out_dtype)
    return g


@torch.no_grad()
def predict_volume(model, vol, patch_shape: Sequence[int], overlap=(0, 0, 0), padding=(0, 0, 0), batch_size: int = 4,
                   head_activations: Optional[List[str]] = None, pad_type: str = "reflect", out_dtype=torch.float32,
                   rank: int = 0, world: int = 1, tta: bool = False, tta_mode: str = "mean", tta_group: str = "auto",
                   gather: str = "all", group=None, stats: Optional[dict] = None):
    """vol: (Z, Y, X, C) numpy array or CUDA tensor, or a ``_stitch.VolumeShard`` holding just the planes this rank's patches read
    (``shard_planes`` tells which).  Returns the merged prediction (Z, Y, X, C_out) with the same container type.

    world > 1: rank r crops and predicts only the patches ``[n r / world, n (r + 1) / world)`` of the grid and owns
    the output planes ``[Z r / world, Z (r + 1) / world)``.  After the forward passes the ranks exchange just the z-pieces of patch
    predictions that reach into a neighbour's slab (``dist.plan_slab_exchange``; for the 512^3 / 128^3 / 25 % grid on 8 ranks about
    0.4 GB received per rank instead of the 1.6 GB of an all-gather) and every rank merges its own slab with the same gather kernel
    -- the patch order and every float operation per output element are those of world = 1, so the result is bit-identical.
    `gather`: ``"all"`` every rank ends with the full volume (one broadcast per slab), ``"rank0"`` only rank 0 does (the others
    return None) -- the reference's rank-0-only result (``base_workflow.py:1552-1559``) --, ``"none"`` returns
    ``(slab, (z0, z1))`` and leaves the volume sharded.

    ``tta`` = ``TEST.AUGMENTATION``: every patch is predicted in the 16 orientations of ``ensemble_predictions`` (activated
    outputs are ensembled, as ``predict_batches_in_test`` does, ``base_workflow.py:1659-1672``); mode / group =
    ``TEST.AUGMENTATION_MODE`` / ``TEST.AUGMENTATION_GROUP``."""
    from . import dist as bd
    if gather not in ("all", "rank0", "none"):
        raise ValueError(f"gather must be 'all', 'rank0' or 'none', got {gather!r}")
    upload = None
    if (isinstance(vol, _stitch.VolumeShard) and isinstance(vol.data, torch.Tensor) and not vol.data.is_cuda and vol.data.is_pinned()
            and torch.cuda.is_available()):
        # pinned host planes: the upload is pipelined with the forward passes (see _PipelinedUpload); the result is a CUDA tensor
        is_np = False
        upload = _PipelinedUpload(vol)
        dev_vol = upload.shard
        device = dev_vol.data.device
    elif isinstance(vol, _stitch.VolumeShard):
        is_np = isinstance(vol.data, np.ndarray)
        dev_vol = _stitch.VolumeShard(_stitch.to_device(vol.data), vol.z0, vol.depth)
        device = dev_vol.data.device
    else:
        is_np = isinstance(vol, np.ndarray)
        dev_vol = _stitch.to_device(vol)
        device = dev_vol.device
    Z, Y, X, Cin = dev_vol.shape
    axes = [_stitch.Axis(dev_vol.shape[i], patch_shape[i], padding[i], overlap[i]) for i in range(3)]
    starts_c = [a.starts(0) for a in axes]
    n = len(starts_c[0]) * len(starts_c[1]) * len(starts_c[2])
    first, end = bd.deal_patch_range(n, rank, world)
    n_yx = len(starts_c[1]) * len(starts_c[2])
    if upload is None:
        patches = _stitch.crop_device(dev_vol, patch_shape[:3], starts_c, padding, pad_type, patch_range=(first, end))
    else:
        # every copy is queued now, in plane order, on the copy stream; a batch waits only for the planes its patches read
        upload.start([_stitch.planes_needed(Z, int(patch_shape[0]), int(padding[0]), starts_c[0], n_yx,
                                            (first + k, min(first + k + batch_size, end)), pad_type)[1]
                      for k in range(0, end - first, batch_size)])
        patches = None
    c_out = sum(model.output_channels)
    acts = head_activations or ["linear"] * c_out
    # the full-grid array: this rank writes its own patches, the exchange fills in the pieces of the others it needs
    pred = torch.empty((n,) + tuple(int(v) for v in patch_shape[:3]) + (c_out,), dtype=out_dtype, device=device)
    for bi, k in enumerate(range(0, end - first, batch_size)):
        if upload is None:
            xb = patches[k:k + batch_size]
        else:
            upload.wait(bi)
            xb = _stitch.crop_device(dev_vol, patch_shape[:3], starts_c, padding, pad_type,
                                     patch_range=(first + k, min(first + k + batch_size, end)))
        dst = pred[first + k:first + k + xb.shape[0]]
        if tta:
            from ..data.post_processing.post_processing import ensemble_predictions

            def call(b):
                yb = model(b.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
                return apply_head_activations(yb, acts, torch.empty(yb.shape, dtype=torch.float32, device=yb.device)).permute(0, 4, 1, 2, 3)
            for j in range(xb.shape[0]):
                yj = ensemble_predictions(xb[j], call, (0, 2, 3, 4, 1), (0, 4, 1, 2, 3), xb.device, 3, batch_size_value=batch_size,
                                          mode=tta_mode, group=tta_group).permute(0, 2, 3, 4, 1)
                dst[j:j + 1].copy_(yj)
        else:
            g = _graphed_batch(model, xb, acts, out_dtype) if (xb.shape[0] == batch_size and (end - first) >= 2 * batch_size) else None
            if g is not None:
                g(xb, dst)
            else:
                y = model(xb.permute(0, 4, 1, 2, 3))                          # (b, C_out, z, y, x) fp32 view of NDHWC
                apply_head_activations(y.permute(0, 2, 3, 4, 1), acts, dst)
    starts_m = [a.starts(1) for a in axes]
    wins = [a.window() for a in axes]
    if world == 1:
        merged = _stitch.merge_device(pred, (Z, Y, X), starts_m, wins, padding)
        return merged.cpu().numpy() if is_np else merged
    plan = bd.plan_slab_exchange(starts_m[0], len(starts_m[1]) * len(starts_m[2]), axes[0].core, int(padding[0]), Z, world)
    got = bd.exchange_patch_slabs(pred, plan, rank, group)
    if stats is not None:
        stats.update(patches=end - first, exchange_bytes_received=got)
    z0, z1 = bd.slab_range(Z, rank, world)
    if gather == "none":
        slab = _stitch.merge_device(pred, (Z, Y, X), starts_m, wins, padding, z_range=(z0, z1))
        return (slab.cpu().numpy() if is_np else slab), (z0, z1)
    if gather == "rank0" and rank != 0:
        slab = _stitch.merge_device(pred, (Z, Y, X), starts_m, wins, padding, z_range=(z0, z1))
        torch.distributed.send(slab, 0 if group is None else torch.distributed.get_global_rank(group, 0), group=group)
        return None
    full = torch.empty((Z, Y, X, c_out), dtype=out_dtype, device=pred.device)
    _stitch.merge_device(pred, (Z, Y, X), starts_m, wins, padding, z_range=(z0, z1), out=full[z0:z1])
    for r in range(world):
        a, b = bd.slab_range(Z, r, world)
        if b <= a:
            continue
        gr = r if group is None else torch.distributed.get_global_rank(group, r)
        if gather == "all":
            torch.distributed.broadcast(full[a:b], gr, group=group)
        elif r != 0:
            torch.distributed.recv(full[a:b], gr, group=group)
    return full.cpu().numpy() if is_np else full


@torch.no_grad()
def predict_by_chunks(model, vol, patch_shape: Sequence[int], padding=(0, 0, 0), batch_size: int = 4,
                      head_activations: Optional[List[str]] = None, out_dtype=torch.float32, rank: int = 0, world: int = 1,
                      z_start: int = -1, z_end: int = -1, patches_per_tile=(1, 1, 1), out: Optional[torch.Tensor] = None,
                      reduce: bool = True):
    """The reference's multi-GPU inference semantics (``TEST.BY_CHUNKS``, ``base_workflow.py:2559-2614``): non-blended tiles
    of ``patch - 2 * padding`` read with a halo, reflect-padded at the volume border, predicted in batches and written back
    without the halo.  Tiles are dealt to ranks exactly as the reference's ``DistributedSampler`` deals them; every rank
    writes only its own tiles (disjoint regions), so with ``reduce=True`` one NCCL all-reduce (sum of disjoint supports)
    leaves the full prediction on every rank -- the reference gets the same effect by writing into a shared Zarr file.
    vol: (Z, Y, X, C) numpy array or CUDA tensor; returns (Z, Y, X, C_out) in the same container type.

    Streaming: `vol` may be a lazy array-like -- ``zarr.Array``, ``h5py.Dataset``, ``numpy.memmap``: anything with
    ``.shape`` and slice reads -- and `out` a writable one of shape (Z, Y, X, C_out).  Every tile is then read with its halo from
    `vol`, predicted, stripped and assigned to its region of `out` (``extract_patch_from_efficient_file`` /
    ``insert_patch_in_efficient_file``, ``data_3D_manipulation.py:179-351``): neither the volume nor the prediction is ever
    resident as a whole, on the host or on the device, and the ranks write disjoint regions of the shared file (no collective).
    Returns `out`."""
    from ..data.generators.chunked_test_pair_data_generator import chunked_test_pair_data_generator
    lazy = not isinstance(vol, torch.Tensor) and (isinstance(vol, np.memmap) or not isinstance(vol, np.ndarray))
    if lazy:
        if out is None or isinstance(out, torch.Tensor):
            raise ValueError("streaming by-chunks inference writes into `out`: pass a writable (Z, Y, X, C_out) array-like "
                             "(zarr.Array, h5py.Dataset, numpy.memmap or numpy array)")
        gen = chunked_test_pair_data_generator(dict(X=vol, Y=None, X_filename="", X_dir=""), None, "ZYXC", "ZYXC", tuple(patch_shape),
                                               tuple(padding), z_start=z_start, z_end=z_end, patches_per_tile=patches_per_tile)
        c_out = sum(model.output_channels)
        acts = head_activations or ["linear"] * c_out
        if tuple(out.shape) != (gen.z_dim, gen.y_dim, gen.x_dim, c_out):
            raise ValueError(f"`out` must have shape {(gen.z_dim, gen.y_dim, gen.x_dim, c_out)}, got {tuple(out.shape)}")
        todo = gen.rank_patches(world, rank, drop_repeats=True)
        for k in range(0, len(todo), batch_size):
            xb, pads, coords = gen.extract_batch(todo[k:k + batch_size])
            y = model(xb.permute(0, 4, 1, 2, 3))
            ycl = y.permute(0, 2, 3, 4, 1)
            act = torch.empty(ycl.shape, dtype=out_dtype, device=ycl.device)
            apply_head_activations(ycl, acts, act)
            gen.write_batch(act, pads, coords, out)
        if world > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.barrier()                      # the reference's only synchronisation on this path (:2614)
        return out
    is_np = isinstance(vol, np.ndarray)
    gen = chunked_test_pair_data_generator(dict(X=_stitch.to_device(vol), Y=None, X_filename="", X_dir=""), None, "ZYXC", "ZYXC",
                                           tuple(patch_shape), tuple(padding), z_start=z_start, z_end=z_end,
                                           patches_per_tile=patches_per_tile)
    c_out = sum(model.output_channels)
    acts = head_activations or ["linear"] * c_out
    dev = gen.X_parallel_data.device
    if out is None:
        out = torch.zeros((gen.z_dim, gen.y_dim, gen.x_dim, c_out), dtype=out_dtype, device=dev)
    # tiles the sampler repeats to even out the ranks are "predicted but not used" in the reference (:2582-2590): skip them,
    # so that every tile has one owner and the final all-reduce adds disjoint supports
    todo = gen.rank_patches(world, rank, drop_repeats=True)
    for k in range(0, len(todo), batch_size):
        xb, pads, coords = gen.extract_batch(todo[k:k + batch_size])
        y = model(xb.permute(0, 4, 1, 2, 3))
        ycl = y.permute(0, 2, 3, 4, 1)
        act = torch.empty(ycl.shape, dtype=out_dtype, device=dev)
        apply_head_activations(ycl, acts, act)
        gen.insert_batch(act, pads, coords, out=out)
    if world > 1 and reduce:
        torch.distributed.all_reduce(out)
    return out.cpu().numpy() if is_np else out

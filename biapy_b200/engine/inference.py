"""Device-resident sliding-window inference: crop -> batched forward -> head activation -> spline overlap-add.

Mirrors the core of ``Base_Workflow.process_test_sample`` (``biapy/engine/base_workflow.py:1944-1997``):
``crop_3D_data_with_overlap`` -> ``predict_batches_in_test`` (``:1696-1728``: ``model_call_func`` per
``TRAIN.BATCH_SIZE`` patches, head activations of ``apply_model_activations`` ``:1367-1470``) ->
``merge_3D_data_with_overlap``.  The reference moves every batch host->device->host and merges in numpy on rank 0;
here the volume is uploaded once, patches never leave HBM and the prediction buffer is written by the head kernel.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _lib, ops
from ..data import _stitch
from .dist import deal_patches, gather_patch_predictions


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


@torch.no_grad()
def predict_volume(model, vol, patch_shape: Sequence[int], overlap=(0, 0, 0), padding=(0, 0, 0), batch_size: int = 4,
                   head_activations: Optional[List[str]] = None, pad_type: str = "reflect", out_dtype=torch.float32,
                   rank: int = 0, world: int = 1, tta: bool = False, tta_mode: str = "mean", tta_group: str = "auto"):
    """vol: (Z, Y, X, C) numpy array or CUDA tensor.  Returns the merged prediction (Z, Y, X, C_out) with the same
    container type.  With world > 1 every rank predicts the patches ``rank::world`` (embarrassingly parallel, the
    reference's by-chunks dealing, ``chunked_test_pair_data_generator.py:613-618``) and the patch predictions are
    all-gathered over NCCL before each rank merges (every rank ends with the full volume).
    ``tta`` = ``TEST.AUGMENTATION``: every patch is predicted in the 16 orientations of ``ensemble_predictions`` (activated
    outputs are ensembled, as ``predict_batches_in_test`` does, ``base_workflow.py:1659-1672``); mode / group =
    ``TEST.AUGMENTATION_MODE`` / ``TEST.AUGMENTATION_GROUP``."""
    is_np = isinstance(vol, np.ndarray)
    dev_vol = _stitch.to_device(vol)
    Z, Y, X, Cin = dev_vol.shape
    axes_c = [_stitch.Axis(dev_vol.shape[i], patch_shape[i], padding[i], overlap[i]) for i in range(3)]
    starts = [a.starts(0) for a in axes_c]
    patches = _stitch.crop_device(dev_vol, patch_shape[:3], starts, padding, pad_type)      # (n, pz, py, px, Cin)
    n = patches.shape[0]
    c_out = sum(model.output_channels)
    acts = head_activations or ["linear"] * c_out
    pred = torch.empty((n,) + tuple(patches.shape[1:4]) + (c_out,), dtype=out_dtype, device=patches.device)
    mine = deal_patches(n, rank, world) if world > 1 else None
    idx = range(0, n, batch_size) if world == 1 else range(0, len(mine), batch_size)
    for k in idx:
        if world == 1:
            xb = patches[k:k + batch_size]
            sel = slice(k, k + batch_size)
        else:
            ids = torch.tensor(mine[k:k + batch_size], device=patches.device)
            xb = patches.index_select(0, ids)
            sel = ids
        if tta:
            from ..data.post_processing.post_processing import ensemble_predictions

            def call(b):
                yb = model(b.permute(0, 4, 1, 2, 3)).permute(0, 2, 3, 4, 1)
                return apply_head_activations(yb, acts, torch.empty(yb.shape, dtype=torch.float32, device=yb.device)).permute(0, 4, 1, 2, 3)
            ycl = torch.cat([ensemble_predictions(xb[j], call, (0, 2, 3, 4, 1), (0, 4, 1, 2, 3), xb.device, 3, batch_size_value=batch_size,
                                                  mode=tta_mode, group=tta_group).permute(0, 2, 3, 4, 1) for j in range(xb.shape[0])], 0)
        else:
            y = model(xb.permute(0, 4, 1, 2, 3))                              # (b, C_out, z, y, x) fp32 view of NDHWC
            ycl = y.permute(0, 2, 3, 4, 1)
        if tta:
            if world == 1:
                pred[sel] = ycl.to(out_dtype)
            else:
                pred.index_copy_(0, sel, ycl.to(out_dtype))
        elif world == 1:
            apply_head_activations(ycl, acts, pred[sel])
        else:
            tmp = torch.empty(ycl.shape, dtype=out_dtype, device=ycl.device)
            apply_head_activations(ycl, acts, tmp)
            pred.index_copy_(0, sel, tmp)
    if world > 1:
        gather_patch_predictions(pred, n)          # one NCCL all_gather, then every rank merges locally
    axes_m = [_stitch.Axis(dev_vol.shape[i], patch_shape[i], padding[i], overlap[i]) for i in range(3)]
    merged = _stitch.merge_device(pred, (Z, Y, X), [a.starts(1) for a in axes_m], [a.window() for a in axes_m], padding)
    return merged.cpu().numpy() if is_np else merged


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
    vol: (Z, Y, X, C) numpy array or CUDA tensor; returns (Z, Y, X, C_out) in the same container type."""
    from ..data.generators.chunked_test_pair_data_generator import chunked_test_pair_data_generator
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

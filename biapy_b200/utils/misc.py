"""Host <-> device layout helpers with the reference's signatures (``biapy/utils/misc.py:689-733``).

BiaPy keeps images as ``(N, [Z,] Y, X, C)`` numpy arrays and only *permutes* them into ``(N, C, [Z,] Y, X)``
tensors, so the tensor a model receives already has channels-last strides -- exactly the layout the B200 engine
computes in.  No data is re-ordered at the boundary."""
from __future__ import annotations

from typing import Tuple

import torch


def to_pytorch_format(x, axes_order: Tuple, device, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(dtype).permute(axes_order).to(device, non_blocking=True)
    return torch.from_numpy(x).to(dtype).permute(axes_order).to(device, non_blocking=True)


def to_numpy_format(x: torch.Tensor, axes_order_back: Tuple):
    return x.permute(axes_order_back).cpu().numpy()


# ------------------------------------------------------------------------------------------------ checkpoints
# Checkpoint interop with BiaPy (``biapy/utils/misc.py:328-660``): the model classes keep the reference's ``state_dict`` keys,
# so a BiaPy ``.pth`` / ``.safetensors`` file loads with ``strict=True`` and files written here load in BiaPy.
import collections.abc as _abc
import glob as _glob
import os as _os
import pickle as _pickle
from pathlib import Path as _Path


def _cfg_get(cfg, dotted: str, default=None):
    cur = cfg
    for part in dotted.split("."):
        if isinstance(cur, _abc.Mapping):
            if part not in cur:
                return default
            cur = cur[part]
        elif hasattr(cur, part):
            cur = getattr(cur, part)
        else:
            return default
    return cur


def is_main_process() -> bool:
    return not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_rank() == 0


def cfg_to_plain_dict(cfg):
    """Nested mapping (yacs ``CfgNode`` included) -> plain containers, so ``torch.load(weights_only=True)`` accepts the file
    (reference ``misc.py:413-424``)."""
    if isinstance(cfg, _abc.Mapping):
        return {k: cfg_to_plain_dict(v) for k, v in cfg.items()}
    if isinstance(cfg, (list, tuple)):
        return type(cfg)(cfg_to_plain_dict(v) for v in cfg)
    return cfg


def save_on_master(model_dict, checkpoint_path):
    """``.pth`` = the whole dict through ``torch.save``; ``.safetensors`` = the model tensors only (reference ``:389-410``)."""
    if not is_main_process():
        return
    path = str(checkpoint_path)
    if path.endswith(".pth"):
        torch.save(model_dict, checkpoint_path)
    elif path.endswith(".safetensors"):
        from safetensors.torch import save_file
        save_file({k: v.contiguous() for k, v in model_dict["model"].items()}, checkpoint_path)
    else:
        raise ValueError("Unsupported checkpoint extension: {}".format(checkpoint_path))


def save_model(output_dir, cfg, biapy_version, jobname, epoch, model_without_ddp, optimizer, model_build_kwargs=None, extension="pth"):
    """Write ``<jobname>-checkpoint-<epoch>.<extension>`` with the reference's dictionary layout (``misc.py:328-386``):
    ``model_build_kwargs, model, optimizer (list of state dicts), epoch, cfg (plain dict), biapy_version``.  `optimizer`: a list
    of objects with ``state_dict()`` -- torch optimisers or :class:`biapy_b200.engine.train.Trainer` (AdamW-compatible state)."""
    path = _Path(output_dir) / "{}-checkpoint-{}.{}".format(jobname, str(epoch), extension)
    to_save = {
        "model_build_kwargs": model_build_kwargs,
        "model": {k: v.detach().cpu() for k, v in model_without_ddp.state_dict().items()},
        "optimizer": [opt.state_dict() for opt in optimizer],
        "epoch": epoch,
        "cfg": cfg_to_plain_dict(cfg),
        "biapy_version": biapy_version,
    }
    save_on_master(to_save, path)
    return path


def load_checkpoint_file(path, map_location="cpu"):
    """``torch.load`` with ``weights_only=True`` first, full unpickling for old files that embed a ``CfgNode``
    (reference ``misc.py:427-460``)."""
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except _pickle.UnpicklingError as e:
        print("Checkpoint '{}' could not be loaded with 'weights_only=True' ({}). It was probably created with an older BiaPy "
              "version, so it will be loaded with 'weights_only=False'. Only do this with checkpoints coming from a trusted "
              "source.".format(path, e))
        return torch.load(path, map_location=map_location, weights_only=False)


def get_checkpoint_path(cfg, jobname):
    """Checkpoint path without extension from ``PATHS.CHECKPOINT_FILE`` or ``MODEL.LOAD_CHECKPOINT_EPOCH``
    (``last_on_train`` / ``best_on_val``), reference ``misc.py:463-513``."""
    ckpt_dir = _Path(_cfg_get(cfg, "PATHS.CHECKPOINT", "."))
    explicit = _cfg_get(cfg, "PATHS.CHECKPOINT_FILE", "")
    if explicit != "":
        return _os.path.splitext(explicit)[0]
    which = _cfg_get(cfg, "MODEL.LOAD_CHECKPOINT_EPOCH", "best_on_val")
    if which == "last_on_train":
        latest = -1
        for f in _glob.glob(_os.path.join(ckpt_dir, "{}-checkpoint-*".format(jobname))):
            t = f.split("-")[-1].split(".")[0]
            if t.isdigit():
                latest = max(int(t), latest)
        if latest < 0:
            raise FileNotFoundError("no '{}-checkpoint-<epoch>' file in {}".format(jobname, ckpt_dir))
        return _os.path.join(ckpt_dir, "{}-checkpoint-{}".format(jobname, latest))
    if which == "best_on_val":
        return _os.path.join(ckpt_dir, "{}-checkpoint-best".format(jobname))
    raise NotImplementedError


def load_model_checkpoint(cfg, jobname, model_without_ddp, device, optimizer=None, just_extract_checkpoint_info=False,
                          skip_unmatched_layers=False):
    """Mirror of the reference loader (``misc.py:516-660``): finds ``.pth`` / ``.safetensors``, accepts the ``model`` /
    ``model_state_dict`` / ``state_dict`` / bare layouts, loads strictly (or shape-filtered with `skip_unmatched_layers`),
    restores optimiser state and epoch when ``MODEL.ITEMS_TO_LOAD_FROM_CHECKPOINT`` lists them.  Returns
    ``(start_epoch, path)`` -- or ``(cfg dict | None, biapy_version | None)`` with `just_extract_checkpoint_info`."""
    resume = get_checkpoint_path(cfg, jobname)
    for ext in (".pth", ".safetensors"):
        if _os.path.exists(resume + ext):
            resume += ext
            break
    if not _os.path.exists(resume):
        raise FileNotFoundError(f"Checkpoint file {resume} not found (considering .pth and .safetensors extensions)")
    print(("Extracting model from checkpoint file {}" if just_extract_checkpoint_info else "Loading checkpoint from file {}").format(resume))
    if resume.endswith(".safetensors"):
        from safetensors.torch import load_file
        checkpoint = {"model": load_file(resume, device="cpu")}
    else:
        checkpoint = load_checkpoint_file(resume, map_location="cpu")
    if just_extract_checkpoint_info:
        return (checkpoint.get("cfg"), str(checkpoint["biapy_version"]) if "biapy_version" in checkpoint else None)
    for key in ("model", "model_state_dict", "state_dict"):
        if key in checkpoint:
            state = checkpoint[key]
            break
    else:
        state = checkpoint
    if not skip_unmatched_layers:
        model_without_ddp.load_state_dict(state, strict=True)
    else:
        own = model_without_ddp.state_dict()
        keep = {}
        for k, v in state.items():
            if k not in own:
                print(f"Skipping unexpected layer '{k}' not found in model.")
            elif not torch.is_tensor(v):
                print(f"Skipping layer '{k}' because its value is not a tensor (type {type(v)})")
            elif v.shape != own[k].shape:
                print(f"Skipping layer '{k}' due to shape mismatch: checkpoint {v.shape} vs model {own[k].shape}")
            else:
                keep[k] = v
        model_without_ddp.load_state_dict(keep, strict=False)
    print("Model weights loaded!")
    items = _cfg_get(cfg, "MODEL.ITEMS_TO_LOAD_FROM_CHECKPOINT", ["model"])
    if "optimizer" in checkpoint and optimizer is not None and "optimizer" in items:
        saved = checkpoint["optimizer"]
        if isinstance(saved, dict):
            saved = [saved]
        n = 0
        for opt, st in zip(optimizer, saved):
            opt.load_state_dict(st)
            n += 1
        print(f"Optimizer info loaded for {n}/{len(optimizer)} optimizer(s)!")
    start_epoch = 0
    if "epoch" in checkpoint and "epoch" in items:
        start_epoch = checkpoint["epoch"]
        start_epoch = 0 if isinstance(start_epoch, str) else int(start_epoch)
        print("Epoch loaded!")
    return start_epoch, resume

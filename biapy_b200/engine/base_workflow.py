"""Workflow plugin surface of the hot path: the part of ``biapy/engine/base_workflow.py`` that sits directly on the model and the
stitching kernels, with the reference's class / method names so a BiaPy workflow subclass keeps working.

What is here (reference line numbers): constructor contract ``(cfg, job_identifier, device, system_dict, args)`` ``:130-137``,
normalisation module ``:374-394``, ``define_activations_and_channels`` ``:477-520`` (abstract hook), ``prepare_model`` ``:906-995``,
``model_call_func`` ``:832-902``, ``apply_model_activations`` ``:1367-1470``, ``predict_batches_in_test`` ``:1632-1730``,
``process_test_sample`` ``:1874-2013`` (crop -> predict -> merge -> ``after_merge_patches``), and a training loop body on
:class:`biapy_b200.engine.train.Trainer` (``train_engine.py:106-203``).  What is not: file / Zarr I/O, data generators,
augmentation, metrics, logging, BMZ / torchvision model sources -- the data arrives as arrays.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import numpy as np
import torch

from .. import _lib
from ..config.config import first
from ..data import norm as _norm
from ..models import build_model
from ..utils.misc import load_model_checkpoint, to_pytorch_format
from .inference import apply_head_activations, predict_by_chunks, predict_volume
from .train import Trainer


class Base_Workflow:
    """Common machinery of the semantic-segmentation and denoising workflows on the B200 engine."""

    loss_kind = "bce"

    def __init__(self, cfg, job_identifier: str, device, system_dict: Optional[Dict[str, int]] = None, args: Any = None):
        self.cfg = cfg
        self.job_identifier = job_identifier
        self.device = torch.device(device)
        self.test_device = self.device
        self.system_dict = system_dict or {}
        self.args = args
        self.ndim = 3 if cfg.PROBLEM.NDIM == "3D" else 2
        self.dims = self.ndim
        # (N, [Z,] Y, X, C) <-> (N, C, [Z,] Y, X): permutations only, the tensor keeps channels-last strides (misc.py:689-733)
        self.axes_order = (0, 3, 1, 2) if self.ndim == 2 else (0, 4, 1, 2, 3)
        self.axes_order_back = (0, 2, 3, 1) if self.ndim == 2 else (0, 2, 3, 4, 1)
        self.model = None
        self.trainer: Optional[Trainer] = None
        self.model_output_channels: List[int] = []
        self.model_output_channel_info: List[str] = []
        self.head_activations: List[str] = []
        self.separated_class_channel = False
        n = cfg.DATA.NORMALIZATION
        self.norm_module = {                                                              # base_workflow.py:374-387
            "type": n.TYPE, "target_type": "mask", "out_dtype": "float32", "norm_target": False,
            "percentile_clip": n.PERC_CLIP.ENABLE, "per_lower_bound": n.PERC_CLIP.LOWER_PERC, "per_upper_bound": n.PERC_CLIP.UPPER_PERC,
            "lower_bound_val": list(n.PERC_CLIP.LOWER_VALUE), "upper_bound_val": list(n.PERC_CLIP.UPPER_VALUE),
            "mean": list(n.ZERO_MEAN_UNIT_VAR.MEAN_VAL), "std": list(n.ZERO_MEAN_UNIT_VAR.STD_VAL),
        }
        self.test_norm_module = dict(self.norm_module, train_normalization=False, out_dtype="float32")
        self.current_sample: Dict[str, Any] = {}
        self.define_activations_and_channels()

    # ---------------------------------------------------------------------------------------------- hooks
    def define_activations_and_channels(self):
        """Subclasses set ``model_output_channels``, ``model_output_channel_info``, ``head_activations``,
        ``separated_class_channel`` (reference ``base_workflow.py:477-520``) and then call this check."""
        if not self.model_output_channels or not self.head_activations:
            raise ValueError("'model_output_channels' and 'head_activations' need to be defined. Correct define_activations_and_channels() function")
        if len(self.head_activations) != sum(self.model_output_channels):
            raise ValueError("'head_activations' must hold one value per output channel")

    def after_merge_patches(self, pred):
        """Called with the merged prediction of one sample; returns what the workflow keeps of it."""
        return pred

    def after_one_chunk_workflow_process(self, chunks, patch_in_data=None, added_pad=None):
        """By-chunks twin of ``after_merge_patches`` (reference ``base_workflow.py`` hook of the same name)."""
        return [self.after_merge_patches(c) for c in chunks]

    def prepare_targets(self, targets, batch=None):
        return targets

    # ---------------------------------------------------------------------------------------------- model
    def prepare_model(self):
        """``build_model`` + optional checkpoint (reference ``:906-995``)."""
        (self.model, self.model_build_kwargs_name, _, _, _, self.model_build_kwargs, self.network_stride) = build_model(
            self.cfg, self.model_output_channels, self.model_output_channel_info, self.head_activations, self.device)
        if self.cfg.MODEL.LOAD_CHECKPOINT:
            self.start_epoch, self.checkpoint_path = load_model_checkpoint(self.cfg, self.job_identifier, self.model, self.device)
        return self.model

    def set_engine(self, dtype=torch.bfloat16):
        """Storage / tensor-core dtype of the engine (float32 = the exact path)."""
        assert self.model is not None, "call prepare_model() first"
        self.model.set_engine(dtype=dtype)
        return self

    def apply_model_activations(self, pred: torch.Tensor, training: bool = False) -> torch.Tensor:
        """``(N, C, ...)`` logits -> activated prediction.  In training ``ce_sigmoid`` / ``ce_softmax`` stay logits (the loss
        applies them, reference ``:1396-1457``)."""
        acts = [("linear" if (training and a.lower() in ("ce_sigmoid", "ce_softmax")) else a) for a in self.head_activations]
        if all(a.lower() == "linear" for a in acts):
            return pred
        cl = pred.permute(self.axes_order_back)
        if self.ndim == 2:
            cl = cl[:, None]
        out = torch.empty(cl.shape, dtype=torch.float32, device=cl.device)
        apply_head_activations(cl.contiguous(), acts, out)
        if self.ndim == 2:
            out = out[:, 0]
        return out.permute(self.axes_order)

    def model_call_func(self, in_img, is_train: bool = False, apply_act: bool = True):
        """``(N, [Z,] Y, X, C)`` array or tensor -> ``(N, C_out, [Z,] Y, X)`` prediction on the device (reference ``:832-902``)."""
        assert self.model is not None, "call prepare_model() first"
        x = to_pytorch_format(in_img, self.axes_order, self.device)
        pred = self.model(x)
        if apply_act:
            pred = self.apply_model_activations(pred, training=is_train)
        return pred

    # ------------------------------------------------------------------------------------------ inference
    def predict_batches_in_test(self, x_batch, y_batch=None, stats_name: str = "", disable_tqdm: bool = True):
        """Patches ``(n, [z,] y, x, C)`` -> predictions ``(n, [z,] y, x, C_out)`` in ``TRAIN.BATCH_SIZE`` batches, through the
        TTA ensemble when ``TEST.AUGMENTATION`` (reference ``:1632-1730``).  numpy in -> numpy out, CUDA in -> CUDA out."""
        is_np = isinstance(x_batch, np.ndarray)
        x = x_batch if not is_np else torch.from_numpy(np.ascontiguousarray(x_batch)).to(self.device)
        bs = int(self.cfg.TRAIN.BATCH_SIZE)
        outs = []
        with torch.no_grad():
            if self.cfg.TEST.AUGMENTATION:
                from ..data.post_processing.post_processing import ensemble_predictions
                for k in range(x.shape[0]):
                    p = ensemble_predictions(x[k], self.model_call_func, self.axes_order_back, self.axes_order, self.device, self.ndim,
                                             batch_size_value=bs, mode=self.cfg.TEST.AUGMENTATION_MODE,
                                             group=self.cfg.TEST.AUGMENTATION_GROUP)
                    outs.append(p.permute(self.axes_order_back))
            else:
                for k in range(int(math.ceil(x.shape[0] / bs))):
                    outs.append(self.model_call_func(x[k * bs:(k + 1) * bs]).permute(self.axes_order_back).float())
        pred = torch.cat(outs, 0)
        return pred.cpu().numpy() if is_np else pred

    def process_test_sample(self, X, norm: bool = True):
        """One test image ``([z,] y, x, C)`` (raw dtype): normalise -> crop with overlap / padding -> predict -> spline merge ->
        ``after_merge_patches`` (reference ``:1874-2013``).  Returns ``(prediction, what after_merge_patches returned)``."""
        assert self.model is not None, "call prepare_model() first"
        cfg = self.cfg
        self.model.eval()
        if norm:
            X, self.current_sample["norm_info"] = _norm.normalize_image(X, dict(self.test_norm_module))
        vol, patch = X, tuple(cfg.DATA.PATCH_SIZE)
        ov, pad = tuple(cfg.DATA.TEST.OVERLAP), tuple(cfg.DATA.TEST.PADDING)
        if self.ndim == 2 and cfg.TEST.get("FULL_IMG", False):
            return self._process_full_image(vol)
        if self.ndim == 2:
            # one (y, x, C) image: the 2D crop / merge mirrors around predict_batches_in_test (reference :1944-1997 with
            # crop_data_with_overlap / merge_data_with_overlap)
            from ..data.data_2D_manipulation import crop_data_with_overlap, merge_data_with_overlap
            img = vol[None]
            patches, _ = crop_data_with_overlap(img, patch, overlap=ov, padding=pad, verbose=False)
            pp = self.predict_batches_in_test(patches)
            pred = merge_data_with_overlap(pp, tuple(img.shape[:-1]) + (pp.shape[-1],), overlap=ov, padding=pad, verbose=False)[0]
            return pred, self.after_merge_patches(pred)
        if cfg.TEST.BY_CHUNKS.ENABLE:
            pred = predict_by_chunks(self.model, vol, patch, padding=pad, batch_size=int(cfg.TRAIN.BATCH_SIZE),
                                     head_activations=self.head_activations)
            # the by-chunks path post-processes chunk by chunk (semantic seg: fixed 0.5 threshold, semantic_seg.py:502-535);
            # voxel-wise hooks give the same result on the assembled volume
            return pred, self.after_one_chunk_workflow_process([pred])[0]
        else:
            pred = predict_volume(self.model, vol, patch, overlap=ov, padding=pad, batch_size=int(cfg.TRAIN.BATCH_SIZE),
                                  head_activations=self.head_activations, tta=bool(cfg.TEST.AUGMENTATION),
                                  tta_mode=cfg.TEST.AUGMENTATION_MODE, tta_group=cfg.TEST.AUGMENTATION_GROUP)
        return pred, self.after_merge_patches(pred)

    def _process_full_image(self, X):
        """``TEST.FULL_IMG`` (2D only, reference ``:2224-2290``): the whole ``(y, x, C)`` image goes through the model in one
        call -- zero-padded at the bottom / right to a multiple of ``2 ** levels`` (``check_downsample_division``), through the TTA
        ensemble when ``TEST.AUGMENTATION``, cropped back -- and then to ``after_full_image``."""
        from ..utils.util import check_downsample_division
        cfg = self.cfg
        is_np = isinstance(X, np.ndarray)
        Xp, o_shape = check_downsample_division(X[None], len(cfg.MODEL.FEATURE_MAPS) - 1)
        with torch.no_grad():
            if cfg.TEST.AUGMENTATION:
                from ..data.post_processing.post_processing import ensemble_predictions
                xin = torch.from_numpy(np.ascontiguousarray(Xp)).to(self.device) if is_np else Xp
                pred = ensemble_predictions(xin[0], self.model_call_func, self.axes_order_back, self.axes_order, self.device, self.ndim,
                                            batch_size_value=int(cfg.TRAIN.BATCH_SIZE), mode=cfg.TEST.AUGMENTATION_MODE,
                                            group=cfg.TEST.AUGMENTATION_GROUP)
            else:
                pred = self.model_call_func(Xp)
        pred = pred.permute(self.axes_order_back).float()[:, :o_shape[1], :o_shape[2]][0]
        pred = pred.cpu().numpy() if is_np else pred.contiguous()
        return pred, self.after_full_image(pred)

    def after_full_image(self, pred):
        """Called with the full-image prediction of one sample (reference hook ``after_full_image``); the two workflows of
        the hot path post-process it like a merged prediction."""
        return self.after_merge_patches(pred)

    # ------------------------------------------------------------------------------------------- training
    def prepare_trainer(self) -> Trainer:
        cfg = self.cfg
        assert self.model is not None, "call prepare_model() first"
        if getattr(self, "unsupported_loss_options", None):
            raise NotImplementedError(self.unsupported_loss_options)
        opt = str(first(cfg.TRAIN.OPTIMIZER)).lower()
        if opt not in ("adamw", "adam", "sgd"):
            raise NotImplementedError(f"TRAIN.OPTIMIZER={opt!r}: the fused optimiser kernels cover ADAMW, ADAM and SGD")
        sgd = dict(momentum=0.9, nesterov=True) if opt == "sgd" else {}      # timm's 'sgd' (see engine/__init__.py)
        self.trainer = Trainer(self.model, loss=self.loss_kind, optimizer=opt, lr=float(first(cfg.TRAIN.LR)),
                               betas=tuple(first(cfg.TRAIN.OPT_BETAS)), weight_decay=float(cfg.TRAIN.W_DECAY),
                               clip_norm=float(cfg.TRAIN.GRADIENT_CLIP_NORM), ignore_index=int(getattr(self, "ignore_index", -100)),
                               **sgd)
        return self.trainer

    def train(self, train_generator, val_generator=None, cuda_graph: bool = False):
        """The epoch loop of ``Base_Workflow.train`` (reference ``:1017-1290``) on in-memory generators: iterables of
        ``(batch, targets)`` in ``(N, [Z,] Y, X, C)`` layout with a ``len()`` (a list, a ``DataLoader`` ...).  Per epoch:
        ``train_one_epoch`` -> optional ``evaluate`` (+ ``reduceonplateau``) -> best-on-validation bookkeeping -> early stopping.
        File outputs of the reference (checkpoints, tensorboard, charts, log file) are left to the caller: ``self.optimizer[0]``
        is the Trainer (``state_dict()`` in ``torch.optim`` layout for ``save_model``).  Returns the per-epoch statistics."""
        from . import build_callbacks, prepare_optimizer
        from .train_engine import evaluate, train_one_epoch
        cfg = self.cfg
        if self.model is None:
            self.prepare_model()
        self.optimizer, self.lr_scheduler = prepare_optimizer(cfg, self.model, len(train_generator), loss=self.loss_kind)
        self.trainer = self.optimizer[0]
        self.early_stopping = build_callbacks(cfg)
        self.loss_names = ["loss"]
        if cuda_graph:
            b0, t0 = next(iter(train_generator))
            self.trainer.enable_cuda_graph(b0, self.prepare_targets(t0, b0))
        self.val_best_loss = float("inf")
        start = int(getattr(self, "start_epoch", 0) or 0)
        history = []
        for epoch in range(start, int(cfg.TRAIN.EPOCHS)):
            print("~~~ Epoch {}/{} ~~~\n".format(epoch + 1, cfg.TRAIN.EPOCHS))
            sampler = getattr(train_generator, "sampler", None)
            if sampler is not None and hasattr(sampler, "set_epoch"):
                sampler.set_epoch(epoch)
            train_stats, _ = train_one_epoch(cfg, model=self.model, model_call_func=self.model_call_func, loss_function=None,
                                             metric_function=None, prepare_targets=self.prepare_targets,
                                             data_loader=train_generator, optimizer=self.optimizer, device=self.device, epoch=epoch,
                                             lr_scheduler=self.lr_scheduler, verbose=bool(cfg.TRAIN.VERBOSE),
                                             loss_names=self.loss_names)
            log_stats = {**{f"train_{k}": v for k, v in train_stats.items()}, "epoch": epoch}
            if val_generator is not None:
                test_stats = evaluate(cfg, model=self.model, model_call_func=self.model_call_func, loss_function=None,
                                      metric_function=None, prepare_targets=self.prepare_targets, epoch=epoch,
                                      data_loader=val_generator, lr_scheduler=self.lr_scheduler, loss_names=self.loss_names,
                                      optimizer=self.optimizer)
                if test_stats["loss"] < self.val_best_loss:
                    print("Val loss improved from {} to {}".format(self.val_best_loss, test_stats["loss"]))
                    self.val_best_loss = test_stats["loss"]
                log_stats.update({f"test_{k}": v for k, v in test_stats.items()})
            history.append(log_stats)
            if val_generator is not None and self.early_stopping is not None:
                self.early_stopping(test_stats["loss"])
                if self.early_stopping.early_stop:
                    print("Early stopping")
                    break
        print("Finished Training")
        return history

    def train_step(self, batch, targets) -> torch.Tensor:
        """One iteration of ``train_one_epoch`` (``train_engine.py:106-203``) on a ``(N, [Z,] Y, X, C)`` batch; returns the loss
        as a device tensor (no host synchronisation)."""
        if self.trainer is None:
            self.prepare_trainer()
        self.model.train()
        return self.trainer.step(batch, self.prepare_targets(targets, batch))

"""Engine entry points of the hot path (``biapy/engine/__init__.py``): ``prepare_optimizer`` and ``build_callbacks`` with the
reference's signatures.  The "optimizer" returned is the :class:`~biapy_b200.engine.train.Trainer` of the model -- the object
that owns the flat parameter / gradient / moment buffers and launches the fused optimiser kernel -- which exposes the
``param_groups`` / ``state_dict`` / ``zero_grad`` surface the training loop and ``save_model`` use."""
from __future__ import annotations

from typing import List, Optional, Tuple


def _scalar(v, i: int = 0):
    """``TRAIN.LR`` / ``MIN_LR`` / ``OPTIMIZER`` are per-optimiser lists since BiaPy 3.6, scalars in older YAMLs."""
    return v[i] if isinstance(v, (list, tuple)) else v


_SCHEDULERS = ["reduceonplateau", "warmupcosine", "onecycle", "warmupreduceonplateau"]


def check_lr_scheduler(cfg) -> None:
    """The scheduler rules of ``check_configuration.py:3304-3351`` (same messages), applied before anything is built."""
    s = cfg.TRAIN.LR_SCHEDULER
    name = str(s.NAME)
    if name == "":
        return
    if name not in _SCHEDULERS:
        raise ValueError(f"'TRAIN.LR_SCHEDULER.NAME' must be in {_SCHEDULERS}")
    min_lr = list(s.MIN_LR) if isinstance(s.MIN_LR, (list, tuple)) else [s.MIN_LR]      # a bare float in older YAMLs (:3664-3666)
    if name in ("reduceonplateau", "warmupcosine") and all(x == -1.0 for x in min_lr):
        raise ValueError("'TRAIN.LR_SCHEDULER.MIN_LR' needs to be set when 'TRAIN.LR_SCHEDULER.NAME' is between "
                         "['reduceonplateau', 'warmupcosine']")
    if name == "reduceonplateau":
        if s.REDUCEONPLATEAU_PATIENCE == -1:
            raise ValueError("'TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_PATIENCE' needs to be set when 'TRAIN.LR_SCHEDULER.NAME' is "
                             "'reduceonplateau'")
        if cfg.TRAIN.PATIENCE != -1 and s.REDUCEONPLATEAU_PATIENCE >= cfg.TRAIN.PATIENCE:
            raise ValueError("'TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_PATIENCE' needs to be less than 'TRAIN.PATIENCE' ")
    if name == "warmupcosine":
        if s.WARMUP_COSINE_DECAY_EPOCHS == -1:
            raise ValueError("'TRAIN.LR_SCHEDULER.WARMUP_COSINE_DECAY_EPOCHS' needs to be set when 'TRAIN.LR_SCHEDULER.NAME' is "
                             "'warmupcosine'")
        if s.WARMUP_COSINE_DECAY_EPOCHS > cfg.TRAIN.EPOCHS:
            raise ValueError("'TRAIN.LR_SCHEDULER.WARMUP_COSINE_DECAY_EPOCHS' needs to be less than 'TRAIN.EPOCHS'")


def prepare_optimizer(cfg, model_without_ddp, steps_per_epoch: int, loss: Optional[str] = None) -> Tuple[List, List]:
    """Optimiser + LR scheduler per ``TRAIN.OPTIMIZER`` entry (reference ``biapy/engine/__init__.py:21-107``; one entry on the
    hot path).  ``warmupcosine`` starts from ``MIN_LR`` (``:58``); ``timm.create_optimizer_v2`` given a parameter *list* applies
    the weight decay to every parameter, which is what the fused kernel does.  `loss`: the workflow's loss kind
    (``bce`` / ``ce`` / ``n2v_mse``), default from the model's attached workflow or ``bce``."""
    from .schedulers import OneCycleLR, ReduceLROnPlateau, WarmUpCosineDecayScheduler, WarmUpReduceOnPlateauScheduler
    from .train import Trainer

    check_lr_scheduler(cfg)
    name = str(cfg.TRAIN.LR_SCHEDULER.NAME)
    opts = cfg.TRAIN.OPTIMIZER if isinstance(cfg.TRAIN.OPTIMIZER, (list, tuple)) else [cfg.TRAIN.OPTIMIZER]
    if len(opts) != 1:
        raise NotImplementedError("one optimiser per model on the B200 hot path (TRAIN.OPTIMIZER has several entries)")
    opt = str(opts[0]).lower()
    if opt not in ("adamw", "adam", "sgd"):
        raise NotImplementedError(f"TRAIN.OPTIMIZER={opts[0]!r}: the fused optimiser kernels cover ADAMW, ADAM and SGD")
    lr = float(_scalar(cfg.TRAIN.LR_SCHEDULER.MIN_LR if name == "warmupcosine" else cfg.TRAIN.LR))
    betas = cfg.TRAIN.OPT_BETAS
    betas = tuple(betas[0]) if isinstance(betas[0], (list, tuple)) else tuple(betas)
    # timm.optim.create_optimizer_v2(opt='SGD') is torch.optim.SGD(momentum=0.9, nesterov=True) (timm's `momentum` default and its
    # 'sgd' alias of 'nesterov'); 'ADAM' is torch.optim.Adam (L2 weight decay), 'ADAMW' the decoupled form
    sgd = dict(momentum=0.9, nesterov=True) if opt == "sgd" else {}
    trainer = Trainer(model_without_ddp, loss=loss or getattr(model_without_ddp, "loss_kind", "bce"), optimizer=opt, lr=lr,
                      betas=betas, weight_decay=float(cfg.TRAIN.W_DECAY), clip_norm=float(cfg.TRAIN.GRADIENT_CLIP_NORM), **sgd)
    sched = None
    if name == "reduceonplateau":
        sched = ReduceLROnPlateau(trainer, patience=int(cfg.TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_PATIENCE),
                                  factor=float(cfg.TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_FACTOR),
                                  min_lr=float(_scalar(cfg.TRAIN.LR_SCHEDULER.MIN_LR)))
    elif name == "warmupcosine":
        sched = WarmUpCosineDecayScheduler(lr=float(_scalar(cfg.TRAIN.LR)), min_lr=float(_scalar(cfg.TRAIN.LR_SCHEDULER.MIN_LR)),
                                           warmup_epochs=cfg.TRAIN.LR_SCHEDULER.WARMUP_COSINE_DECAY_EPOCHS,
                                           epochs=int(cfg.TRAIN.EPOCHS))
    elif name == "onecycle":
        sched = OneCycleLR(trainer, float(_scalar(cfg.TRAIN.LR)), epochs=int(cfg.TRAIN.EPOCHS), steps_per_epoch=steps_per_epoch)
    elif name == "warmupreduceonplateau":
        sched = WarmUpReduceOnPlateauScheduler(lr=float(_scalar(cfg.TRAIN.LR)), epochs=int(cfg.TRAIN.EPOCHS))
    elif name != "":
        raise ValueError(f"unknown TRAIN.LR_SCHEDULER.NAME {name!r}")
    return [trainer], [sched]


def build_callbacks(cfg):
    """``EarlyStopping(patience=TRAIN.PATIENCE)`` or None for ``-1`` (reference ``:110-132``)."""
    from ..utils.callbacks import EarlyStopping
    return EarlyStopping(patience=int(cfg.TRAIN.PATIENCE)) if int(cfg.TRAIN.PATIENCE) != -1 else None

"""The four learning-rate schedules ``prepare_optimizer`` can build (reference ``biapy/engine/__init__.py:74-104``), working on
anything that has ``param_groups``: BiaPy's two per-iteration warm-up rules (``biapy/engine/schedulers/``) and ``OneCycleLR`` /
``ReduceLROnPlateau`` with the arguments the reference passes and torch's defaults for the rest.

torch's own classes insist on a ``torch.optim.Optimizer`` instance; the engine's optimiser is the fused kernel driven by
:class:`~biapy_b200.engine.train.Trainer`, so the two schedules are restated here (the formulas are the published 1cycle /
plateau rules; ``tests/test_host_schedulers.py`` holds them to torch's sequences step by step)."""
from __future__ import annotations

import math
from typing import List, Sequence, Union

import numpy as np


def _write_lr(optimizer, lr: float) -> None:
    """`lr` into every parameter group, scaled by the group's optional ``lr_scale``."""
    for group in optimizer.param_groups:
        group["lr"] = lr * group.get("lr_scale", 1.0) if "lr_scale" in group else lr


class WarmUpCosineDecayScheduler:
    """Linear warm-up from 0 to `lr` over `warmup_epochs`, then half a cosine down to `min_lr` at `epochs` (reference
    ``schedulers/warmup_cosine_decay.py:13-79``).  Driven per iteration with a fractional epoch,
    ``step / len(data_loader) + epoch`` (``train_engine.py:113-116``)."""

    def __init__(self, lr: float, min_lr: float, warmup_epochs: int, epochs: int):
        self.lr, self.min_lr, self.warmup_epochs, self.epochs = lr, min_lr, warmup_epochs, epochs

    def lr_at(self, epoch: float) -> float:
        if epoch < self.warmup_epochs:
            return self.lr * epoch / self.warmup_epochs
        # same operation order as the reference (pi * elapsed / span): the doubles agree bit for bit
        angle = math.pi * (epoch - self.warmup_epochs) / (self.epochs - self.warmup_epochs)
        return self.min_lr + (self.lr - self.min_lr) * 0.5 * (1.0 + math.cos(angle))

    def adjust_learning_rate(self, optimizer, epoch: float | int) -> float:
        lr = self.lr_at(epoch)
        _write_lr(optimizer, lr)
        return lr


def plateau_table(lr: float, n_epochs: int) -> np.ndarray:
    """One rate per epoch: ten points of a linear ramp 0 -> lr, then lr; runs longer than 100 (300) epochs trade their last
    50 (100) plateau epochs for ten halvings lasting 5 (10) epochs each (reference
    ``schedulers/warmup_reduce_on_plateau.py:17-32``)."""
    table = list(np.linspace(0, lr, 10)) + [lr] * max(0, n_epochs - 10)
    halving_run = 10 if n_epochs > 300 else 5 if n_epochs > 100 else 0
    if halving_run:
        del table[len(table) - 10 * halving_run:]
        for _ in range(10):
            table += [table[-1] / 2] * halving_run
    return np.asarray(table, dtype=np.float64)


class WarmUpReduceOnPlateauScheduler:
    """The tabulated schedule of :func:`plateau_table`, looked up with the integer part of the (fractional) epoch and clamped
    to its last entry (reference ``schedulers/warmup_reduce_on_plateau.py:6-46``)."""

    def __init__(self, lr: float, epochs: int):
        self.lr, self.epochs = lr, epochs
        self.LR = plateau_table(lr, epochs)

    def adjust_learning_rate(self, optimizer, epoch: float | int) -> float:
        lr = float(self.LR[min(int(epoch), len(self.LR) - 1)])
        _write_lr(optimizer, lr)
        return lr


def _cos_anneal(start: float, end: float, pct: float) -> float:
    return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1.0)


class OneCycleLR:
    """Two-phase cosine 1cycle policy: ``max_lr/25 -> max_lr`` over the first 30 % of the steps, then down to
    ``max_lr/25e4``; Adam's beta1 (SGD's momentum) is cycled the opposite way between 0.95 and 0.85.  As in torch the
    constructor already applies step 0 and each ``step()`` moves one iteration on (``train_engine.py:172-173``)."""

    def __init__(self, optimizer, max_lr: Union[float, Sequence[float]], total_steps: int | None = None, epochs: int | None = None,
                 steps_per_epoch: int | None = None, pct_start: float = 0.3, cycle_momentum: bool = True, base_momentum: float = 0.85,
                 max_momentum: float = 0.95, div_factor: float = 25.0, final_div_factor: float = 1e4):
        if total_steps is None:
            if epochs is None or steps_per_epoch is None or epochs <= 0 or steps_per_epoch <= 0:
                raise ValueError("You must define either total_steps OR (epochs AND steps_per_epoch)")
            total_steps = epochs * steps_per_epoch
        if total_steps <= 0:
            raise ValueError(f"Expected positive integer total_steps, but got {total_steps}")
        if not 0 <= pct_start <= 1:
            raise ValueError(f"Expected float between 0 and 1 pct_start, but got {pct_start}")
        self.optimizer = optimizer
        self.total_steps = total_steps
        groups = optimizer.param_groups
        max_lrs = list(max_lr) if isinstance(max_lr, (list, tuple)) else [max_lr] * len(groups)
        if len(max_lrs) != len(groups):
            raise ValueError(f"Expected {len(groups)} values for max_lr, got {len(max_lrs)}")
        self._ends = (float(pct_start * total_steps) - 1.0, float(total_steps) - 1.0)
        self.cycle_momentum = cycle_momentum
        for g, m in zip(groups, max_lrs):
            g["max_lr"] = m
            g["initial_lr"] = m / div_factor
            g["min_lr"] = g["initial_lr"] / final_div_factor
            if cycle_momentum:
                if "betas" not in g and "momentum" not in g:
                    raise ValueError("optimizer must support momentum or beta1 with `cycle_momentum` option enabled")
                g["max_momentum"] = max_momentum
                g["base_momentum"] = base_momentum
        self.last_epoch = -1
        self._last_lr: List[float] = []
        self.step()

    def _values(self, g, step: int):
        first_end, last_end = self._ends
        if step <= first_end:
            pct = step / first_end
            lr = _cos_anneal(g["initial_lr"], g["max_lr"], pct)
            mom = _cos_anneal(g["max_momentum"], g["base_momentum"], pct) if self.cycle_momentum else None
        else:
            pct = (step - first_end) / (last_end - first_end)
            lr = _cos_anneal(g["max_lr"], g["min_lr"], pct)
            mom = _cos_anneal(g["base_momentum"], g["max_momentum"], pct) if self.cycle_momentum else None
        return lr, mom

    def step(self) -> None:
        self.last_epoch += 1
        if self.last_epoch > self.total_steps:
            raise ValueError(f"Tried to step {self.last_epoch} times. The specified number of total steps is {self.total_steps}")
        self._last_lr = []
        for g in self.optimizer.param_groups:
            lr, mom = self._values(g, self.last_epoch)
            g["lr"] = lr
            if mom is not None:
                if "betas" in g:
                    g["betas"] = (mom,) + tuple(g["betas"][1:])
                else:
                    g["momentum"] = mom
            self._last_lr.append(lr)

    def get_last_lr(self) -> List[float]:
        return list(self._last_lr)


class ReduceLROnPlateau:
    """``mode='min'``, relative threshold 1e-4, no cool-down (torch defaults): after more than `patience` epochs without a
    new best the rate is multiplied by `factor`, not below `min_lr`; updates smaller than `eps` are skipped."""

    def __init__(self, optimizer, patience: int = 10, factor: float = 0.1, min_lr: Union[float, Sequence[float]] = 0.0,
                 threshold: float = 1e-4, cooldown: int = 0, eps: float = 1e-8):
        if factor >= 1.0:
            raise ValueError("Factor should be < 1.0.")
        self.optimizer = optimizer
        groups = optimizer.param_groups
        self.min_lrs = list(min_lr) if isinstance(min_lr, (list, tuple)) else [min_lr] * len(groups)
        if len(self.min_lrs) != len(groups):
            raise ValueError(f"expected {len(groups)} min_lrs, got {len(self.min_lrs)}")
        self.patience, self.factor, self.threshold, self.cooldown, self.eps = patience, factor, threshold, cooldown, eps
        self.best = math.inf
        self.num_bad_epochs = 0
        self.cooldown_counter = 0
        self.last_epoch = 0
        self._last_lr = [g["lr"] for g in groups]

    def step(self, metrics, epoch=None) -> None:
        current = float(metrics)
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        if current < self.best * (1.0 - self.threshold):
            self.best = current
            self.num_bad_epochs = 0
        else:
            self.num_bad_epochs += 1
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.num_bad_epochs = 0
        if self.num_bad_epochs > self.patience:
            for g, floor in zip(self.optimizer.param_groups, self.min_lrs):
                old = float(g["lr"])
                new = max(old * self.factor, floor)
                if old - new > self.eps:
                    g["lr"] = new
            self.cooldown_counter = self.cooldown
            self.num_bad_epochs = 0
        self._last_lr = [g["lr"] for g in self.optimizer.param_groups]

    def get_last_lr(self) -> List[float]:
        return list(self._last_lr)

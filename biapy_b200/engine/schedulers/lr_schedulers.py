"""``OneCycleLR`` and ``ReduceLROnPlateau`` with the arguments ``prepare_optimizer`` passes (reference
``biapy/engine/__init__.py:74-96``) and torch's defaults for the rest, working on anything that has ``param_groups``.

torch's own classes insist on a ``torch.optim.Optimizer`` instance; the engine's optimiser is the fused kernel driven by
:class:`~biapy_b200.engine.train.Trainer`, so the two schedules are restated here (the formulas are the published 1cycle /
plateau rules; ``tests/test_host_schedulers.py`` holds them to torch's sequences step by step)."""
from __future__ import annotations

import math
from typing import List, Sequence, Union


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

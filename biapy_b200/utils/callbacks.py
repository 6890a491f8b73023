"""``EarlyStopping`` with the interface of ``biapy/utils/callbacks.py:20-100``: called once per epoch with the validation loss,
it raises the ``early_stop`` flag after `patience` consecutive epochs whose loss is worse than the best one by more than `delta`.
Attribute names (``counter``, ``best_score``, ``val_loss_min``, ``early_stop``) are the ones BiaPy code reads."""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


class EarlyStopping:
    def __init__(self, patience: int = 7, delta: float = 0, trace_func: Callable = print):
        self.patience, self.delta, self.trace_func = patience, delta, trace_func
        self.counter = 0
        self.early_stop = False
        self.best_score: Optional[float] = None        # minus the best loss: "larger is better", as in the reference
        self.val_loss_min = np.inf                      # set on the first improvement *after* the first epoch

    def __call__(self, val_loss: float) -> None:
        first_epoch = self.best_score is None
        worse = (not first_epoch) and (-val_loss < self.best_score + self.delta)      # an equal loss is not worse
        if worse:
            self.counter += 1
            self.trace_func(f"EarlyStopping counter: {self.counter} out of {self.patience}")
            self.early_stop = self.early_stop or self.counter >= self.patience
            return
        self.best_score = -val_loss
        if not first_epoch:
            self.val_loss_min, self.counter = val_loss, 0

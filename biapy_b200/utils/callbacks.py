"""``EarlyStopping`` with the interface of ``biapy/utils/callbacks.py:20-100``: called with the validation loss of every epoch,
raises ``early_stop`` after `patience` epochs in a row without an improvement of more than `delta`."""
from __future__ import annotations

from typing import Callable

import numpy as np


class EarlyStopping:
    def __init__(self, patience: int = 7, delta: float = 0, trace_func: Callable = print):
        self.patience = patience
        self.counter = 0
        self.best_score = None
        self.early_stop = False
        self.val_loss_min = np.inf
        self.delta = delta
        self.trace_func = trace_func

    def __call__(self, val_loss: float):
        score = -val_loss
        if self.best_score is None:                       # first epoch: only the reference point is taken
            self.best_score = score
        elif score < self.best_score + self.delta:
            self.counter += 1
            self.trace_func(f"EarlyStopping counter: {self.counter} out of {self.patience}")
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            self.best_score = score
            self.val_loss_min = val_loss
            self.counter = 0

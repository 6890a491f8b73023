"""Tabulated schedule: 10-epoch linear warm-up, a plateau, then halvings towards the end of long runs
(reference ``schedulers/warmup_reduce_on_plateau.py:6-46``)."""
from __future__ import annotations

import numpy as np

from .warmup_cosine_decay import _set_lr


class WarmUpReduceOnPlateauScheduler:
    def __init__(self, lr: float, epochs: int):
        self.lr = lr
        self.epochs = epochs
        self.LR = self._build_schedule(lr, epochs)

    @staticmethod
    def _build_schedule(learning_rate: float, n_epochs: int) -> np.ndarray:
        table = np.concatenate([np.linspace(0, learning_rate, 10), np.full(max(0, n_epochs - 10), learning_rate, dtype=np.float64)])
        # long runs give up their last 100 (50) plateau epochs for ten halvings of 10 (5) epochs each
        if n_epochs > 300:
            cut, run = 100, 10
        elif n_epochs > 100:
            cut, run = 50, 5
        else:
            return table
        table = table[:-cut]
        for _ in range(10):
            table = np.concatenate([table, np.full(run, table[-1] / 2)])
        return table

    def adjust_learning_rate(self, optimizer, epoch: float | int) -> float:
        lr = float(self.LR[min(int(epoch), len(self.LR) - 1)])
        _set_lr(optimizer, lr)
        return lr

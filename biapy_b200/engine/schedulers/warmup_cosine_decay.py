"""Per-iteration linear warm-up followed by a half-cosine decay (reference ``schedulers/warmup_cosine_decay.py:13-79``)."""
from __future__ import annotations

import math


def _set_lr(optimizer, lr: float) -> None:
    """``lr`` (times the group's optional ``lr_scale``) into every parameter group."""
    for group in optimizer.param_groups:
        group["lr"] = lr * group["lr_scale"] if "lr_scale" in group else lr


class WarmUpCosineDecayScheduler:
    def __init__(self, lr: float, min_lr: float, warmup_epochs: int, epochs: int):
        self.lr = lr
        self.min_lr = min_lr
        self.warmup_epochs = warmup_epochs
        self.epochs = epochs

    def adjust_learning_rate(self, optimizer, epoch: float | int) -> float:
        """`epoch` is fractional: ``step / len(data_loader) + epoch`` (``train_engine.py:113-116``)."""
        if epoch < self.warmup_epochs:
            lr = self.lr * epoch / self.warmup_epochs
        else:
            # operation order of the reference (pi * elapsed / span), so the doubles agree bit for bit
            angle = math.pi * (epoch - self.warmup_epochs) / (self.epochs - self.warmup_epochs)
            lr = self.min_lr + (self.lr - self.min_lr) * 0.5 * (1.0 + math.cos(angle))
        _set_lr(optimizer, lr)
        return lr

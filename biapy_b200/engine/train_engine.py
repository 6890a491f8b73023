"""``train_one_epoch`` / ``evaluate`` with the reference's signatures (``biapy/engine/train_engine.py:25-207, 209-330``) on the
fused training step of :class:`~biapy_b200.engine.train.Trainer`.

What the reference does per iteration -- ``model_call_func`` -> ``loss_function`` -> ``backward`` -> ``clip_grad_norm_`` ->
``optimizer.step`` -> ``zero_grad`` -- is ONE call here (``optimizer[0].step(batch, targets)``: forward, loss, backward, gradient
all-reduce, clipping and the optimiser kernel, captured in a CUDA graph when enabled), so `model_call_func`, `loss_function`
and `metric_function` are accepted for signature compatibility and only `metric_function` is called if given.  Kept from the
reference: the per-iteration schedulers (``:111-116``), the patch-shape check (``:121-125``), the 1cycle step after every
update (``:172-173``), ``sys.exit(1)`` on a non-finite loss (``:158-162``), the ``{loss_name: average, lr_name: last lr}``
statistics and the returned step index.  Different on purpose: the loss is read from the device with a lag of one iteration
(an asynchronous copy + event) instead of ``.item()`` + ``synchronize()`` every step, so the host never stalls the GPU; a
non-finite loss therefore stops the run one iteration later."""
from __future__ import annotations

import math
import sys
from typing import Callable, List, Optional, Sequence

import torch

from .schedulers import OneCycleLR, ReduceLROnPlateau, WarmUpCosineDecayScheduler, WarmUpReduceOnPlateauScheduler


class _LaggedLoss:
    """Device loss -> host float without a stream synchronisation: copy into pinned memory, record an event, read it when the
    next iteration has been queued."""

    _POOL: List = []           # pinned one-element buffers, reused across epochs (cudaHostAlloc per step would cost ~0.1 ms)

    def __init__(self):
        self.pending = []      # [(pinned tensor, event | None)]
        self.values: List[float] = []

    def push(self, loss) -> None:
        if isinstance(loss, torch.Tensor) and loss.is_cuda:
            host, ev = self._POOL.pop() if self._POOL else (torch.empty(1, dtype=torch.float64, pin_memory=True), torch.cuda.Event())
            host.copy_(loss.detach().reshape(-1)[:1], non_blocking=True)
            ev.record(torch.cuda.current_stream(loss.device))
            self.pending.append((host, ev))
        else:
            self.pending.append((torch.as_tensor(loss, dtype=torch.float64).reshape(-1), None))

    def drain(self, keep: int) -> None:
        """Read every pending loss but the newest `keep`; exit like the reference on a non-finite value."""
        while len(self.pending) > keep:
            host, ev = self.pending.pop(0)
            if ev is not None:
                ev.synchronize()
            v = float(host[0])
            if ev is not None:
                self._POOL.append((host, ev))
            if not math.isfinite(v):
                print("Loss is {}, stopping training".format(v))
                sys.exit(1)
            self.values.append(v)


def _global_avg(values: List[float], device) -> float:
    """MetricLogger.synchronize_between_processes + global_avg (reference misc.py): sum of values / sum of counts over the ranks."""
    tot, cnt = float(sum(values)), float(len(values))
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        t = torch.tensor([tot, cnt], dtype=torch.float64,
                         device=device if (device is not None and torch.distributed.get_backend() == "nccl") else "cpu")
        torch.distributed.all_reduce(t)
        tot, cnt = float(t[0]), float(t[1])
    return tot / cnt if cnt else float("nan")


def _max_lr(opt) -> float:
    return max(float(g["lr"]) for g in opt.param_groups)


def train_one_epoch(cfg, model, model_call_func: Optional[Callable], loss_function: Optional[Callable],
                    metric_function: Optional[Callable], prepare_targets: Callable, data_loader, optimizer: Sequence, device,
                    epoch: int, log_writer=None, lr_scheduler: Optional[Sequence] = None, verbose: bool = False, memory_bank=None,
                    total_iters: int = 0, contrast_warmup_iters: int = 0, loss_names: Optional[List[str]] = None):
    """One pass over `data_loader` (an iterable of ``(batch, targets)`` in BiaPy's ``(N, [Z,] Y, X, C)`` layout with a
    ``len()``).  Returns ``(stats, last step index)`` like the reference."""
    if memory_bank is not None:
        raise NotImplementedError("contrastive training (memory bank) is outside the B200 hot path")
    loss_names = list(loss_names) if loss_names else ["loss"]
    lr_names = [n.replace("loss", "lr", 1) for n in loss_names]
    lr_scheduler = list(lr_scheduler) if lr_scheduler is not None else [None] * len(optimizer)
    if len(optimizer) != 1:
        raise NotImplementedError("one optimiser per model on the B200 hot path")
    trainer, sched = optimizer[0], lr_scheduler[0]
    model.train(True)
    trainer.zero_grad()
    name = str(cfg.TRAIN.LR_SCHEDULER.NAME)
    patch = tuple(cfg.DATA.PATCH_SIZE[:-1])
    n_batches = len(data_loader)
    lag = _LaggedLoss()
    step, last_lr = -1, _max_lr(trainer)
    for step, (batch, targets) in enumerate(data_loader):
        # per-iteration (not per-epoch) schedules: fractional epoch = step / len(data_loader) + epoch
        if name in ("warmupcosine", "warmupreduceonplateau") and isinstance(
                sched, (WarmUpCosineDecayScheduler, WarmUpReduceOnPlateauScheduler)):
            sched.adjust_learning_rate(trainer, step / n_batches + epoch)
        targets = prepare_targets(targets, batch)
        if tuple(batch.shape[1:-1]) != patch:
            raise ValueError("Trying to input data with different shape than 'DATA.PATCH_SIZE'. Check your configuration."
                             f" Input: {tuple(batch.shape[1:-1])} vs PATCH_SIZE: {patch}")
        last_lr = _max_lr(trainer)                      # the rate this update runs with
        loss = trainer.step(batch, targets)             # forward + loss + backward + all-reduce + clip + update
        if isinstance(sched, OneCycleLR) and name == "onecycle":
            sched.step()
        lag.push(loss)
        lag.drain(keep=1)
        if log_writer is not None and lag.values:
            log_writer.update(head="loss", **{loss_names[0]: lag.values[-1]})
            log_writer.update(head="opt", **{lr_names[0]: last_lr})
        if verbose and step % 10 == 0 and lag.values:
            print("Epoch: [{}]  [{}/{}]  {}: {:.4f}  {}: {:.6f}".format(epoch + 1, step, n_batches, loss_names[0], lag.values[-1],
                                                                       lr_names[0], last_lr))
    lag.drain(keep=0)
    if getattr(trainer, "loss_kind", None) == "ce":
        trainer.check_labels()                          # the host is synchronised here anyway
    avg = _global_avg(lag.values, device)
    stats = {loss_names[0]: avg, lr_names[0]: last_lr}
    print("[Train] averaged stats:", "  ".join(f"{k}: {v:.6f}" for k, v in stats.items()))
    return stats, step


@torch.no_grad()
def evaluate(cfg, model, model_call_func: Optional[Callable], loss_function: Optional[Callable], metric_function: Optional[Callable],
             prepare_targets: Callable, epoch: int, data_loader, lr_scheduler: Optional[Sequence] = None, memory_bank=None,
             loss_names: Optional[List[str]] = None, optimizer: Optional[Sequence] = None):
    """Validation pass: eval mode, forward + loss per batch (``Trainer.evaluate``), average, then the ``reduceonplateau`` step
    on that average (reference ``:209-330``).  `optimizer` (the list holding the Trainer) is an extra argument: the reference
    computes the loss with `loss_function` on torch tensors, here the loss kernel belongs to the Trainer."""
    if memory_bank is not None:
        raise NotImplementedError("contrastive training (memory bank) is outside the B200 hot path")
    if not optimizer:
        raise ValueError("evaluate() needs optimizer=[trainer]: the loss kernels are driven by the Trainer")
    trainer = optimizer[0]
    loss_names = list(loss_names) if loss_names else ["loss"]
    model.eval()
    lag = _LaggedLoss()
    for batch in data_loader:
        images, targets = batch[0], batch[1]
        targets = prepare_targets(targets, images)
        lag.push(trainer.evaluate(images, targets))
        lag.drain(keep=2)
    lag.drain(keep=0)
    # metric_logger.synchronize_between_processes() + global_avg (reference :318-321): every rank must feed the SAME validation
    # loss to ReduceLROnPlateau / EarlyStopping / the best-checkpoint test, or learning rates and stop epochs diverge
    avg = _global_avg(lag.values, getattr(trainer, "device", None))
    stats = {loss_names[0]: avg}
    print("[Val] averaged stats:", "  ".join(f"{k}: {v:.6f}" for k, v in stats.items()))
    if lr_scheduler and str(cfg.TRAIN.LR_SCHEDULER.NAME) == "reduceonplateau":
        for sched in lr_scheduler:
            if isinstance(sched, ReduceLROnPlateau):
                sched.step(avg, epoch=epoch)
    return stats

"""Training step of the U-Net path on the B200 engine: forward, loss, backward, gradient all-reduce, optimiser.

Mirrors the hot loop of ``biapy/engine/train_engine.py:106-203`` (``model_call_func(is_train=True)`` ->
``loss_function`` -> ``loss.backward()`` -> optional ``clip_grad_norm_`` -> ``optimizer.step()``) for the two
workflows of the hot path:

* semantic segmentation: ``BCEWithLogits`` for ``N_CLASSES <= 2`` / ``CrossEntropy`` otherwise
  (``biapy/engine/metrics.py:544-546, 577-586``);
* denoising: Noise2Void masked MSE (``metrics.py:2265-2286``).

B200-first choices: parameters, gradients and Adam moments live in three flat fp32 buffers (the ``nn.Parameter``
objects become views, so ``state_dict`` and checkpoints are unchanged); one fused optimiser kernel updates all
6.7 M parameters; data parallelism is ONE ``all_reduce`` over the flat gradient buffer on NCCL/NVLink instead of
DDP's bucketed hooks (``base_workflow.py:951-958``); nothing in the step synchronises the host (the reference
calls ``.item()`` twice and ``synchronize()`` once per step, ``train_engine.py:159,183,187``).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .. import _lib, ops
from .dist import allreduce_mean_
from .tape import TT, Tape


def _pow2_floor(n: int) -> int:
    return 1 << (max(1, int(n)).bit_length() - 1)


class FlatParams:
    """Re-home every parameter of `model` into one contiguous fp32 buffer (views keep the module API intact)."""

    def __init__(self, model: torch.nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        dev = params[0].device
        _lib.require_cuda(params[0], "model parameters")
        sizes = [p.numel() for p in params]
        # keep every parameter 16-byte aligned inside the flat buffer
        offs, total = [], 0
        for s in sizes:
            offs.append(total)
            total += (s + 3) // 4 * 4
        self.params = params
        self.offsets = offs
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            v = self.flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
        self.grad_views: Dict[torch.nn.Parameter, torch.Tensor] = {
            p: self.grad[o:o + p.numel()].view(p.shape) for p, o in zip(params, offs)}


def choose_allreduce_split(first: Dict[int, int], total: int, n_steps: int):
    """`first`: flat offset of a parameter gradient -> lowest backward-step index that touches it; `total`: elements of the flat
    buffer; `n_steps`: steps of the pass.  Returns (M, off) -- once the steps >= M have run, grad[off:] is final -- or None: `off` is
    the largest parameter boundary with at most 5 % of the elements in front of it, and at least a tenth of the steps must remain
    after the split to hide the collective."""
    cands = [o for o in sorted(first) if 0 < o <= total // 20]
    if not cands:
        return None
    off = cands[-1]
    m = min(i for o, i in first.items() if o >= off)
    if m <= 0 or m >= n_steps or m < n_steps // 10:
        return None
    return (m, off)


class Trainer:
    """One object = model + loss + optimiser state; ``step(x, target)`` runs one training iteration.

    x: ``(N, [Z,] Y, X, C)`` channels-last host (numpy / pinned torch) or device tensor -- BiaPy's batch layout.
    target: same layout; float mask for ``bce``, class indices in channel 0 for ``ce``, ``target||mask`` for ``n2v_mse``.
    """

    def __init__(self, model, loss: str = "bce", optimizer: str = "adamw", lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0, momentum: float = 0.0, clip_norm: float = 0.0,
                 process_group=None, nesterov: bool = False, loss_scale: Optional[float] = None, ignore_index: int = -100):
        """`optimizer`: 'adamw' | 'adam' | 'sgd' -- BiaPy's three ``TRAIN.OPTIMIZER`` values (``engine/__init__.py:58-70``).
        `loss_scale`: factor on the loss gradient that is divided out again inside the optimiser kernel; None = automatic
        (fp16 engine: the gradient of the *summed* loss is propagated so that it stays inside fp16's range, and a step whose
        gradient overflowed is skipped like ``torch.amp.GradScaler`` does; other dtypes: 1).  `ignore_index`: the
        ``CrossEntropyLoss(ignore_index=...)`` label of the reference's wrapper (``metrics.py:534-546``)."""
        self.model = model
        self.loss_kind = loss.lower()
        assert self.loss_kind in ("bce", "ce", "n2v_mse"), loss
        self.opt_kind = optimizer.lower()
        assert self.opt_kind in ("adamw", "adam", "sgd"), optimizer
        if nesterov and not momentum > 0:
            raise ValueError("Nesterov momentum requires a momentum")
        # one parameter group in torch.optim layout: the LR schedulers (engine/schedulers) and BiaPy's per-iteration
        # `adjust_learning_rate` write `param_groups[i]["lr"]` (and 1cycle the first beta); the optimiser launch reads it back
        self.param_groups = [{"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay, "momentum": momentum,
                              "nesterov": bool(nesterov)}]
        if self.opt_kind not in ("adamw", "adam"):
            del self.param_groups[0]["betas"]
            self._betas = tuple(betas)
        else:
            del self.param_groups[0]["nesterov"]
        self.clip_norm = clip_norm
        self.loss_scale = loss_scale
        self.ignore_index = int(ignore_index)
        self.fp = FlatParams(model)
        self.m = torch.zeros_like(self.fp.flat)
        self.v = torch.zeros_like(self.fp.flat) if self.opt_kind in ("adamw", "adam") else None
        self.t = 0
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.device = self.fp.flat.device
        self.ndim = model.ndim
        self._graph = None
        self._graph_b = None           # second half of the captured pass when the gradient all-reduce overlaps it (enable_cuda_graph)
        self._split = None             # (tape step index, flat offset): steps >= index finish every gradient at / after the offset
        self._on_split = None
        self._probe_log = None
        self._tail_reduced = False
        self.graph_launches = 0
        self._arena = ops.ZeroArena()
        # device-resident optimiser inputs (ops.optim_step_dev): hyper-parameters, step counters, scratch, gradient norm
        self._hp_dev = torch.zeros(ops.HP_SIZE, dtype=torch.float32, device=self.device)
        self._hp_last = None
        self._opt_state = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._derived = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._gsq = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._bad_labels = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._pack_plan = None         # (engine dtype, {tape key: pack job}) recorded by the first pass, replayed as one launch
        self._unscale = 1.0            # set by _loss_and_grad: what the optimiser multiplies the raw gradient with
        self._denom = None             # device divisor of the gradient (mask / counted-voxel count), or None
        self.sync_parameters()

    def sync_parameters(self):
        """What ``DistributedDataParallel`` does at construction (reference ``base_workflow.py:951-958``): rank 0's parameters
        and buffers (BatchNorm running statistics) replace every other rank's -- the reference seeds each process with
        SEED + rank (``misc.py:285``), so the replicas are NOT initialised identically.  Call again after loading a checkpoint
        on one rank only."""
        if self.world <= 1:
            return
        import torch.distributed as dist
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        dist.broadcast(self.fp.flat, src=src, group=self.pg)
        for b in self.model.buffers():
            if b.numel():
                dist.broadcast(b.data, src=src, group=self.pg)
        self.model.__dict__.pop("_pack_cache", None)

    # hyper-parameters live in `param_groups[0]` (see __init__); attribute access stays for callers and checkpoints
    lr = property(lambda self: self.param_groups[0]["lr"], lambda self, v: self.param_groups[0].__setitem__("lr", v))
    wd = property(lambda self: self.param_groups[0]["weight_decay"],
                  lambda self, v: self.param_groups[0].__setitem__("weight_decay", v))
    eps = property(lambda self: self.param_groups[0]["eps"], lambda self, v: self.param_groups[0].__setitem__("eps", v))
    momentum = property(lambda self: self.param_groups[0]["momentum"],
                        lambda self, v: self.param_groups[0].__setitem__("momentum", v))
    nesterov = property(lambda self: bool(self.param_groups[0].get("nesterov", False)))

    @property
    def betas(self):
        g = self.param_groups[0]
        return tuple(g["betas"]) if "betas" in g else self._betas

    @betas.setter
    def betas(self, v):
        if "betas" in self.param_groups[0]:
            self.param_groups[0]["betas"] = tuple(v)
        else:
            self._betas = tuple(v)

    def zero_grad(self, set_to_none: bool = False):
        """torch.optim API used by ``train_one_epoch``; the gradient buffer is cleared at the start of every pass anyway."""
        self.fp.grad.zero_()

    # ------------------------------------------------------------------------------------- optimiser state
    def state_dict(self) -> Dict:
        """Optimiser state in ``torch.optim.AdamW`` / ``SGD`` ``state_dict()`` layout (per-parameter ``step`` / ``exp_avg`` /
        ``exp_avg_sq`` or ``momentum_buffer``, one param group), so BiaPy checkpoints (``misc.py:328-386``) carry it and a
        torch optimiser can resume from it."""
        state = {}
        self.t = int(self._opt_state[0].item())          # the device counter is the truth (fp16 steps may have been skipped)
        for i, (p, o) in enumerate(zip(self.fp.params, self.fp.offsets)):
            sl = slice(o, o + p.numel())
            if self.opt_kind in ("adamw", "adam"):
                if self.t > 0:
                    state[i] = {"step": torch.tensor(float(self.t)), "exp_avg": self.m[sl].view(p.shape).detach().cpu().clone(),
                                "exp_avg_sq": self.v[sl].view(p.shape).detach().cpu().clone()}
            elif self.t > 0 and self.momentum:
                state[i] = {"momentum_buffer": self.m[sl].view(p.shape).detach().cpu().clone()}
        group = {"lr": self.lr, "weight_decay": self.wd, "params": list(range(len(self.fp.params)))}
        if self.opt_kind in ("adamw", "adam"):
            group.update(betas=tuple(self.betas), eps=self.eps, amsgrad=False)
        else:
            group.update(momentum=self.momentum, nesterov=self.nesterov)
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict, strict: bool = False):
        """Inverse of :meth:`state_dict`; also accepts the ``state_dict()`` of a torch AdamW / SGD over the same parameters."""
        st = sd.get("state", {})
        steps = []
        for i, (p, o) in enumerate(zip(self.fp.params, self.fp.offsets)):
            e = st.get(i, st.get(str(i)))
            if e is None:
                continue
            sl = slice(o, o + p.numel())
            if self.opt_kind in ("adamw", "adam"):
                self.m[sl].view(p.shape).copy_(e["exp_avg"])
                self.v[sl].view(p.shape).copy_(e["exp_avg_sq"])
                steps.append(int(float(e["step"])))
            elif "momentum_buffer" in e and e["momentum_buffer"] is not None:
                self.m[sl].view(p.shape).copy_(e["momentum_buffer"])
                steps.append(1)
        if steps:
            self.t = max(steps)
            self._opt_state[0] = self.t
        groups = sd.get("param_groups") or [{}]
        g = groups[0]
        self.lr = g.get("lr", self.lr)
        self.wd = g.get("weight_decay", self.wd)
        if "betas" in g:
            self.betas = tuple(g["betas"])
        self.eps = g.get("eps", self.eps)
        self.momentum = g.get("momentum", self.momentum)
        if "nesterov" in g and "nesterov" in self.param_groups[0]:
            self.param_groups[0]["nesterov"] = bool(g["nesterov"])

    # ------------------------------------------------------------------------------------------------- data
    def _to_device_cl(self, a, dtype=None) -> torch.Tensor:
        """host/device array in BiaPy layout -> (N, D, H, W, C) CUDA tensor (async copy from pinned memory)."""
        if not isinstance(a, torch.Tensor):
            a = torch.from_numpy(a)
        if not a.is_cuda:
            a = a.to(self.device, non_blocking=True)
        if self.ndim == 2:
            a = a.unsqueeze(1)
        return a

    # ------------------------------------------------------------------------------------------------- step
    def step(self, x, target) -> torch.Tensor:
        """Returns the (mean) loss as a 1-element float64 CUDA tensor; no host synchronisation."""
        if self._graph is not None:
            return self._step_graphed(x, target)
        xd = self._to_device_cl(x)
        td = self._to_device_cl(target)
        loss = self._forward_backward(xd, td)
        self._reduce_and_update()
        return loss

    # ---------------------------------------------------------------------------------------- CUDA graph mode
    def enable_cuda_graph(self, x_example, t_example):
        """Capture forward + loss + backward (~900 kernel launches of fixed shape) into one CUDA graph; the gradient
        all-reduce and the optimiser kernel stay eager (NCCL call, step-dependent bias correction).  Inputs are staged
        through static device buffers; H2D copies stay outside the graph on the same stream."""
        xs = self._to_device_cl(x_example)
        ts = self._to_device_cl(t_example)
        self._x_static = torch.empty_like(xs)
        self._t_static = torch.empty_like(ts)
        self._x_static.copy_(xs)
        self._t_static.copy_(ts)
        prof, ops.PROFILE = ops.PROFILE, None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                                   # warm-up: allocator, cudaFuncSetAttribute, caches
                self._forward_backward(self._x_static, self._t_static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = ops.LAUNCHES
        split = self._plan_allreduce_overlap() if self.world > 1 else None
        if split is not None:
            # Data-parallel run: the pass is captured as TWO graphs.  The first ends once backward has finished every gradient of
            # the flat buffer's tail (everything but the first encoder levels, > 95 % of the bytes); its all-reduce then runs on a
            # side stream under the second graph -- the expensive full-resolution encoder levels -- instead of after the pass.
            import gc
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            self._split = split
            cap = torch.cuda.Stream()
            gc.collect()
            torch.cuda.synchronize()
            cap.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(cap):
                ga.capture_begin()

                def on_split():
                    ops.flush_unpacks()                          # the weight gradients finished so far leave their packed form now
                    ga.capture_end()
                    gb.capture_begin(pool=ga.pool())

                self._on_split = on_split
                try:
                    self._loss_static = self._forward_backward(self._x_static, self._t_static)
                finally:
                    self._on_split = None
                gb.capture_end()
            torch.cuda.current_stream().wait_stream(cap)
            torch.cuda.synchronize()
            self._graph, self._graph_b = ga, gb
            self._comm_stream = torch.cuda.Stream()
            self._split_event = torch.cuda.Event()
        else:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._loss_static = self._forward_backward(self._x_static, self._t_static)
            self._graph = g
        self.graph_launches = ops.LAUNCHES - n0
        ops.PROFILE = prof
        return self

    def _plan_allreduce_overlap(self):
        """(M, off) or None.  One eager pass logs which backward step touches which parameter gradient; `off` is the largest
        parameter boundary of the flat buffer with at most 5 % of the elements in front of it, M the lowest step index that touches a
        gradient at or behind `off`: once the steps >= M have run, grad[off:] is final.  Only for the plain mean losses (no per-rank
        device divisor in front of the reduction) and when at least a tenth of the backward steps remain to hide the collective.
        `B200_OVERLAP_ALLREDUCE=0` keeps the single graph + all-reduce after the pass."""
        import os
        if os.environ.get("B200_OVERLAP_ALLREDUCE", "1") == "0" or self.loss_kind != "bce":
            return None
        self._probe_log = []
        try:
            self._forward_backward(self._x_static, self._t_static)
        finally:
            log, self._probe_log = self._probe_log, None
        torch.cuda.synchronize()
        n_steps = self._probe_nsteps
        if not log or not n_steps:
            return None
        base = self.fp.grad.data_ptr()
        first = {}
        for i, p in log:
            v = self.fp.grad_views.get(p)
            if v is None:
                return None                                      # a gradient outside the flat buffer: keep the simple form
            o = (v.data_ptr() - base) // 4
            first[o] = min(first.get(o, i), i)
        return choose_allreduce_split(first, self.fp.grad.numel(), n_steps)

    def _stage_host_batch(self, xs: torch.Tensor, ts: torch.Tensor):
        """Host batch -> device staging buffers on a dedicated copy stream (two slots), then a device-to-device copy
        into the graph's static inputs on the compute stream.  `step()` never blocks the host, so the host-to-device copy
        of step i+1 runs while step i computes; the compute stream only waits for the copy of its own batch."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [(torch.empty_like(self._x_static, dtype=xs.dtype), torch.empty_like(self._t_static, dtype=ts.dtype))
                           for _ in range(2)]
            self._stage_ready = [torch.cuda.Event() for _ in range(2)]
            self._stage_free = [torch.cuda.Event() for _ in range(2)]
            self._stage_k = 0
        k = self._stage_k
        self._stage_k ^= 1
        sx, st = self._stage[k]
        if sx.dtype != xs.dtype or st.dtype != ts.dtype or sx.shape != xs.shape or st.shape != ts.shape:
            sx = torch.empty(xs.shape, dtype=xs.dtype, device=self.device)
            st = torch.empty(ts.shape, dtype=ts.dtype, device=self.device)
            self._stage[k] = (sx, st)
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._stage_free[k])        # slot k was last read two steps ago
            sx.copy_(xs, non_blocking=True)
            st.copy_(ts, non_blocking=True)
            self._stage_ready[k].record(self._copy_stream)
        main.wait_event(self._stage_ready[k])
        self._x_static.copy_(sx, non_blocking=True)
        self._t_static.copy_(st, non_blocking=True)
        self._stage_free[k].record(main)

    def _step_graphed(self, x, target) -> torch.Tensor:
        xs = x if isinstance(x, torch.Tensor) else torch.from_numpy(x)
        ts = target if isinstance(target, torch.Tensor) else torch.from_numpy(target)
        if self.ndim == 2:
            xs, ts = xs.unsqueeze(1), ts.unsqueeze(1)
        if tuple(xs.shape) != tuple(self._x_static.shape) or tuple(ts.shape) != tuple(self._t_static.shape):
            # a trailing batch of another size (drop_last=False): the captured step has fixed shapes, run this one eagerly
            loss = self._forward_backward(xs.to(self.device, non_blocking=True), ts.to(self.device, non_blocking=True))
            self._reduce_and_update()
            return loss
        if not xs.is_cuda and not ts.is_cuda:
            self._stage_host_batch(xs, ts)
        else:
            if xs.data_ptr() != self._x_static.data_ptr():
                self._x_static.copy_(xs, non_blocking=True)
            if ts.data_ptr() != self._t_static.data_ptr():
                self._t_static.copy_(ts, non_blocking=True)
        self._graph.replay()
        if self._graph_b is not None:
            # grad[off:] is final: reduce it on the side stream while the second graph runs the rest of backward
            main = torch.cuda.current_stream()
            self._split_event.record(main)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(self._split_event)
                torch.distributed.all_reduce(self.fp.grad[self._split[1]:], group=self.pg)
            self._graph_b.replay()
            main.wait_stream(self._comm_stream)
            self._tail_reduced = True
        ops.LAUNCHES += self.graph_launches
        self._reduce_and_update()
        return self._loss_static

    def _forward_backward(self, xd: torch.Tensor, td: torch.Tensor) -> torch.Tensor:
        model = self.model
        tape = Tape(model.engine_dtype, self.device, training=True, conv_impl=model.conv_impl)
        tape.param_grads = dict(self.fp.grad_views)           # gradients land directly in the flat buffer
        if model.training:
            tape.rng_seed = model.next_rng_seed(self.device)
        ops.zero_(self.fp.grad)
        self._arena.begin(self.device)            # one fill for all the accumulators of this pass
        # Weight packs: the first pass launches them one by one and records the jobs whose source is a view of the flat
        # parameter buffer (stable address, contents rewritten by the optimiser); every later pass replays them as ONE batched
        # launch into the same packed tensors and hands those to the tape.  Weight-gradient un-packs are queued during backward
        # and leave as one launch, too (~95 of the ~310 launches of a config-[1] step were these 5-10 us permutation kernels).
        plan = self._pack_plan if (self._pack_plan is not None and self._pack_plan[0] == model.engine_dtype) else None
        if plan is not None and model.engine_dtype != torch.float32:
            ops.pack_batch([j for j in plan[1].values()], model.engine_dtype)
            tape._packed.update({k: j[2] for k, j in plan[1].items()})
        else:
            tape.pack_record = {}
        ops.UNPACK_QUEUE = [] if model.engine_dtype != torch.float32 else None
        try:
            loss = self._run_pass(model, tape, xd, td)
            ops.flush_unpacks()
            if tape.pack_record is not None:
                lo = self.fp.flat.data_ptr()
                hi = lo + self.fp.flat.numel() * 4
                keep = {k: j for k, j in tape.pack_record.items() if lo <= j[1].data_ptr() < hi and j[1].is_contiguous()}
                self._pack_plan = (model.engine_dtype, keep)
            return loss
        finally:
            ops.UNPACK_QUEUE = None
            self._arena.end()

    def _run_pass(self, model, tape, xd: torch.Tensor, td: torch.Tensor) -> torch.Tensor:
        if xd.dtype == model.engine_dtype and xd.is_contiguous():
            x_tt = TT(xd, requires_grad=False)
        else:
            if xd.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                xd = xd.float()
            x_tt = TT(torch.empty(xd.shape, dtype=model.engine_dtype, device=self.device), requires_grad=False)
            ops.convert(xd.contiguous(), x_tt.data)
        pred, cls = model._run(tape, x_tt)
        assert cls is None, "class heads are not part of the semantic-seg / denoising training step"
        loss = self._loss_and_grad(pred, td)
        pred.mark_written()
        if self._probe_log is not None:
            tape.pgrad_log = self._probe_log
        if self._on_split is not None and self._split is not None:
            tape.backward(split_at=self._split[0], on_split=self._on_split)
        else:
            tape.backward()
        self._probe_nsteps = tape.n_steps
        return loss

    def evaluate(self, x, target) -> torch.Tensor:
        """Mean loss of one batch, forward only (the validation loop, ``train_engine.py:266-317``): no tape closures, no
        gradient buffers, no update.  The caller puts the model in eval mode (``model.eval()``), as the reference does."""
        xd = self._to_device_cl(x)
        td = self._to_device_cl(target)
        model = self.model
        tape = Tape(model.engine_dtype, self.device, training=False, conv_impl=model.conv_impl)
        if model.training:
            tape.rng_seed = model.next_rng_seed(self.device)
        if xd.dtype == model.engine_dtype and xd.is_contiguous():
            x_tt = TT(xd, requires_grad=False)
        else:
            if xd.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                xd = xd.float()
            x_tt = TT(torch.empty(xd.shape, dtype=model.engine_dtype, device=self.device), requires_grad=False)
            ops.convert(xd.contiguous(), x_tt.data)
        pred, _ = model._run(tape, x_tt)
        numel = pred.data.numel()
        if self.loss_kind == "bce":
            t32 = td if td.dtype == torch.float32 and td.is_contiguous() else self._as_f32(td)
            return ops.bce_logits(pred.data, t32, None) / numel
        if self.loss_kind == "ce":
            cls = td[..., 0].long().contiguous()
            sums = ops.softmax_ce(pred.data, cls, None, ignore_index=self.ignore_index)
            self._bad_labels += sums[2:3]
            return sums[0:1] / sums[1:2]
        t32 = td if td.dtype == torch.float32 and td.is_contiguous() else self._as_f32(td)
        sums = ops.n2v_mse_sums(pred.data, t32)
        return sums[0:1] / sums[1:2]

    def check_labels(self):
        """Raise if a cross-entropy target held a label that is neither a class nor `ignore_index` (torch asserts on the device
        for those; the kernel skips and counts them).  Reads one scalar: call it where the host synchronises anyway."""
        bad = int(self._bad_labels.item())
        if bad:
            self._bad_labels.zero_()
            raise ValueError(f"cross-entropy target: {bad} voxels carry a label outside [0, n_classes) that is not "
                             f"ignore_index={self.ignore_index}")

    def _loss_and_grad(self, pred: TT, td: torch.Tensor) -> torch.Tensor:
        """Loss value (mean) and the gradient of the prediction.  The gradient is written as `S / P * dloss` with P a power of
        two near the loss's divisor and S the loss scale; `_reduce_and_update` hands `P / S` (and the exact divisor when it
        only exists on the device) to the optimiser kernel, so fp16 gradients stay in range and nothing is read on the host."""
        numel = pred.data.numel()
        fp16 = self.model.engine_dtype == torch.float16
        self._denom = None
        if self.loss_kind == "bce":
            S = float(self.loss_scale) if self.loss_scale else (float(_pow2_floor(numel)) if fp16 else 1.0)
            t32 = td if td.dtype == torch.float32 and td.is_contiguous() else self._as_f32(td)
            s = ops.bce_logits(pred.data, t32, pred.grad(), grad_scale=S / numel)
            self._unscale = 1.0 / S
            return s / numel
        S = float(self.loss_scale) if self.loss_scale else 1.0
        if self.loss_kind == "ce":
            cls = td[..., 0].long().contiguous()
            P = float(_pow2_floor(cls.numel()))
            sums = ops.softmax_ce(pred.data, cls, pred.grad(), grad_scale=S / P, ignore_index=self.ignore_index)
            self._bad_labels += sums[2:3]
        else:
            P = float(_pow2_floor(max(1, numel // 512)))          # Noise2Void masks ~0.2 % of the voxels (3d_denoising.yaml:11)
            t32 = td if td.dtype == torch.float32 and td.is_contiguous() else self._as_f32(td)
            sums = ops.n2v_mse_fused(pred.data, t32, pred.grad(), S / P)
        self._unscale = P / S
        self._denom = sums[1:2]
        return sums[0:1] / sums[1:2]

    def _as_f32(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype in (torch.float16, torch.bfloat16):
            out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
            ops.convert(t.contiguous(), out)
            return out
        return t.float().contiguous()

    def _reduce_and_update(self):
        """Gradient all-reduce + optimiser, no host synchronisation: the hyper-parameters go to the device as kernel arguments of
        a tiny write kernel (only when they changed), the clip factor, the fp16 overflow test, the bias corrections and the
        step counter are evaluated on the device (`b200_optim_step_dev`)."""
        g = self.fp.grad
        unscale, denom = self._unscale, self._denom
        if denom is not None and self.world > 1:
            # DDP averages the gradients of the per-rank *mean* losses: divide by this rank's own count before the reduction
            ops.scale_by_dev(g, denom, unscale)
            unscale, denom = 1.0, None
        if self._tail_reduced:
            # the tail went out under the second graph (_step_graphed); what is left are the first encoder levels' few parameters
            self._tail_reduced = False
            scale = allreduce_mean_(g[:self._split[1]], self.pg)
        else:
            scale = allreduce_mean_(g, self.pg)                  # one NCCL all-reduce over NVLink (no-op for 1 rank)
        clip = float(self.clip_norm) if self.clip_norm and self.clip_norm > 0 else 0.0
        need_gsq = clip > 0 or self.model.engine_dtype == torch.float16
        if need_gsq:
            ops.zero_(self._gsq)
            ops.sumsq(g, out=self._gsq)
        b1, b2 = self.betas
        hp = (float(self.lr), float(b1), float(b2), float(self.eps), float(self.wd), float(self.momentum),
              1.0 if self.nesterov else 0.0, float(scale * unscale), clip)
        if hp != self._hp_last:
            ops.write_floats(self._hp_dev, hp)
            self._hp_last = hp
        ops.optim_step_dev(self.opt_kind, self.fp.flat, g, self.m, self.v, self._hp_dev, self._opt_state, self._derived,
                           gsq=self._gsq if need_gsq else None, denom=denom)
        self.t += 1
        self.model.__dict__.pop("_pack_cache", None)            # eval-mode packed weights are stale now

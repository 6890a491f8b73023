"""Define-by-run executor for the U-Net family on the B200 kernels.

The model classes (``biapy_b200.models``) walk their reference-shaped module tree and call the methods of a
:class:`Tape`; each method launches the forward kernels immediately and, in training mode, records a closure
that launches the matching backward kernels.  :meth:`Tape.backward` replays the closures in reverse.

Design points (B200-first):

* activations are channels-last ``(N, D, H, W, C)`` in the engine dtype (bf16 by default, fp32 for the exact path);
* ``torch.cat([up, skip], 1)`` never happens: producers write straight into channel slices of a pre-allocated
  buffer (:meth:`Tape.new` + :meth:`TT.slice`) and gradients are read back through the same slices;
* the residual ``block(x) + shortcut(x)`` is an accumulate-in-place epilogue of the last convolution, so the
  incoming gradient of the sum is consumed by both producers without an extra pass;
* gradient buffers are allocated on first write; later writers accumulate in their own epilogue.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .. import _lib, ops


_DEBUG = os.environ.get("B200_TAPE_DEBUG", "0") != "0"


class TT:
    """A tensor on the tape: data + lazily allocated gradient; may be a channel slice of a parent buffer."""

    __slots__ = ("_data", "pending", "_meta", "_grad", "_ready", "parent", "off", "requires_grad", "padded", "sums", "dense_grad", "grad_sums")

    def __init__(self, data: Optional[torch.Tensor], requires_grad: bool = True, parent: "Optional[TT]" = None, off: int = 0):
        self._data = data
        self.pending = None         # lazy act(norm(src)): (materialise callback) until a consumer needs the tensor (Tape.norm_act)
        self._meta = None           # (shape, dtype, device) of a lazy tensor
        self._grad = None
        self._ready = False
        self.parent = parent
        self.off = off
        self.requires_grad = requires_grad
        self.padded = None          # zero-padded 16-channel copy (network inputs with < 16 channels, see Tape.conv)
        self.sums = None            # (N, C, 2) float64 channel sums left by the producing convolution's epilogue (Tape.conv)
        self.dense_grad = None      # slices only: the finished gradient as a dense tensor (left by Tape.maxpool's backward)
        # (C,) fp32 = sum over samples and voxels of this tensor's gradient, while every writer so far could state its share without
        # a pass over the data (normalisation backward: from its reductions; pointwise input gradients: W^T sums); None otherwise.
        # A convolution that produced the tensor takes its bias gradient from here (Tape._conv_backward, Tape.convT).
        self.grad_sums = None

    @property
    def data(self) -> torch.Tensor:
        """The tensor.  A lazy normalisation + activation result (Tape.norm_act(lazy=True)) is written out on first access --
        unless the consuming convolution applied it on its own operand path (Tape.conv, x-line kernel)."""
        if self._data is None and self.pending is not None:
            fn, self.pending = self.pending, None
            self._data = fn()
        return self._data

    @property
    def is_lazy(self) -> bool:
        return self._data is None and self.pending is not None

    @property
    def shape(self):
        return self._meta[0] if self._data is None else self._data.shape

    @property
    def c(self) -> int:
        return self.shape[-1]

    def slice(self, off: int, c: int) -> "TT":
        return TT(self.data[..., off:off + c], self.requires_grad, parent=self, off=off)

    # -- gradient plumbing -----------------------------------------------------------------------------
    @property
    def grad_ready(self) -> bool:
        return self.parent.grad_ready if self.parent is not None else self._ready

    def grad(self) -> torch.Tensor:
        """Gradient buffer (allocated on first use).  For slices: a view into the parent's gradient -- or, once the last writer
        has handed the finished sum on as a dense tensor (Tape.maxpool), that tensor."""
        if self.dense_grad is not None:
            return self.dense_grad
        if self.parent is not None:
            return self.parent.grad()[..., self.off:self.off + self.c]
        if self._grad is None:
            if self._data is None:
                shape, dtype, device = self._meta
                self._grad = torch.empty(shape, dtype=dtype, device=device)
            else:
                self._grad = torch.empty_like(self._data)
        return self._grad

    def mark_written(self):
        if self.parent is not None:
            # a slice can only be written first if the whole parent is zero-filled before; keep it simple: the
            # parent buffer is zero-initialised once, then every writer accumulates.
            self.parent.mark_written()
        else:
            self._ready = True

    def prepare_accumulate(self) -> bool:
        """Return the `accumulate` flag for a kernel about to write this gradient, and mark it written."""
        if self.parent is not None:
            root = self.parent
            while root.parent is not None:
                root = root.parent
            if not root._ready:
                ops.zero_(root.grad())
                root._ready = True
            root.grad_sums = None                   # a writer into a channel slice: the root's sums are no longer known
            return True
        acc = self._ready
        self._ready = True
        self.grad_sums = None                       # the writer restores it when it knows its share (see grad_sums)
        return acc


class Tape:
    def __init__(self, dtype: torch.dtype, device, training: bool, conv_impl: int = _lib.IMPL_AUTO,
                 pack_cache: Optional[Dict] = None):
        """`pack_cache`: a dict owned by the model that keeps packed weights across forward calls while the module is in
        eval mode (sliding-window inference runs 54 batches on frozen weights); entries are validated against the
        parameters' autograd version counters and the cache is dropped by `train()` / `load_state_dict()`."""
        self.dtype = dtype
        self._pack_cache = pack_cache
        self.rng_seed: Optional[torch.Tensor] = None        # int64[1] on the device; set by the model / trainer for dropout
        self._dropout_layers = 0
        self.device = device
        self.training = training
        self.impl = conv_impl
        self.steps: List[Callable[[], None]] = []
        self.use_xfold = os.environ.get("B200_XFOLD", "1") != "0"
        # conv -> norm: channel sums of the normalisation produced by the convolution epilogue (x-slab kernels)
        self.fuse_stats = os.environ.get("B200_FUSE_STATS", "1") != "0"
        self.pool_dense = os.environ.get("B200_POOL_DENSE", "1") != "0"
        self.dbias_share = os.environ.get("B200_DBIAS_SHARE", "1") != "0"
        # bias gradients from the normalisation backward's reductions instead of a pass over dy (TT.grad_sums)
        self.dbias_analytic = os.environ.get("B200_DBIAS_ANALYTIC", "1") != "0"
        # x-line kernel (csrc/conv_xline.cu): B200_XLINE = 0 off, 1 (default) where it measured faster than the x-folded kernels
        # (profiles/xline_probe_r2_*.log: the Cin = 48 launches, 0.39 vs 0.62 ms), 2 every launch it supports (Cin = 16 as well).
        # B200_XLINE_FUSE = 1: GroupNorm-apply + SiLU on the operand path of that convolution (the fused Conv3D + GN + SiLU launch).
        # Off by default: measured on the B200 the fused launch costs what the plain launch + the stand-alone HBM-speed apply pass
        # cost for 16 channels (0.31 vs 0.22 + 0.09 ms) and more for 48 (0.77 vs 0.40 + 0.28 ms) -- the activation arithmetic sits on
        # the issue-bound staging warps and every input line is staged 1.4-1.7 times (band / z-chunk halo); DESIGN 3.2d.
        self.xline = int(os.environ.get("B200_XLINE", "1"))
        self.xline_fuse = os.environ.get("B200_XLINE_FUSE", "0").lower()
        self.param_grads: Dict[torch.nn.Parameter, torch.Tensor] = {}
        # key -> (pack job, packed tensor) of every weight pack this pass launched on its own (Trainer: replayed as one launch)
        self.pack_record: Optional[Dict] = None
        self._packed: Dict[Tuple, torch.Tensor] = {}
        self._pad16: Dict[int, torch.Tensor] = {}
        # convolutions that wrote the same output tensor (block(x) + shortcut(x): the accumulate epilogue) have the same bias
        # gradient, the channel sums of that tensor's gradient: id(out) -> (out, bias gradient of the first one that ran)
        self._dbias_memo: Dict[int, Tuple[TT, torch.Tensor]] = {}
        self._dbias_touched: set = set()
        self.pgrad_log: Optional[list] = None        # [(step index, parameter)] of the backward pass when a list is put here
        self._cur_step: Optional[int] = None
        self.n_steps = 0

    # ------------------------------------------------------------------------------------------ helpers
    def new(self, like: torch.Tensor, channels: int, spatial: Optional[Sequence[int]] = None) -> TT:
        n = like.shape[0]
        sp = tuple(spatial) if spatial is not None else tuple(like.shape[1:4])
        return TT(torch.empty((n,) + sp + (channels,), dtype=self.dtype, device=self.device))

    @staticmethod
    def _grad_sums_of(t: TT) -> Optional[torch.Tensor]:
        """Channel sums of t's gradient if they are known (TT.grad_sums; for a channel slice: the slice of its parent's)."""
        if t.dense_grad is not None:
            return None
        if t.parent is None:
            return t.grad_sums
        if t.parent.parent is None and t.parent.grad_sums is not None:
            return t.parent.grad_sums[t.off:t.off + t.c]
        return None

    def _pgrad(self, p: torch.nn.Parameter) -> torch.Tensor:
        if self.pgrad_log is not None and self._cur_step is not None:
            self.pgrad_log.append((self._cur_step, p))          # which backward step touches which parameter gradient
        g = self.param_grads.get(p)
        if g is None:
            g = torch.zeros(p.shape, dtype=torch.float32, device=self.device)
            self.param_grads[p] = g
        return g

    def _pack(self, w, flip: bool, xfold: bool = False, wsrc=None) -> torch.Tensor:
        """Packed copy of a weight for this step; `w` is the cache key, `wsrc` the tensor to pack (defaults to w)."""
        src = w if wsrc is None else wsrc
        return self._cached((id(w), flip, xfold), src,
                            lambda: ops.pack_conv_weight_xfold(src, self.dtype, flip) if xfold
                            else ops.pack_conv_weight(src, self.dtype, flip))

    def _cached(self, key, src: torch.Tensor, build: Callable[[], torch.Tensor]) -> torch.Tensor:
        """Packed form of `src` for this step; served from the model's eval-mode cache when the parameter is unchanged."""
        t = self._packed.get(key)
        if t is None:
            ver = getattr(src, "_version", None)
            hit = self._pack_cache.get((key, self.dtype)) if self._pack_cache is not None else None
            if hit is not None and hit[0] == ver and hit[2] == src.data_ptr():
                t = hit[1]
            else:
                ops.LAST_PACK = None
                t = build()
                if self._pack_cache is not None:
                    self._pack_cache[(key, self.dtype)] = (ver, t, src.data_ptr())
                job = ops.LAST_PACK
                if self.pack_record is not None and job is not None and job[2] is t:
                    self.pack_record[key] = job
            self._packed[key] = t
        return t

    def _dense_for_xfold(self, t: torch.Tensor, other_c: int) -> torch.Tensor:
        """Gradients that arrive as a channel slice of a concat buffer are copied once into a dense tensor when the
        layer is small-channel: both the x-folded dgrad and wgrad need contiguous voxel rows, and the streaming copy is
        much cheaper than running the direct kernels on 32-byte TMA rows."""
        if (self.dtype == torch.float32 or not self.use_xfold or self.impl != _lib.IMPL_AUTO or t.stride(3) == t.shape[4]
                or t.shape[4] > 96 or other_c > 96 or t.shape[3] % 4 != 0 or t.shape[3] < 8 or t.shape[1] * t.shape[2] < 128):
            return t
        dense = torch.empty(t.shape, dtype=t.dtype, device=t.device)
        ops.binary(t, None, dense, ops.OP_COPY)
        return dense

    def _conv_launch(self, x: torch.Tensor, w, flip: bool, bias, y: torch.Tensor, k, accumulate: bool, wkey=None,
                     stats: bool = False) -> Optional[torch.Tensor]:
        """Pick the kernel family for (x -> y) and launch it with the matching weight packing.  `stats` asks for the channel
        sums of y from the epilogue; returns the (N, C, 2) float64 sums when the kernel produced them, else None (the
        normalisation then runs the stand-alone reduction)."""
        impl = self.impl
        wkey = w if wkey is None else wkey
        if self._xline_ok(x, y, k, accumulate, stats=stats and self.fuse_stats and not accumulate):
            return self._xline_launch(x, w, flip, bias, y, accumulate, wkey, stats)
        if impl == _lib.IMPL_AUTO and self.dtype != torch.float32 and self.use_xfold:
            x = self._dense_for_xfold(x, y.shape[4])
            if ops.conv_impl_query(x, y, k) == _lib.IMPL_XFOLD:
                impl = _lib.IMPL_XFOLD
            elif 64 < y.shape[4] <= 128 and y.shape[4] % 16 == 0 and x.shape[4] <= 96:
                # Cout in (64, 128]: two x-folded launches over output-channel slices (N = 256 and N = 4*(Cout-64)) beat the
                # direct kernel on 64-byte TMA rows (32->96 @64^3: 0.44 -> 0.25 ms); the slices are views of y, no copy
                cuts = [(0, 64), (64, y.shape[4])]
                if all(ops.conv_impl_query(x, y[..., a:b], k) == _lib.IMPL_XFOLD for a, b in cuts):
                    for a, b in cuts:
                        wp = self._cached((id(wkey), flip, "xfold", a, b), w,
                                          lambda a=a, b=b: ops.pack_conv_weight_xfold(
                                              (w[:, a:b] if flip else w[a:b]).detach().contiguous(), self.dtype, flip))
                        ops.conv_fprop(x, wp, None if bias is None else bias[a:b], y[..., a:b], k, accumulate=accumulate,
                                       impl=_lib.IMPL_XFOLD)
                    return
        wp = self._pack(wkey, flip, impl == _lib.IMPL_XFOLD, wsrc=w)
        if stats and impl == _lib.IMPL_XFOLD and self.fuse_stats and not accumulate and y.shape[4] == 16:
            sums = ops.zeros(y.shape[0] * y.shape[4] * 2, torch.float64, y.device)
            return sums if ops.conv_fprop_stats(x, wp, bias, y, k, sums, accumulate=accumulate) else None
        ops.conv_fprop(x, wp, bias, y, k, accumulate=accumulate, impl=impl)
        return None

    def _xline_ok(self, x: torch.Tensor, y: torch.Tensor, k, accumulate: bool, fused: bool = False, stats: bool = False) -> bool:
        if (self.xline <= 0 or self.impl != _lib.IMPL_AUTO or self.dtype == torch.float32 or tuple(k) != (3, 3, 3)
                or x.shape[3] != 128):
            return False
        cin, cout = x.shape[4], y.shape[4]
        if cout == 48:          # the input gradient of the 48 -> 16 layer: plain launches only
            if cin != 16 or accumulate or fused or stats:
                return False
        elif cout != 16 or cin not in (16, 48):
            return False
        elif self.xline == 1 and not fused and cin == 16:
            return False        # 16 -> 16: on a par with the x-slab kernel (0.22-0.28 vs 0.23-0.27 ms over five boxes); 48 -> 16: 1.6x faster
        return ops.conv_xline_supported(x, y, k)

    def _xline_launch(self, x, w, flip, bias, y, accumulate, wkey, stats, scale=None, shift=None, fuse=0, a_out=None):
        wp = self._cached((id(wkey), flip, "xline"), w, lambda: ops.pack_conv_weight_xline(w, self.dtype, flip))
        sums = None
        if stats and self.fuse_stats and not accumulate:
            sums = ops.zeros(y.shape[0] * y.shape[4] * 2, torch.float64, y.device)
        ops.conv_fprop_xline(x, wp, bias, y, accumulate=accumulate, scale=scale, shift=shift, fuse=fuse, a_out=a_out, sums=sums)
        return sums

    def _fuse_mode(self, x: torch.Tensor) -> int:
        """0 = keep normalisation-apply + SiLU as its own launch; 1 / 2 = apply it inside the x-line convolution with the exact /
        one-MUFU chain -- always the chain `ops.scale_shift_act` would run for this tensor, so both routes store the same bits."""
        if self.xline <= 0 or self.xline_fuse not in ("1", "on") or self.dtype == torch.float32:
            return 0
        return 2 if ops.norm_fast_ok(x) else 1

    @staticmethod
    def _k3(k) -> Tuple[int, int, int]:
        k = tuple(int(v) for v in k)
        return k if len(k) == 3 else (1,) + k

    def _f32(self, p):
        if p is None:
            return None
        t = p.detach()
        return t if t.dtype == torch.float32 else t.float()

    # --------------------------------------------------------------------------------------------- ops
    def conv(self, x: TT, mod: torch.nn.Module, out: Optional[TT] = None, accumulate: bool = False, stats: bool = False) -> TT:
        """y = conv(x) + bias, stride 1, 'same' padding.  `accumulate` adds into `out` (residual epilogue).  `stats`: the
        result goes straight into a normalisation -- leave its channel sums in `out.sums` when the kernel can."""
        w, b = mod.weight, mod.bias
        k = self._k3(w.shape[2:])
        cout, cin = w.shape[0], w.shape[1]
        assert x.c == cin, (x.c, cin)
        if x.is_lazy:
            fused = self._conv_fused(x, mod, out, accumulate, stats, k, cout, cin)
            if fused is not None:
                return fused
        if out is None:
            out = self.new(x.data, cout)
        # image-fed layers (Cin = 2, 4, 8): the x-folded slab kernels take the narrow input as it is (3x3x3), the pointwise
        # shortcut is an HBM stream on the CUDA cores; anything else with Cin < 16 is zero-padded to 16 channels
        narrow = (cin in (2, 4, 8) and self.dtype != torch.float32 and self.impl == _lib.IMPL_AUTO and self.use_xfold
                  and (k == (1, 1, 1) or ops.conv_impl_query(x.data, out.data, k) == _lib.IMPL_XFOLD))
        if (cin < 16 and cout % 16 == 0 and not x.requires_grad and self.dtype != torch.float32
                and self.impl != _lib.IMPL_SIMT and not narrow):
            return self._conv_padded_input(x, mod, out, accumulate, k, cout, cin)
        out.sums = self._conv_launch(x.data, w, False, self._f32(b), out.data, k, accumulate, stats=stats and not accumulate)
        if self.training:
            self._conv_backward(x, out, w, b, k, cout, cin, narrow)
        return out

    def _conv_backward(self, x: TT, out: TT, w, b, k, cout: int, cin: int, narrow: bool):
        def bwd(x=x, out=out, w=w, b=b, k=k, cout=cout, cin=cin, narrow=narrow):
            dy = out.grad() if (narrow and k == (1, 1, 1)) else self._dense_for_xfold(out.grad(), cin)
            assert out.grad_ready, "conv output gradient was never produced"
            if w.requires_grad:
                gb = self._pgrad(b) if (b is not None and b.requires_grad) else None
                first = w not in self.param_grads
                gs = self._grad_sums_of(out) if gb is not None else None
                memo = self._dbias_memo.get(id(out)) if gb is not None else None
                if _DEBUG:
                    print(f"[tape] conv bwd {cin}->{cout} k{k} @{tuple(x.shape[1:4])}: bias from "
                          f"{'grad_sums' if gs is not None else 'memo' if memo is not None else 'pass over dy' if gb is not None else 'none'}"
                          f" (out slice: {out.parent is not None}, dense_grad: {out.dense_grad is not None})")
                dysum = None                        # a tensor that holds sum(dy) per channel at this point of the stream
                if gs is not None:
                    # every writer of out.grad stated its channel sums (normalisation backward): that IS the bias gradient
                    ops.conv_wgrad(x.data, dy, cout, cin, k, self._pgrad(w), None, accumulate=not first, impl=self.impl)
                    ops.queue_float_add(gs, gb)
                    self._dbias_touched.add(id(b))
                    dysum = gs
                elif memo is not None and memo[1].numel() == gb.numel():
                    # second producer of `out`: its bias gradient is the first one's (copied with the batched un-packs)
                    ops.conv_wgrad(x.data, dy, cout, cin, k, self._pgrad(w), None, accumulate=not first, impl=self.impl)
                    ops.queue_float_add(memo[1], gb)
                    self._dbias_touched.add(id(b))
                    dysum = memo[1]
                else:
                    ops.conv_wgrad(x.data, dy, cout, cin, k, self._pgrad(w), gb, accumulate=not first, impl=self.impl)
                    if gb is not None and id(b) not in self._dbias_touched:
                        dysum = gb                  # holds nothing but this layer's sums (zero before this pass)
                        if self.dbias_share:
                            self._dbias_memo[id(out)] = (out, gb)
                    if gb is not None:
                        self._dbias_touched.add(id(b))
            else:
                dysum = None
            if x.requires_grad:
                # pointwise input gradient: dx = W^T dy voxel by voxel, so its channel sums are W^T sum(dy)
                want_sums = self.dbias_analytic and tuple(k) == (1, 1, 1) and dysum is not None and x.parent is None
                sums_prev = x.grad_sums
                acc = x.prepare_accumulate()
                self._conv_launch(dy, w, True, None, x.grad(), k, acc)
                if want_sums and (not acc or sums_prev is not None):
                    sums_x = sums_prev if acc else ops.zeros(cin, torch.float32, self.device)
                    ops.sums_through_pointwise(self._f32(w).contiguous(), dysum, sums_x)
                    x.grad_sums = sums_x
        self.steps.append(bwd)

    def _conv_fused(self, x: TT, mod, out: Optional[TT], accumulate: bool, stats: bool, k, cout: int, cin: int) -> Optional[TT]:
        """x is a lazy act(norm(src)) (Tape.norm_act): run the convolution on `src` with the normalisation-apply + SiLU on the
        operand path of the x-line kernel -- the fused Conv3D + GroupNorm + SiLU launch.  In training the kernel also writes the
        activated tensor (the weight gradient and nothing else reads it).  None when this convolution cannot take it."""
        src, st, fuse = x._meta[3]
        shape, dtype, device = x._meta[:3]
        if cout != 16 or tuple(k) != (3, 3, 3) or self.xline <= 0:
            return None
        if out is None:
            out = TT(torch.empty(tuple(shape[:4]) + (cout,), dtype=self.dtype, device=self.device))
        if not self._xline_ok(src.data, out.data, k, accumulate, fused=True):
            return None
        w, b = mod.weight, mod.bias
        a_buf = torch.empty(shape, dtype=dtype, device=device) if self.training else None
        out.sums = self._xline_launch(src.data, w, False, self._f32(b), out.data, accumulate, w, stats and not accumulate,
                                      scale=st.scale, shift=st.shift, fuse=fuse, a_out=a_buf)
        x.pending = None
        x._data = a_buf             # eval mode: the activated tensor never exists
        if self.training:
            self._conv_backward(x, out, w, b, k, cout, cin, narrow=False)
        return out

    def _conv_padded_input(self, x: TT, mod, out: TT, accumulate: bool, k, cout: int, cin: int) -> TT:
        """Image-fed convolutions (Cin = 1..3) run on the tensor cores by zero-padding the input and the weights to
        16 channels (K = taps*16 instead of a CUDA-core kernel); the padded gradient columns are discarded."""
        w, b = mod.weight, mod.bias
        if x.padded is None:
            buf = torch.zeros(tuple(x.shape[:4]) + (16,), dtype=self.dtype, device=self.device)
            ops.convert(x.data, buf[..., :cin])
            x.padded = buf
        w16 = self._pad16.get(id(w))
        if w16 is None:
            w16 = torch.zeros((cout, 16) + tuple(w.shape[2:]), dtype=torch.float32, device=self.device)
            w16[:, :cin] = w.detach()
            self._pad16[id(w)] = w16
        self._conv_launch(x.padded, w16, False, self._f32(b), out.data, k, accumulate)
        if self.training:
            def bwd(x=x, out=out, w=w, b=b, k=k, cout=cout, cin=cin):
                assert out.grad_ready, "conv output gradient was never produced"
                if w.requires_grad:
                    gb = self._pgrad(b) if (b is not None and b.requires_grad) else None
                    dw16 = torch.empty((cout, 16) + tuple(w.shape[2:]), dtype=torch.float32, device=self.device)
                    # the block output may live in a channel slice of a concat buffer: dense copy -> x-folded wgrad
                    dy = self._dense_for_xfold(out.grad(), 16)
                    ops.conv_wgrad(x.padded, dy, cout, 16, k, dw16, gb, accumulate=False, impl=self.impl, defer_unpack=False)
                    self._pgrad(w).add_(dw16[:, :cin])
            self.steps.append(bwd)
        return out

    def convT(self, x: TT, mod: torch.nn.Module, out: Optional[TT] = None) -> TT:
        w, b = mod.weight, mod.bias          # (Cin, Cout, *s)
        s = self._k3(w.shape[2:])
        cin, cout = w.shape[0], w.shape[1]
        assert x.c == cin
        if out is None:
            sp = (x.shape[1] * s[0], x.shape[2] * s[1], x.shape[3] * s[2])
            out = self.new(x.data, cout, sp)
        wf = self._f32(w).contiguous()
        tc = self.dtype != torch.float32 and self.impl != _lib.IMPL_SIMT and ops.convT_tc_supported(x.data, out.data, s)
        if tc:
            wp = self._cached((id(w), "convT"), w, lambda: ops.pack_convT_weight(wf, self.dtype, False))
            ops.convT_fprop_tc(x.data, wp, self._f32(b), out.data, s)
        else:
            ops.convT_fprop(x.data, wf, self._f32(b), out.data, s)
        if self.training:
            def bwd(x=x, out=out, w=w, b=b, s=s, wf=wf, tc=tc):
                dy = out.grad()
                assert out.grad_ready
                if w.requires_grad:
                    gb = self._pgrad(b) if (b is not None and b.requires_grad) else None
                    gs = self._grad_sums_of(out) if (gb is not None and tc) else None
                    if _DEBUG:
                        print(f"[tape] convT bwd {cin}->{cout} @{tuple(x.shape[1:4])}: bias from {'grad_sums' if gs is not None else 'pass over dy'}"
                              f" (out slice: {out.parent is not None}, parent sums: {out.parent is not None and out.parent.grad_sums is not None})")
                    if gs is not None:              # the concat gradient's channel sums are known: no pass over dy for the bias
                        ops.queue_float_add(gs, gb)
                        gb = None
                    if tc:
                        ops.convT_wgrad_tc(x.data, dy, self._pgrad(w), gb, s, accumulate=True)
                    else:
                        ops.convT_wgrad(x.data, dy, self._pgrad(w), gb, s)
                if x.requires_grad:
                    acc = x.prepare_accumulate()
                    if tc:
                        wpt = self._cached((id(w), "convT_d"), w, lambda: ops.pack_convT_weight(wf, self.dtype, True))
                        ops.convT_dgrad_tc(dy, wpt, x.grad(), s, accumulate=acc)
                    else:
                        ops.convT_dgrad(dy, wf, x.grad(), s, accumulate=acc)
            self.steps.append(bwd)
        return out

    def norm_act(self, x: TT, norm: Optional[torch.nn.Module], act: Optional[str], out: Optional[TT] = None,
                 lazy: bool = False) -> TT:
        """out = act(norm(x)).  norm: GroupNorm / InstanceNorm(affine) parameter holder or None.  `lazy`: the caller hands the
        result straight to a convolution -- the statistics are computed now, the apply pass is left to the consumer (see TT.data)."""
        act = (act or "none").lower()
        if norm is None and act in ("none", "linear"):
            return x
        sums, x.sums = x.sums, None
        # lazy route: GroupNorm / InstanceNorm + SiLU of a dense 16- or 48-channel tensor at W = 128 whose apply pass the x-line
        # convolution can run on its operand path (chain chosen by _fuse_mode: the one scale_shift_act would run)
        fuse = 0
        if (lazy and out is None and norm is not None and act == "silu" and x.parent is None and x.shape[3] == 128
                and x.c in (16, 48) and not isinstance(norm, torch.nn.modules.batchnorm._BatchNorm) and x.data.is_contiguous()):
            fuse = self._fuse_mode(x.data)
        if out is None and not fuse:
            out = self.new(x.data, x.c)
        if norm is None:
            ops.scale_shift_act(x.data, None, None, act, out.data)
            if self.training:
                def bwd(x=x, out=out, act=act):
                    assert out.grad_ready
                    if x.requires_grad:
                        acc = x.prepare_accumulate()
                        ops.act_bwd(x.data, out.grad(), act, x.grad(), accumulate=acc)
                self.steps.append(bwd)
            return out
        groups = norm_groups(norm, x.c)
        gamma, beta = getattr(norm, "weight", None), getattr(norm, "bias", None)
        g32, b32 = self._f32(gamma), self._f32(beta)
        if isinstance(norm, torch.nn.modules.batchnorm._BatchNorm):
            # BatchNorm / SyncBatchNorm (reference blocks.py:2117-2120): batch statistics + running-statistics update in
            # train mode, running statistics in eval mode.  Statistics over (N, D, H, W) = the per-channel sums of all samples.
            use_batch = norm.training or norm.running_mean is None
            if use_batch:
                sync = isinstance(norm, torch.nn.SyncBatchNorm) and norm.training
                pg = getattr(norm, "process_group", None)
                st = ops.norm_stats(x.data, groups, g32, b32, eps=float(norm.eps), batch_stats=True,
                                    sync_group=(pg if pg is not None else True) if sync else False, sums=sums)
                if norm.training and norm.track_running_stats and norm.running_mean is not None:
                    if norm.momentum is None:
                        raise NotImplementedError("BatchNorm(momentum=None) (cumulative average) is not implemented by the B200 "
                                                  "engine; BiaPy always passes a momentum")
                    count = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3] * st.world
                    ops.bn_update_running(st, x.c, count, float(norm.eps), float(norm.momentum), norm.running_mean,
                                          norm.running_var)
                    norm.num_batches_tracked.add_(1)
            else:
                if self.training and x.requires_grad:
                    raise NotImplementedError("gradients through eval-mode BatchNorm are not implemented by the B200 engine")
                st = ops.bn_eval_stats(x.data, norm.running_mean, norm.running_var, g32, b32, float(norm.eps))
        else:
            st = ops.norm_stats(x.data, groups, g32, b32, eps=float(norm.eps), sums=sums)
        if fuse:
            # the consumer decides: a 3x3x3 convolution into 16 channels applies scale / shift / SiLU on its own operand path
            # (Tape._conv_fused); anything else reads `.data`, which runs the stand-alone launch first
            out = TT(None)
            out._meta = (tuple(x.shape), x.data.dtype, x.data.device, (x, st, fuse))
            out.pending = lambda x=x, st=st, act=act, meta=out._meta: ops.scale_shift_act(
                x.data, st.scale, st.shift, act, torch.empty(meta[0], dtype=meta[1], device=meta[2]))
        else:
            ops.scale_shift_act(x.data, st.scale, st.shift, act, out.data)
        if self.training and st.mean is not None:
            def bwd(x=x, out=out, st=st, gamma=gamma, beta=beta, g32=g32, b32=b32, act=act):
                assert out.grad_ready
                dg = self._pgrad(gamma) if (gamma is not None and gamma.requires_grad) else None
                db = self._pgrad(beta) if (beta is not None and beta.requires_grad) else None
                dx, acc, gs = None, False, None
                if x.requires_grad:
                    gs_old = x.grad_sums
                    acc = x.prepare_accumulate()
                    dx = x.grad()
                    if (self.dbias_analytic and x.parent is None and getattr(st, "sums", None) is not None and (st.world or 1) == 1
                            and (not acc or gs_old is not None)):
                        gs = gs_old if acc else ops.zeros(x.c, torch.float32, self.device)
                if _DEBUG:
                    print(f"[tape] norm bwd c{x.c} @{tuple(x.shape[1:4])}: acc {acc}, sums {'yes' if gs is not None else 'no'} (slice: {x.parent is not None}, "
                          f"st.sums: {getattr(st, 'sums', None) is not None})")
                ops.norm_act_bwd(x.data, out.grad(), st, g32, b32, act, dx, dg, db, accumulate=acc, dy_dead=True, dx_sums=gs)
                if gs is not None:
                    x.grad_sums = gs
            self.steps.append(bwd)
        return out

    def dropout(self, x: TT, p: float, out: Optional[TT] = None) -> TT:
        """nn.Dropout(p) in training mode.  The mask is re-derived in backward from (seed, layer counter), never stored."""
        if out is None:
            out = self.new(x.data, x.c)
        if self.rng_seed is None:
            raise _lib.B200Error("dropout needs the tape's rng_seed (a 1-element int64 CUDA tensor)")
        layer = self._dropout_layers
        self._dropout_layers += 1
        ops.dropout(x.data, out.data, p, self.rng_seed, layer)
        if self.training:
            def bwd(x=x, out=out, p=p, layer=layer):
                assert out.grad_ready
                if x.requires_grad:
                    acc = x.prepare_accumulate()
                    ops.dropout(out.grad(), x.grad(), p, self.rng_seed, layer, accumulate=acc)
            self.steps.append(bwd)
        return out

    def upsample_linear(self, x: TT, scale: Sequence[int], out: Optional[TT] = None) -> TT:
        """nn.Upsample(mode='bilinear'|'trilinear', align_corners=False) by integer factors (reference blocks.py:605)."""
        s = self._k3(scale)
        sp = (x.shape[1] * s[0], x.shape[2] * s[1], x.shape[3] * s[2])
        if out is None:
            out = self.new(x.data, x.c, sp)
        ops.upsample_linear_fwd(x.data, out.data)
        if self.training:
            def bwd(x=x, out=out):
                assert out.grad_ready
                if x.requires_grad:
                    acc = x.prepare_accumulate()
                    ops.upsample_linear_bwd(out.grad(), x.grad(), accumulate=acc)
            self.steps.append(bwd)
        return out

    def maxpool(self, x: TT, window: Sequence[int]) -> TT:
        p = self._k3(window)
        sp = (x.shape[1] // p[0], x.shape[2] // p[1], x.shape[3] // p[2])
        out = self.new(x.data, x.c, sp)
        ops.maxpool_fwd(x.data, out.data, p)
        if self.training:
            def bwd(x=x, out=out, p=p):
                assert out.grad_ready
                if x.requires_grad:
                    acc = x.prepare_accumulate()
                    if x.parent is not None and acc and self.pool_dense and self.dtype != torch.float32:
                        # The pooled path is the last writer of a skip tensor's gradient, which lives in a channel slice of the
                        # concat buffer's gradient; the x-folded dgrad / wgrad of the producing block need it dense.  Write the
                        # sum (slice + routed gradient) straight into a dense tensor and hand that on, instead of updating the
                        # slice in place and copying it out afterwards (five strided copies per step, 0.28 ms).
                        dense = torch.empty(x.shape, dtype=x.data.dtype, device=x.data.device)
                        if ops.maxpool_bwd_to(x.data, out.grad(), x.grad(), dense, p):
                            x.dense_grad = dense
                            return
                    ops.maxpool_bwd(x.data, out.data, out.grad(), x.grad(), p, accumulate=acc)
            self.steps.append(bwd)
        return out

    def _route_grad(self, t: TT, d: torch.Tensor):
        """Add (or copy) a finished gradient tensor `d` into t's gradient."""
        if not t.requires_grad:
            return
        if t.prepare_accumulate():
            ops.binary(t.grad(), d, t.grad(), ops.OP_ADD)
        else:
            ops.binary(d, None, t.grad(), ops.OP_COPY)

    def add(self, a: TT, b: TT, out: Optional[TT] = None) -> TT:
        """out = a + b (out may be `b` itself: in-place residual add)."""
        if out is None:
            out = self.new(a.data, a.c)
        ops.binary(a.data, b.data, out.data, ops.OP_ADD)
        if self.training:
            def bwd(a=a, b=b, out=out):
                assert out.grad_ready
                for t in (a, b):
                    if t is not out:
                        self._route_grad(t, out.grad())
            self.steps.append(bwd)
        return out

    def copy_into(self, src: TT, dst: TT) -> TT:
        """dst = src (used when a skip tensor must be replicated into several concat buffers)."""
        ops.binary(src.data, None, dst.data, ops.OP_COPY)
        if self.training:
            def bwd(src=src, dst=dst):
                assert dst.grad_ready
                self._route_grad(src, dst.grad())
            self.steps.append(bwd)
        return dst

    def add_relu(self, a: TT, b: TT) -> TT:
        out = self.new(a.data, a.c)
        ops.binary(a.data, b.data, out.data, ops.OP_ADD_RELU)
        if self.training:
            def bwd(a=a, b=b, out=out):
                assert out.grad_ready
                d = torch.empty_like(out.data)
                ops.relu_mask_bwd(out.data, out.grad(), d)
                for t in (a, b):
                    self._route_grad(t, d)
            self.steps.append(bwd)
        return out

    def gate(self, psi: TT, x: TT, out: Optional[TT] = None) -> TT:
        """out = psi * x with a single-channel psi (attention gate, blocks.py:1116)."""
        if out is None:
            out = self.new(x.data, x.c)
        ops.binary(x.data, psi.data, out.data, ops.OP_MUL)
        if self.training:
            def bwd(psi=psi, x=x, out=out):
                assert out.grad_ready
                acc_x = x.prepare_accumulate()
                dpsi = torch.empty_like(psi.data)
                ops.gate_bwd(x.data, psi.data, out.grad(), dpsi, x.grad(), accumulate=acc_x)
                self._route_grad(psi, dpsi)
            self.steps.append(bwd)
        return out

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, split_at: Optional[int] = None, on_split=None):
        """Run the recorded steps in reverse.  `split_at` = M: `on_split()` is called once every step with index >= M has run (the
        Trainer ends one CUDA graph there and starts the next, see Trainer.enable_cuda_graph)."""
        n = len(self.steps)
        for i in range(n - 1, -1, -1):
            if on_split is not None and split_at is not None and i == split_at - 1:
                on_split()
            self._cur_step = i
            self.steps[i]()
        self._cur_step = None
        self.n_steps = n
        self.steps = []


def norm_groups(norm: torch.nn.Module, channels: int) -> int:
    if isinstance(norm, torch.nn.GroupNorm):
        return norm.num_groups
    if isinstance(norm, (torch.nn.InstanceNorm2d, torch.nn.InstanceNorm3d)):
        return channels
    if isinstance(norm, torch.nn.modules.batchnorm._BatchNorm):
        return channels                     # per-channel statistics; the batch axis is folded in by batch_stats=True
    raise NotImplementedError(f"normalization layer {type(norm).__name__} is not supported by the B200 engine "
                              "(supported: 'gn', 'in', 'none')")

"""Shared implementation of ``U_Net`` / ``ResUNet`` / ``Attention_U_Net`` for the B200 engine.

The three reference classes (``biapy/models/unet.py:29``, ``resunet.py:27``, ``attention_unet.py:34``) differ only
in the block types they instantiate; their constructor surface, attribute names, ``state_dict`` layout and
``forward`` return convention are reproduced here once.  ``forward`` runs the whole network through
:class:`biapy_b200.engine.tape.Tape` inside ONE ``torch.autograd.Function`` so that ``loss.backward()``,
``torch.optim`` and ``DistributedDataParallel`` keep working unchanged on top of hand-written kernels.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from .. import _lib, ops
from ..engine.tape import TT, Tape
from .blocks import (ConvBlock, ResConvBlock, ResUpBlock, UpBlock, get_decoder_feature_maps, init_weights)

_DTYPES = {"bf16": torch.bfloat16, "bfloat16": torch.bfloat16, "fp16": torch.float16, "float16": torch.float16,
           "fp32": torch.float32, "float32": torch.float32}


def default_engine_dtype() -> torch.dtype:
    return _DTYPES[os.environ.get("BIAPY_B200_DTYPE", "bf16").lower()]


class _NetFn(torch.autograd.Function):
    """Whole-network autograd node: forward = tape forward, backward = tape backward (all CUDA kernels ours)."""

    @staticmethod
    def forward(ctx, model, x, *params):
        outs, tape, out_tts, x_tt = model._execute(x, record=True, x_requires_grad=x.requires_grad)
        ctx.tape, ctx.out_tts, ctx.x_tt, ctx.params = tape, out_tts, x_tt, params
        ctx.ndim = model.ndim
        ctx.model_ref = model
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        tape: Tape = ctx.tape
        if tape is None:
            raise RuntimeError("biapy_b200: backward called twice on the same forward (activations were freed)")
        # fp16 engine: the gradient of a mean-reduced loss is ~1/numel per element, below fp16's normal range at patch sizes.  It is
        # propagated times the power of two that brings its largest element to ~1 and divided out of the (fp32) results -- what
        # torch.amp.GradScaler does for autocast models, chosen from the data because this autograd-compatibility path cannot know
        # the loss (one scalar read; the Trainer scales inside its loss kernel instead and never synchronises).
        scale = 1.0
        if tape.dtype == torch.float16:
            amax = max((float(g.abs().max()) for g in grads if g is not None), default=0.0)
            if amax > 0.0 and amax == amax and amax != float("inf"):
                import math
                scale = 2.0 ** (-math.ceil(math.log2(amax)))
        for tt, g in zip(ctx.out_tts, grads):
            if g is None:
                tt.grad().zero_()
            else:
                if ctx.ndim == 2:
                    g = g.unsqueeze(2)
                gcl = g.permute(0, 2, 3, 4, 1)
                if gcl.dtype != torch.float32 or not gcl.is_contiguous():
                    gcl = gcl.float().contiguous()
                if scale != 1.0:
                    gcl = gcl * scale
                ops.convert(gcl, tt.grad())
            tt.mark_written()
        arena = None
        if ops.ARENA is None and not torch.cuda.is_current_stream_capturing():
            arena = ctx.model_ref.__dict__.setdefault("_arena_bwd", ops.ZeroArena())
            arena.begin(ctx.x_tt.data.device)
        try:
            tape.backward()
        finally:
            if arena is not None:
                arena.end()
        gx = None
        if ctx.x_tt.requires_grad:
            gx32 = torch.empty(ctx.x_tt.shape, dtype=torch.float32, device=ctx.x_tt.data.device)
            ops.convert(ctx.x_tt.grad(), gx32)
            if scale != 1.0:
                gx32.mul_(1.0 / scale)
            gx = gx32.permute(0, 4, 1, 2, 3)
            if ctx.ndim == 2:
                gx = gx.squeeze(2)
        pg = tuple(tape.param_grads.get(p) if p.requires_grad else None for p in ctx.params)
        if scale != 1.0:
            torch._foreach_mul_([g for g in pg if g is not None], 1.0 / scale)
        ctx.tape = None
        return (None, gx) + pg


class UNetFamily(nn.Module):
    """Common constructor / forward of the three U-Net variants.  `variant` in {'unet','resunet','attention_unet'}."""

    variant = "unet"

    def __init__(self, image_shape=(256, 256, 1), activation="ELU", feature_maps=[32, 64, 128, 256],
                 drop_values=[0.1, 0.1, 0.1, 0.1], normalization="none", k_size=3, upsample_layer="convtranspose",
                 yx_down=[2, 2, 2, 2], z_down=[2, 2, 2, 2], output_channels=[1], separated_decoders=False,
                 divide_decoder_feature_maps=False, output_channel_info=["F"], explicit_activations: bool = False,
                 head_activations: List[str] = ["ce_sigmoid"], upsampling_factor=(), upsampling_position="pre",
                 isotropy=False, larger_io=True, conv_layers: List[int] = [2, 2, 2, 2, 2], contrast: bool = False,
                 contrast_proj_dim: int = 256, return_one_tensor: bool = False, conv_block_order: str = "conv_norm_act"):
        super().__init__()
        if len(output_channels) == 0:
            raise ValueError("'output_channels' needs to has at least one value")
        if contrast and len(output_channels) > 2:
            raise ValueError("If 'contrast' is True, 'output_channels' can only have two values at max: one for the main output and one for the class.")
        if contrast:
            raise NotImplementedError("contrastive heads are outside the B200 hot path")
        if len(upsampling_factor) > 0:
            raise NotImplementedError("super-resolution pre/post up-sampling is outside the B200 hot path")
        print("Selected output channels:")
        for i, info in enumerate(output_channel_info):
            print(f"  - {i} channel for {info} output")

        self.depth = len(feature_maps) - 1
        self.ndim = 3 if len(image_shape) == 4 else 2
        self.z_down = z_down
        self.yx_down = yx_down
        self.output_channels = output_channels
        self.output_channel_info = output_channel_info
        self.return_class = True if "class" in output_channel_info else False
        self.contrast = contrast
        self.explicit_activations = explicit_activations
        self.return_one_tensor = return_one_tensor
        if self.explicit_activations:
            assert len(head_activations) == sum(output_channels), \
                "If 'explicit_activations' is True, 'head_activations' needs to have the same number of values as 'output_channels'"
            self._head_act_names = [a.lower() for a in head_activations]
        activation = activation.lower() if isinstance(activation, str) else activation
        if type(isotropy) == bool:
            isotropy = [isotropy] * len(feature_maps)
        nd = self.ndim
        conv = nn.Conv3d if nd == 3 else nn.Conv2d
        convtranspose = nn.ConvTranspose3d if nd == 3 else nn.ConvTranspose2d
        pooling = nn.MaxPool3d if nd == 3 else nn.MaxPool2d
        res = self.variant == "resunet"

        def ksize(level, extra=0):
            k = k_size + extra
            if nd == 2:
                return (k, k)
            return (k, k, k) if isotropy[level] else (1, k, k)

        def pool_of(level):
            return (z_down[level], yx_down[level], yx_down[level]) if nd == 3 else (yx_down[level], yx_down[level])

        self.pre_upsampling = None
        # ENCODER
        self.down_path = nn.ModuleList()
        self.mpooling_layers = nn.ModuleList()
        in_channels = image_shape[-1]
        if larger_io:
            self.conv_in = ConvBlock(conv=conv, in_size=in_channels, out_size=feature_maps[0], k_size=ksize(0, 2),
                                     act=activation, norm=normalization, order=conv_block_order)
            in_channels = feature_maps[0]
        else:
            self.conv_in = None
        for i in range(self.depth):
            if res:
                blk = ResConvBlock(conv=conv, in_size=in_channels, out_size=feature_maps[i], k_size=ksize(i), act=activation,
                                   norm=normalization, dropout=drop_values[i], first_block=(i == 0), nconvs=conv_layers[i],
                                   order=conv_block_order)
            else:
                blk = ConvBlock(conv=conv, in_size=in_channels, out_size=feature_maps[i], k_size=ksize(i), act=activation,
                                norm=normalization, dropout=drop_values[i], nconvs=conv_layers[i], order=conv_block_order)
            self.down_path.append(blk)
            self.mpooling_layers.append(pooling(pool_of(i)))
            in_channels = feature_maps[i]
        if res:
            self.bottleneck = ResConvBlock(conv=conv, in_size=in_channels, out_size=feature_maps[-1], k_size=ksize(-1),
                                           act=activation, norm=normalization, dropout=drop_values[-1],
                                           nconvs=conv_layers[-1], order=conv_block_order)
        else:
            self.bottleneck = ConvBlock(conv=conv, in_size=in_channels, out_size=feature_maps[-1], k_size=ksize(-1),
                                        act=activation, norm=normalization, dropout=drop_values[-1], nconvs=conv_layers[-1],
                                        order=conv_block_order)
        # DECODER
        self.num_decoders = 1 if not separated_decoders else len(output_channels)
        dec_fm = get_decoder_feature_maps(feature_maps, self.num_decoders, divide_decoder_feature_maps)
        self.up_paths = nn.ModuleList([nn.ModuleList() for _ in range(self.num_decoders)])
        for j in range(self.num_decoders):
            in_channels = feature_maps[-1]
            for i in range(self.depth - 1, -1, -1):
                if res:
                    up = ResUpBlock(ndim=nd, convtranspose=convtranspose, in_size=in_channels, out_size=dec_fm[i],
                                    in_size_bridge=feature_maps[i], z_down=z_down[i], yx_down=yx_down[i],
                                    up_mode=upsample_layer, conv=conv, k_size=ksize(i), act=activation, norm=normalization,
                                    dropout=drop_values[i], nconvs=conv_layers[i], order=conv_block_order)
                else:
                    up = UpBlock(ndim=nd, convtranspose=convtranspose, in_size=in_channels, out_size=dec_fm[i],
                                 in_size_bridge=feature_maps[i], z_down=z_down[i], yx_down=yx_down[i], up_mode=upsample_layer,
                                 conv=conv, k_size=ksize(i), act=activation, norm=normalization, dropout=drop_values[i],
                                 attention_gate=(self.variant == "attention_unet"), nconvs=conv_layers[i],
                                 order=conv_block_order)
                self.up_paths[j].append(up)
                in_channels = dec_fm[i]
        if larger_io:
            self.conv_out = nn.ModuleList([
                ConvBlock(conv=conv, in_size=dec_fm[0], out_size=dec_fm[0], k_size=ksize(0, 2), act=activation,
                          norm=normalization, order=conv_block_order) for _ in range(self.num_decoders)])
        else:
            self.conv_out = None
        self.post_upsampling = None
        self.heads = nn.Sequential()
        for out_ch in output_channels:
            self.heads.append(conv(dec_fm[0], out_ch, kernel_size=1, padding="same"))
        init_weights(self)

        # engine configuration (not part of the reference surface)
        self.engine_dtype: torch.dtype = default_engine_dtype()
        self.conv_impl: int = _lib.IMPL_AUTO

    # ------------------------------------------------------------------------------------------ engine
    def set_engine(self, dtype=None, conv_impl: Optional[int] = None):
        """Choose the activation/compute storage dtype (torch.bfloat16 | float16 | float32) and conv kernel family."""
        if dtype is not None:
            self.engine_dtype = _DTYPES[dtype] if isinstance(dtype, str) else dtype
        if conv_impl is not None:
            self.conv_impl = conv_impl
        return self

    def next_rng_seed(self, device) -> torch.Tensor:
        """Device-resident dropout seed (int64[1]), started from torch.initial_seed() and advanced once per training
        forward with an in-stream add -- a captured CUDA graph therefore draws new masks on every replay."""
        s = self.__dict__.get("_rng_seed")
        if s is None or s.device != torch.device(device):
            s = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device=device)
            self.__dict__["_rng_seed"] = s
        else:
            s.add_(1)
        return s

    # packed-weight cache of eval-mode forwards (see Tape): dropped whenever the weights may change outside autograd's view
    def train(self, mode: bool = True):
        self.__dict__.pop("_pack_cache", None)
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_pack_cache", None)
        return super().load_state_dict(*args, **kwargs)

    def _run(self, tape: Tape, x: TT):
        """Encoder / bottleneck / decoder(s) / heads on the tape.  Returns (pred TT, class TT | None)."""
        if self.conv_in is not None:
            x = self.conv_in.run(tape, x)
        skips: List[TT] = []
        cats: List[List[Optional[TT]]] = [[None] * self.depth for _ in range(self.num_decoders)]
        for i, (down, pool) in enumerate(zip(self.down_path, self.mpooling_layers)):
            ups = [self.up_paths[j][self.depth - 1 - i] for j in range(self.num_decoders)]
            c_skip = _out_channels(down)
            out = None
            for j, up in enumerate(ups):
                cats[j][i] = tape.new(x.data, up.up_channels + c_skip)
            if ups[0].bridge_in_cat:
                # the encoder block writes its result straight into the concat buffer of decoder 0
                out = cats[0][i].slice(ups[0].up_channels, c_skip)
            x = down.run(tape, x, out=out)
            for j, up in enumerate(ups[1:], start=1):
                if up.bridge_in_cat:
                    tape.copy_into(x, cats[j][i].slice(up.up_channels, c_skip))
            skips.append(x)
            x = tape.maxpool(x, _pool_window(pool))
        x_bot = self.bottleneck.run(tape, x)
        feats = []
        for j in range(self.num_decoders):
            x = x_bot
            for n, up in enumerate(self.up_paths[j]):
                lvl = self.depth - 1 - n
                x = up.run(tape, x, skips[lvl], cats[j][lvl])
            feats.append(x)
        if self.conv_out is not None:
            feats = [self.conv_out[j].run(tape, feats[j]) for j in range(self.num_decoders)]
        n_pred = sum(c for c, info in zip(self.output_channels, self.output_channel_info) if "class" not in info)
        n_cls = sum(c for c, info in zip(self.output_channels, self.output_channel_info) if "class" in info)
        pred = tape.new(feats[0].data, n_pred) if n_pred else None
        cls = tape.new(feats[0].data, n_cls) if n_cls else None
        po = co = 0
        for i, head in enumerate(self.heads):
            feat = feats[i] if self.num_decoders > 1 else feats[0]
            c = self.output_channels[i]
            if "class" in self.output_channel_info[i]:
                tape.conv(feat, head, out=cls.slice(co, c))
                co += c
            else:
                tape.conv(feat, head, out=pred.slice(po, c))
                po += c
        return pred, cls

    def _execute(self, x: torch.Tensor, record: bool, x_requires_grad: bool = False):
        if not x.is_cuda:
            raise _lib.B200Error("biapy_b200 models run on CUDA only (no CPU / PyTorch fallback); got a CPU tensor")
        if x.dim() != self.ndim + 2:
            raise ValueError(f"expected a {self.ndim + 2}-D input (N, C, [Z,] Y, X), got {tuple(x.shape)}")
        xs = x.detach()
        if self.ndim == 2:
            xs = xs.unsqueeze(2)
        xcl = xs.permute(0, 2, 3, 4, 1)                     # (N, D, H, W, C): BiaPy's host layout, no copy
        if not xcl.is_contiguous():
            xcl = xcl.contiguous()
        if xcl.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            xcl = xcl.float()
        cache = None
        if not record and not self.training:
            cache = self.__dict__.setdefault("_pack_cache", {})
        tape = Tape(self.engine_dtype, x.device, training=record, conv_impl=self.conv_impl, pack_cache=cache)
        if self.training:
            tape.rng_seed = self.next_rng_seed(x.device)
        if xcl.dtype == self.engine_dtype:
            x_tt = TT(xcl, requires_grad=x_requires_grad)
        else:
            x_tt = TT(torch.empty(xcl.shape, dtype=self.engine_dtype, device=x.device), requires_grad=x_requires_grad)
            ops.convert(xcl, x_tt.data)
        # the zero-initialised accumulators of the pass (channel sums ...) come out of ONE cleared buffer instead of a torch fill
        # kernel each; they are all consumed before the forward returns.  Not inside a CUDA-graph capture of the caller (the
        # arena may grow) and not nested inside a Trainer pass, which has its own.
        arena = None
        if ops.ARENA is None and not torch.cuda.is_current_stream_capturing():
            arena = self.__dict__.setdefault("_arena_fwd", ops.ZeroArena())
            arena.begin(x.device)
        try:
            pred, cls = self._run(tape, x_tt)
        finally:
            if arena is not None:
                arena.end()
        outs, tts = [], []
        for t in (pred, cls):
            if t is None:
                continue
            o32 = torch.empty(t.shape, dtype=torch.float32, device=x.device)
            ops.convert(t.data, o32)
            o = o32.permute(0, 4, 1, 2, 3)
            outs.append(o.squeeze(2) if self.ndim == 2 else o)
            tts.append(t)
        return outs, tape, tts, x_tt

    def _apply_explicit(self, t: torch.Tensor, names: Sequence[str]) -> torch.Tensor:
        """explicit_activations=True (inference only): per-channel head activations on the fp32 output."""
        xs = t.unsqueeze(2) if self.ndim == 2 else t
        cl = xs.permute(0, 2, 3, 4, 1)
        if not cl.is_contiguous():
            cl = cl.contiguous()
        out = torch.empty_like(cl)
        for i, a in enumerate(names):
            a = {"ce_sigmoid": "sigmoid", "linear": "none"}.get(a, a)
            if a in ("ce_softmax", "softmax"):
                raise NotImplementedError("explicit softmax head activations: use the workflow-level activation path")
            ops.scale_shift_act(cl[..., i:i + 1], None, None, a, out[..., i:i + 1])
        o = out.permute(0, 4, 1, 2, 3)
        return o.squeeze(2) if self.ndim == 2 else o

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, x) -> Dict | torch.Tensor:
        """Same contract as the reference ``forward`` (``unet.py:351-445``): input ``(N, C, [Z,] Y, X)``; returns the
        prediction tensor, or a dict with ``"pred"`` / ``"class"`` when class heads exist."""
        params = [p for p in self.parameters()]
        need_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        if need_grad:
            if self.explicit_activations:
                raise NotImplementedError("explicit_activations=True is supported for inference only")
            outs = _NetFn.apply(self, x, *params)
        else:
            outs = self._execute(x, record=False)[0]
        out_dict = {"pred": outs[0]}
        if self.return_class:
            out_dict["class"] = outs[1]
        if self.explicit_activations:
            n_pred = out_dict["pred"].shape[1]
            out_dict["pred"] = self._apply_explicit(out_dict["pred"], self._head_act_names[:n_pred])
            if self.return_class:
                out_dict["class"] = self._apply_explicit(out_dict["class"], self._head_act_names[n_pred:])
        if len(out_dict) == 1:
            return out_dict["pred"]
        if self.return_one_tensor:
            if "class" in out_dict:
                return torch.cat((out_dict["pred"], torch.argmax(out_dict["class"], dim=1).unsqueeze(1)), dim=1)
            return out_dict["pred"]
        return out_dict


def _out_channels(block) -> int:
    if isinstance(block, ResConvBlock):
        return block.shortcut[0].out_channels
    b = block
    while isinstance(b, ConvBlock) and b.nconvs > 1:
        b = b.block[-1]
    return b.block[b._conv_idx].out_channels


def _pool_window(pool: nn.Module):
    k = pool.kernel_size
    return tuple(k) if isinstance(k, (tuple, list)) else (k,) * 2

"""Building blocks of the U-Net family, laid out like ``biapy/models/blocks.py`` but executed by the B200 engine.

Every class keeps the reference's constructor arguments and -- crucially -- the same sub-module nesting, so
``state_dict()`` keys and shapes are identical and BiaPy checkpoints load with ``strict=True``:

=================  ===========================  ======================================================
class              reference                    notes
=================  ===========================  ======================================================
``ConvBlock``      ``blocks.py:25-192``         conv_norm_act / norm_act_conv, ``nconvs`` nesting
``UpBlock``        ``blocks.py:510-668``        ConvTranspose -> norm -> act, optional attention gate
``AttentionBlock`` ``blocks.py:1014-1116``      ``w_x`` carries no norm (reference quirk, replicated)
``ResConvBlock``   ``blocks.py:1194-1459``      post- and pre-activation residual blocks
``ResUpBlock``     ``blocks.py:1462-1655``
=================  ===========================  ======================================================

The torch leaf modules (``nn.Conv3d``, ``nn.GroupNorm`` ...) are *parameter holders only*: nothing here calls
their ``forward``.  Computation goes through ``run(tape, x)``, which drives the CUDA kernels via
:class:`biapy_b200.engine.tape.Tape`.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from ..engine.tape import TT, Tape

# activations whose reference module is built with inplace=True (blocks.py:1986-1998)
_INPLACE_ACTS = ("relu", "leaky_relu", "elu", "silu")


def get_activation(activation: str = "relu") -> nn.Module:
    """Placeholder module per activation name (reference ``blocks.py:1962-1999``); keeps Sequential indices."""
    table = {
        "relu": lambda: nn.ReLU(inplace=True), "tanh": nn.Tanh, "leaky_relu": lambda: nn.LeakyReLU(inplace=True),
        "elu": lambda: nn.ELU(alpha=1.0, inplace=True), "gelu": nn.GELU, "silu": lambda: nn.SiLU(inplace=True),
        "sigmoid": nn.Sigmoid, "softmax": lambda: nn.Softmax(dim=1), "linear": nn.Identity, "softplus": nn.Softplus,
        "none": nn.Identity,
    }
    assert activation in table, "Get unknown activation key {}".format(activation)
    return table[activation]()


def _get_norm(ndim: int, norm: str, channels: int, bn_momentum: float = 0.1) -> nn.Module:
    assert norm in ["bn", "sync_bn", "gn", "in", "none"], "Get unknown normalization layer key {}".format(norm)
    if norm == "gn":
        # The reference call nn.GroupNorm(C, num_groups=8|16) raises TypeError (blocks.py:2124-2125, 2162-2163);
        # the intended layer is GroupNorm(8, C) in 3D and GroupNorm(16, C) in 2D (SURVEY.md finding 1).
        return nn.GroupNorm(8 if ndim == 3 else 16, channels)
    if norm == "in":
        return (nn.InstanceNorm3d if ndim == 3 else nn.InstanceNorm2d)(channels, affine=True, momentum=bn_momentum)
    if norm == "bn":
        return (nn.BatchNorm3d if ndim == 3 else nn.BatchNorm2d)(channels, momentum=bn_momentum)
    if norm == "sync_bn":
        return nn.SyncBatchNorm(channels, momentum=bn_momentum)
    return nn.Identity()


def get_norm_3d(norm: str, out_channels: int, bn_momentum: float = 0.1) -> nn.Module:
    return _get_norm(3, norm, out_channels, bn_momentum)


def get_norm_2d(norm: str, out_channels: int, bn_momentum: float = 0.1) -> nn.Module:
    return _get_norm(2, norm, out_channels, bn_momentum)


def get_decoder_feature_maps(feature_maps: List[int], num_decoders: int, divide_feature_maps: bool) -> List[int]:
    """Reference ``blocks.py:2083-2090``."""
    if num_decoders <= 1 or not divide_feature_maps:
        return list(feature_maps)
    return [max(1, x // num_decoders) for x in feature_maps]


def init_weights(model: nn.Module):
    """Xavier-uniform conv / linear weights, zero biases; transposed convs keep the PyTorch default
    (reference ``blocks.py:2301-2339``)."""
    info = getattr(model, "output_channel_info", [""])
    heads = getattr(model, "heads", None)
    hm_head = None
    if info and "bbox_heatmap" in info and heads is not None and len(heads) > 0:
        hm_head = heads[info.index("bbox_heatmap")]

    def _init(m):
        if isinstance(m, (nn.Conv2d, nn.Conv3d, nn.Linear)):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, -4.59 if (hm_head is not None and m is hm_head) else 0)
        elif isinstance(m, nn.LayerNorm):
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
            if m.weight is not None:
                nn.init.constant_(m.weight, 1.0)

    model.apply(_init)


def _ndim_of(conv) -> int:
    return 2 if conv == nn.Conv2d else 3


def _no_direct_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is executed through its parent biapy_b200 model "
                       "(run(tape, x)); it has no standalone torch forward")


def _norm_or_none(m: nn.Module) -> Optional[nn.Module]:
    return None if isinstance(m, nn.Identity) else m


def _act_name(m: Optional[nn.Module]) -> Optional[str]:
    if m is None:
        return None
    return {nn.ReLU: "relu", nn.Tanh: "tanh", nn.LeakyReLU: "leaky_relu", nn.ELU: "elu", nn.GELU: "gelu", nn.SiLU: "silu",
            nn.Sigmoid: "sigmoid", nn.Identity: "none", nn.Softplus: "softplus"}[type(m)]


class ConvBlock(nn.Module):
    """Conv (+norm +act +dropout), possibly repeated ``nconvs`` times (reference ``blocks.py:25-192``)."""

    forward = _no_direct_forward

    def __init__(self, conv, in_size, out_size, k_size, padding: int | str = "same", stride=1, bias=True, act=None,
                 norm="none", dropout=0, se_block=False, nconvs=1, order="conv_norm_act"):
        super().__init__()
        if nconvs < 1:
            raise ValueError(f"'nconvs' must be >= 1, but {nconvs} was given")
        if order not in ("conv_norm_act", "norm_act_conv"):
            raise ValueError(f"'order' must be 'conv_norm_act' or 'norm_act_conv', but {order!r} was given")
        if se_block:
            raise NotImplementedError("Squeeze-and-Excitation blocks are outside the B200 hot path")
        if padding != "same" or stride != 1:
            raise NotImplementedError("the B200 engine implements stride-1 'same' convolutions (the U-Net path)")
        self.order = order
        self.nconvs = nconvs
        if nconvs > 1:
            self.block = nn.Sequential(*[
                ConvBlock(conv, in_size if i == 0 else out_size, out_size, k_size, bias=bias, act=act, norm=norm,
                          dropout=dropout, order=order) for i in range(nconvs)])
            return
        ndim = _ndim_of(conv)
        self._conv_idx = self._norm_idx = self._act_idx = None
        self.dropout_p = float(dropout)
        layers: List[nn.Module] = []

        def add(m):
            layers.append(m)
            return len(layers) - 1

        if order == "norm_act_conv":
            if norm != "none":
                self._norm_idx = add(_get_norm(ndim, norm, in_size))
            if act:
                self._act_idx = add(get_activation(act))
            self._conv_idx = add(conv(in_size, out_size, kernel_size=k_size, padding="same", stride=1, bias=bias))
        else:
            self._conv_idx = add(conv(in_size, out_size, kernel_size=k_size, padding="same", stride=1, bias=bias))
            if norm != "none":
                self._norm_idx = add(_get_norm(ndim, norm, out_size))
            if act:
                self._act_idx = add(get_activation(act))
        if dropout > 0:
            add(nn.Dropout(dropout))
        self.block = nn.Sequential(*layers)

    def _parts(self):
        b = self.block
        norm = _norm_or_none(b[self._norm_idx]) if self._norm_idx is not None else None
        act = _act_name(b[self._act_idx]) if self._act_idx is not None else None
        return b[self._conv_idx], norm, act

    def run(self, tape: Tape, x: TT, out: Optional[TT] = None, into: Optional[TT] = None, lazy: bool = False) -> TT:
        """`out`: write the block result there.  `into`: accumulate the (bare) final convolution into it.  `lazy`: the caller
        feeds the result to a convolution and nothing else (Tape.norm_act)."""
        if self.nconvs > 1:
            for i, sub in enumerate(self.block):
                last = i == self.nconvs - 1
                x = sub.run(tape, x, out=out if last else None, into=into if last else None, lazy=lazy if last else True)
            return x
        conv, norm, act = self._parts()
        if self.dropout_p > 0 and self.training:
            # nn.Dropout is the last layer of the block (reference blocks.py:162-163); `into` is never set here because
            # a block with dropout does not end with a bare convolution
            assert into is None
            if self.order == "norm_act_conv":
                y = tape.conv(tape.norm_act(x, norm, act), conv)
            elif norm is None and (act in (None, "none")):
                y = tape.conv(x, conv)
            else:
                y = tape.norm_act(tape.conv(x, conv, stats=norm is not None), norm, act)
            return tape.dropout(y, self.dropout_p, out=out)
        if self.order == "norm_act_conv":
            h = tape.norm_act(x, norm, act, lazy=True)      # consumed by the convolution only: fusable on its operand path
            if into is not None:
                return tape.conv(h, conv, out=into, accumulate=True)
            return tape.conv(h, conv, out=out)
        if norm is None and (act in (None, "none")):
            if into is not None:
                return tape.conv(x, conv, out=into, accumulate=True)
            return tape.conv(x, conv, out=out)
        assert into is None, "only a bare convolution can be accumulated into a residual"
        # conv -> norm: the channel sums of the normalisation come out of the convolution's epilogue when the kernel has one.
        # `lazy`: inside a residual block the result only feeds the next convolution, which can apply norm + act on its operand path
        return tape.norm_act(tape.conv(x, conv, stats=norm is not None), norm, act, out=out, lazy=lazy)

    @property
    def ends_with_bare_conv(self) -> bool:
        if self.nconvs > 1:
            return self.block[-1].ends_with_bare_conv
        if self.order == "norm_act_conv":
            return not (self.dropout_p > 0)
        return self._norm_idx is None and self._act_idx is None and not (self.dropout_p > 0)


class AttentionBlock(nn.Module):
    """Attention gate (reference ``blocks.py:1014-1116``): ``psi = sigmoid(norm(conv(relu(w_g(g) + w_x(x))))); out = psi*x``.

    The reference appends the ``w_x`` norm to the already-consumed ``w_g`` list (``:1063-1072``), so ``w_x`` is a
    bare 1x1 convolution; replicated for state_dict and numerical parity."""

    forward = _no_direct_forward

    def __init__(self, conv, in_size, out_size, norm="none", in_size_bridge=None):
        super().__init__()
        ndim = _ndim_of(conv)
        if in_size_bridge is None:
            in_size_bridge = in_size
        w_g = [conv(in_size, out_size, kernel_size=1, stride=1, padding=0, bias=True)]
        if norm != "none":
            w_g.append(_get_norm(ndim, norm, out_size))
        self.w_g = nn.Sequential(*w_g)
        self.w_x = nn.Sequential(conv(in_size_bridge, out_size, kernel_size=1, stride=1, padding=0, bias=True))
        psi = [conv(out_size, 1, kernel_size=1, stride=1, padding=0, bias=True)]
        if norm != "none":
            psi.append(_get_norm(ndim, norm, 1))
        psi.append(nn.Sigmoid())
        self.psi = nn.Sequential(*psi)
        self.relu = nn.ReLU(inplace=True)
        self._has_norm = norm != "none"

    def run(self, tape: Tape, g: TT, x: TT, out: Optional[TT] = None) -> TT:
        g1 = tape.conv(g, self.w_g[0], stats=self._has_norm)
        if self._has_norm:
            g1 = tape.norm_act(g1, self.w_g[1], None)
        x1 = tape.conv(x, self.w_x[0])
        s = tape.add_relu(g1, x1)
        p = tape.conv(s, self.psi[0])
        p = tape.norm_act(p, self.psi[1] if self._has_norm else None, "sigmoid")
        return tape.gate(p, x, out=out)


class UpBlock(nn.Module):
    """ConvTranspose(in->out) -> norm -> act, concat with the (optionally gated) skip, ConvBlock
    (reference ``blocks.py:510-668``)."""

    forward = _no_direct_forward

    def __init__(self, ndim, convtranspose, in_size, out_size, z_down, up_mode, conv, k_size, yx_down=2, act=None,
                 norm="none", dropout=0, attention_gate=False, se_block=False, nconvs=2, order="conv_norm_act",
                 in_size_bridge=None):
        super().__init__()
        self.ndim = ndim
        if in_size_bridge is None:
            in_size_bridge = out_size
        self.out_size, self.in_size_bridge = out_size, in_size_bridge
        mpool = (z_down, yx_down, yx_down) if ndim == 3 else (yx_down, yx_down)
        self.up_mode, self.mpool = up_mode, mpool
        if up_mode == "convtranspose":
            layers: List[nn.Module] = [convtranspose(in_size, out_size, kernel_size=mpool, stride=mpool)]
        elif up_mode == "upsampling":
            # reference blocks.py:604-606: nn.Upsample(bi/trilinear) followed by a 1x1 convolution
            layers = [nn.Upsample(mode="bilinear" if ndim == 2 else "trilinear", scale_factor=mpool),
                      conv(in_size, out_size, kernel_size=1)]
        else:
            raise ValueError(f"unknown up_mode {up_mode!r}")
        self._norm_idx = self._act_idx = None
        if norm != "none":
            layers.append(_get_norm(ndim, norm, out_size))
            self._norm_idx = len(layers) - 1
        if act is not None:
            layers.append(get_activation(act))
            self._act_idx = len(layers) - 1
        self.up = nn.Sequential(*layers)
        self.attention_gate = AttentionBlock(conv=conv, in_size=out_size, out_size=out_size // 2, norm=norm,
                                             in_size_bridge=in_size_bridge) if attention_gate else None
        self.conv_block = ConvBlock(conv=conv, in_size=out_size + in_size_bridge, out_size=out_size, k_size=k_size, act=act,
                                    norm=norm, dropout=dropout, se_block=se_block, nconvs=nconvs, order=order)

    # channels the up-sampled tensor occupies in the concat buffer
    up_channels = property(lambda self: self.out_size)
    bridge_in_cat = property(lambda self: self.attention_gate is None)

    def run(self, tape: Tape, x: TT, bridge: TT, cat: TT) -> TT:
        up_slot = cat.slice(0, self.out_size)
        norm = _norm_or_none(self.up[self._norm_idx]) if self._norm_idx is not None else None
        act = _act_name(self.up[self._act_idx]) if self._act_idx is not None else None
        plain = norm is None and act in (None, "none")
        if self.up_mode == "convtranspose":
            if plain:
                tape.convT(x, self.up[0], out=up_slot)
            else:
                tape.norm_act(tape.convT(x, self.up[0]), norm, act, out=up_slot)
        else:
            # Upsample then 1x1 conv == 1x1 conv then Upsample (both linear, interpolation weights sum to one, so the bias
            # passes through): the convolution runs at the coarse resolution, 1/s^3 of the reference's work
            t = tape.conv(x, self.up[1])
            if plain:
                tape.upsample_linear(t, self.mpool, out=up_slot)
            else:
                tape.norm_act(tape.upsample_linear(t, self.mpool), norm, act, out=up_slot)
        if self.attention_gate is not None:
            self.attention_gate.run(tape, up_slot, bridge, out=cat.slice(self.out_size, self.in_size_bridge))
        return self.conv_block.run(tape, cat)


class ResConvBlock(nn.Module):
    """Residual block ``block(x) + shortcut(x)`` (reference ``blocks.py:1194-1459``)."""

    forward = _no_direct_forward

    def __init__(self, conv, in_size, out_size, k_size, act=None, norm="none", dropout=0,
                 skip_k_size: int | Tuple[int, ...] = 1, skip_norm="none", first_block=False, se_block=False,
                 extra_conv=False, nconvs=2, order="conv_norm_act"):
        super().__init__()
        if nconvs < 1:
            raise ValueError(f"'nconvs' must be >= 1, but {nconvs} was given")
        if order not in ("conv_norm_act", "norm_act_conv"):
            raise ValueError(f"'order' must be 'conv_norm_act' or 'norm_act_conv', but {order!r} was given")
        if se_block or extra_conv or skip_norm != "none":
            raise NotImplementedError("se_block / extra_conv / skip_norm are outside the B200 hot path")
        ndim = _ndim_of(conv)
        self.order = order
        self.pre_conv = None
        self._pre_norm_idx = self._pre_act_idx = None
        self._act = act
        self._norm = norm
        layers: List[nn.Module] = []
        if order == "norm_act_conv":
            # reference _build_pre_activation (blocks.py:1389-1432): every conv is a norm->act->conv ConvBlock
            layers.append(ConvBlock(conv, in_size, out_size, k_size, act=act, norm=norm, dropout=dropout, order=order))
            for _ in range(max(0, nconvs - 1)):
                layers.append(ConvBlock(conv, out_size, out_size, k_size, act=act, norm=norm, dropout=dropout, order=order))
        else:
            if not first_block:
                if norm != "none":
                    layers.append(_get_norm(ndim, norm, in_size))
                    self._pre_norm_idx = len(layers) - 1
                if act is not None:
                    layers.append(get_activation(act))
                    self._pre_act_idx = len(layers) - 1
            layers.append(ConvBlock(conv, in_size, out_size, k_size, act=act, norm=norm, dropout=dropout))
            for _ in range(max(0, nconvs - 2)):
                layers.append(ConvBlock(conv, out_size, out_size, k_size, act=act, norm=norm, dropout=dropout))
            if nconvs >= 2:
                layers.append(ConvBlock(conv, out_size, out_size, k_size))
        shortcut = nn.Sequential(conv(in_size, out_size, kernel_size=skip_k_size, padding="same"))
        if order == "norm_act_conv":      # registration order of the reference (blocks.py:1416-1427): shortcut, then block
            self.shortcut = shortcut
            self.block = nn.Sequential(*layers)
        else:                             # blocks.py:1368-1378: block, then shortcut
            self.block = nn.Sequential(*layers)
            self.shortcut = shortcut
        self.se_block = nn.Identity()

    def run(self, tape: Tape, x: TT, out: Optional[TT] = None) -> TT:
        h = x
        shortcut_in = x
        first = 0
        if self.order == "conv_norm_act" and (self._pre_norm_idx is not None or self._pre_act_idx is not None):
            norm = self.block[self._pre_norm_idx] if self._pre_norm_idx is not None else None
            act = _act_name(self.block[self._pre_act_idx]) if self._pre_act_idx is not None else None
            h = tape.norm_act(x, norm, act, lazy=norm is not None)      # feeds the first convolution only
            first = max(i for i in (self._pre_norm_idx, self._pre_act_idx) if i is not None) + 1
            # inplace=True activation applied straight to the block input (no norm in front) also changes what the
            # shortcut sees in the reference (blocks.py:1458 evaluates block(x) first)
            if norm is None and act in _INPLACE_ACTS:
                shortcut_in = h
        convs = list(self.block)[first:]
        res = tape.conv(shortcut_in, self.shortcut[0], out=out)
        if convs[-1].ends_with_bare_conv:
            for cb in convs[:-1]:
                h = cb.run(tape, h, lazy=True)
            return convs[-1].run(tape, h, into=res)
        for i, cb in enumerate(convs):
            h = cb.run(tape, h, lazy=i + 1 < len(convs))
        return tape.add(h, res, out=res)


class ResUpBlock(nn.Module):
    """ConvTranspose(in->in), concat with the skip, ResConvBlock (reference ``blocks.py:1462-1655``)."""

    forward = _no_direct_forward

    def __init__(self, ndim, convtranspose, in_size, out_size, in_size_bridge, z_down, up_mode, conv, k_size, yx_down=2,
                 act=None, norm="none", skip_k_size: int | Tuple[int, ...] = 1, skip_norm="none", dropout=0, se_block=False,
                 extra_conv=False, nconvs=2, order="conv_norm_act"):
        super().__init__()
        self.ndim = ndim
        self.in_size, self.in_size_bridge = in_size, in_size_bridge
        mpool = (z_down, yx_down, yx_down) if ndim == 3 else (yx_down, yx_down)
        self.up_mode, self.mpool = up_mode, mpool
        if up_mode == "convtranspose":
            self.up = convtranspose(in_size, in_size, kernel_size=mpool, stride=mpool)
        elif up_mode == "upsampling":
            self.up = nn.Upsample(mode="bilinear" if ndim == 2 else "trilinear", scale_factor=mpool)   # reference blocks.py:1608
        else:
            raise ValueError(f"unknown up_mode {up_mode!r}")
        self.conv_block = ResConvBlock(conv=conv, in_size=in_size + in_size_bridge, out_size=out_size, k_size=k_size, act=act,
                                       norm=norm, dropout=dropout, skip_k_size=skip_k_size, skip_norm=skip_norm,
                                       se_block=se_block, extra_conv=extra_conv, nconvs=nconvs, order=order)

    up_channels = property(lambda self: self.in_size)
    bridge_in_cat = True

    def run(self, tape: Tape, x: TT, bridge: TT, cat: TT) -> TT:
        if self.up_mode == "convtranspose":
            tape.convT(x, self.up, out=cat.slice(0, self.in_size))
        else:
            tape.upsample_linear(x, self.mpool, out=cat.slice(0, self.in_size))
        return self.conv_block.run(tape, cat)

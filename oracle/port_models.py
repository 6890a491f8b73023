"""TEST INFRASTRUCTURE ONLY -- CPU oracle for BiaPy's U_Net / ResUNet / Attention_U_Net forward (and, through
torch autograd on the same functional graph, backward).

A functional fp32 restatement driven by a reference-layout ``state_dict`` (same key names as the reference
classes), so the same weights feed the reference, this oracle and the CUDA engine.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Reference sites restated:
  * ConvBlock            biapy/models/blocks.py:25-192   (conv_norm_act and norm_act_conv orders, nconvs nesting)
  * UpBlock              biapy/models/blocks.py:510-668
  * AttentionBlock       biapy/models/blocks.py:1014-1116 (w_x has NO norm: the reference appends it to the
                         already-consumed w_g list, :1063-1072 -- replicated)
  * ResConvBlock         biapy/models/blocks.py:1194-1459 (post- and pre-activation variants; in-place activation
                         aliasing when norm == 'none', see _res_block)
  * ResUpBlock           biapy/models/blocks.py:1462-1655
  * norms / activations  biapy/models/blocks.py:1962-1999, 2092-2165  ('gn' = GroupNorm(8|16, C): intended
                         semantics of the broken call at :2124-2125/:2162-2163, SURVEY.md finding 1)
  * U_Net.forward        biapy/models/unet.py:351-445; ResUNet.forward resunet.py:352-446;
                         Attention_U_Net.forward attention_unet.py:366-459
  * head activations     biapy/engine/base_workflow.py:1367-1470 (apply_model_activations)
  * losses               biapy/engine/metrics.py:493-586 (CrossEntropyLoss_wrapper), :2265-2286 (n2v_loss_mse)

The arithmetic itself (conv / norm / pool) is PyTorch ATen on CPU, as in the reference (torch 2.11 here; the
reference pins torch>=2.12,<2.13 -- CPU conv summation-order drift is far below the 1e-3 parity bar).
Pinned against the reference classes themselves in tests/test_oracle_vs_reference.py (container only) and
through the golden fixtures of tests/golden/.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- leaf ops
def _act(name, x):
    if name is None:
        return x
    name = name.lower()
    if name == "relu":
        return F.relu(x)
    if name == "elu":
        return F.elu(x, alpha=1.0)
    if name == "silu":
        return F.silu(x)
    if name == "leaky_relu":
        return F.leaky_relu(x, 0.01)
    if name == "gelu":
        return F.gelu(x)
    if name == "tanh":
        return torch.tanh(x)
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "softplus":
        return F.softplus(x)
    if name in ("linear", "none"):
        return x
    raise ValueError(name)


def _has_act(name):
    # nn.Identity modules still occupy a slot in the reference nn.Sequential, so `act="none"` counts as a layer
    return bool(name)


class Oracle:
    """Functional network bound to a state_dict."""

    def __init__(self, arch: str, sd: Dict[str, torch.Tensor], *, image_shape, activation="elu",
                 feature_maps=(16, 32, 64, 128, 256), normalization="none", k_size=3,
                 upsample_layer="convtranspose", yx_down=None, z_down=None, output_channels=(1,),
                 isotropy=True, larger_io=False, conv_layers=None, conv_block_order="conv_norm_act",
                 training=False, **_ignored):
        self.arch = arch.lower()
        self.sd = sd
        self.ndim = 3 if len(image_shape) == 4 else 2
        self.act = activation.lower()
        self.fm = list(feature_maps)
        self.depth = len(self.fm) - 1
        self.norm = normalization
        self.k = k_size
        self.up_mode = upsample_layer
        self.yx = list(yx_down) if yx_down is not None else [2] * self.depth
        self.z = list(z_down) if z_down is not None else [2] * self.depth
        self.out_ch = list(output_channels)
        self.iso = [isotropy] * len(self.fm) if isinstance(isotropy, bool) else list(isotropy)
        self.larger_io = larger_io
        self.nconvs = list(conv_layers) if conv_layers is not None else [2] * len(self.fm)
        self.order = conv_block_order
        self.training = training

    # -- primitives ---------------------------------------------------------------------------------
    def _conv(self, x, key, same=True):
        w = self.sd[key + ".weight"]
        b = self.sd.get(key + ".bias")
        pad = [k // 2 for k in w.shape[2:]] if same else 0
        return (F.conv3d if self.ndim == 3 else F.conv2d)(x, w, b, padding=pad)

    def _convT(self, x, key, stride):
        f = F.conv_transpose3d if self.ndim == 3 else F.conv_transpose2d
        return f(x, self.sd[key + ".weight"], self.sd.get(key + ".bias"), stride=stride)

    def _norm(self, x, key, kind=None):
        kind = kind or self.norm
        if kind == "none":
            return x
        w, b = self.sd[key + ".weight"], self.sd[key + ".bias"]
        if kind == "gn":
            return F.group_norm(x, 8 if self.ndim == 3 else 16, w, b, eps=1e-5)
        if kind == "in":
            return F.instance_norm(x, None, None, w, b, use_input_stats=True, eps=1e-5)
        if kind in ("bn", "sync_bn"):
            rm, rv = self.sd[key + ".running_mean"], self.sd[key + ".running_var"]
            return F.batch_norm(x, rm, rv, w, b, training=self.training, momentum=0.1, eps=1e-5)
        raise ValueError(kind)

    def _kernel(self, level, extra=0):
        k = self.k + extra
        if self.ndim == 2:
            return (k, k)
        return (k, k, k) if self.iso[level] else (1, k, k)

    def _pool(self, level):
        return (self.z[level], self.yx[level], self.yx[level]) if self.ndim == 3 else (self.yx[level],) * 2

    # -- blocks -------------------------------------------------------------------------------------
    def _single_conv_block(self, x, pfx, norm, act, order):
        """ConvBlock with nconvs == 1 (blocks.py:146-166)."""
        if order == "norm_act_conv":
            i = 0
            if norm != "none":
                x = self._norm(x, f"{pfx}.block.0", norm)
                i += 1
            if _has_act(act):
                x = _act(act, x)
                i += 1
            return self._conv(x, f"{pfx}.block.{i}")
        x = self._conv(x, f"{pfx}.block.0")
        if norm != "none":
            x = self._norm(x, f"{pfx}.block.1", norm)
        return _act(act, x)

    def _conv_block(self, x, pfx, nconvs, norm=None, act="__default__", order=None):
        norm = self.norm if norm is None else norm
        act = self.act if act == "__default__" else act
        order = order or self.order
        if nconvs > 1:                                         # blocks.py:124-144
            for i in range(nconvs):
                x = self._single_conv_block(x, f"{pfx}.block.{i}", norm, act, order)
            return x
        return self._single_conv_block(x, pfx, norm, act, order)

    def _res_block(self, x, pfx, nconvs, first_block):
        """ResConvBlock.forward: block(x) + shortcut(x) (blocks.py:1456-1459)."""
        if self.order == "norm_act_conv":                      # _build_pre_activation, blocks.py:1389-1432
            h = x
            for i in range(max(1, nconvs)):
                h = self._single_conv_block(h, f"{pfx}.block.{i}", self.norm, self.act, "norm_act_conv")
            return h + self._conv(x, f"{pfx}.shortcut.0")
        h = x
        i = 0
        shortcut_in = x
        if not first_block:
            if self.norm != "none":
                h = self._norm(h, f"{pfx}.block.{i}")
                i += 1
            if _has_act(self.act):
                h = _act(self.act, h)
                i += 1
                # The reference activations are inplace=True (blocks.py:1987-1992).  With norm == 'none' the
                # activation therefore overwrites the block input before `self.shortcut(x)` reads it.
                if self.norm == "none" and self.act in ("relu", "leaky_relu", "elu", "silu"):
                    shortcut_in = h
        h = self._single_conv_block(h, f"{pfx}.block.{i}", self.norm, self.act, "conv_norm_act")
        i += 1
        for _ in range(max(0, nconvs - 2)):
            h = self._single_conv_block(h, f"{pfx}.block.{i}", self.norm, self.act, "conv_norm_act")
            i += 1
        if nconvs >= 2:
            h = self._single_conv_block(h, f"{pfx}.block.{i}", "none", None, "conv_norm_act")
        return h + self._conv(shortcut_in, f"{pfx}.shortcut.0")

    def _attention(self, g, x, pfx):
        """AttentionBlock.forward (blocks.py:1112-1116)."""
        g1 = self._conv(g, f"{pfx}.w_g.0", same=False)
        if self.norm != "none":
            g1 = self._norm(g1, f"{pfx}.w_g.1")
        x1 = self._conv(x, f"{pfx}.w_x.0", same=False)
        psi = F.relu(g1 + x1)
        psi = self._conv(psi, f"{pfx}.psi.0", same=False)
        if self.norm != "none":
            psi = self._norm(psi, f"{pfx}.psi.1")
        return torch.sigmoid(psi) * x

    def _upsample(self, x, level):
        return F.interpolate(x, scale_factor=tuple(float(s) for s in self._pool(level)),
                             mode="trilinear" if self.ndim == 3 else "bilinear")

    def _up_block(self, x, bridge, pfx, level, attention):
        """UpBlock.forward (blocks.py:659-668)."""
        if self.up_mode == "convtranspose":
            up = self._convT(x, f"{pfx}.up.0", self._pool(level))
            i = 1
        else:
            up = self._conv(self._upsample(x, level), f"{pfx}.up.1", same=False)
            i = 2
        if self.norm != "none":
            up = self._norm(up, f"{pfx}.up.{i}")
        up = _act(self.act, up)
        other = self._attention(up, bridge, f"{pfx}.attention_gate") if attention else bridge
        return self._conv_block(torch.cat([up, other], 1), f"{pfx}.conv_block", self.nconvs[level])

    def _res_up_block(self, x, bridge, pfx, level):
        """ResUpBlock.forward (blocks.py:1652-1655)."""
        up = self._convT(x, f"{pfx}.up", self._pool(level)) if self.up_mode == "convtranspose" else self._upsample(x, level)
        return self._res_block(torch.cat([up, bridge], 1), f"{pfx}.conv_block", self.nconvs[level], first_block=False)

    # -- network ------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        res = self.arch == "resunet"
        att = self.arch == "attention_unet"
        if self.larger_io:
            x = self._conv_block(x, "conv_in", 1)
        skips = []
        for i in range(self.depth):
            x = self._res_block(x, f"down_path.{i}", self.nconvs[i], first_block=(i == 0)) if res \
                else self._conv_block(x, f"down_path.{i}", self.nconvs[i])
            skips.append(x)
            x = (F.max_pool3d if self.ndim == 3 else F.max_pool2d)(x, self._pool(i))
        x = self._res_block(x, "bottleneck", self.nconvs[-1], first_block=False) if res \
            else self._conv_block(x, "bottleneck", self.nconvs[-1])
        for n, i in enumerate(range(self.depth - 1, -1, -1)):
            pfx = f"up_paths.0.{n}"
            x = self._res_up_block(x, skips[i], pfx, i) if res else self._up_block(x, skips[i], pfx, i, att)
        if self.larger_io:
            x = self._conv_block(x, "conv_out.0", 1)
        outs = [self._conv(x, f"heads.{h}") for h in range(len(self.out_ch))]
        return torch.cat(outs, 1)


def forward(arch: str, state_dict, x: torch.Tensor, **kwargs) -> torch.Tensor:
    return Oracle(arch, state_dict, **kwargs).forward(x)


# ------------------------------------------------------------------------------- head activations / losses
def apply_head_activations(pred: torch.Tensor, head_activations: Sequence[str], training: bool) -> torch.Tensor:
    """Restatement of Base_Workflow.apply_model_activations for a plain tensor whose channels are all
    'class' channels or all per-channel activations (base_workflow.py:1396-1457)."""
    acts = [a.lower() for a in head_activations]
    outs: List[torch.Tensor] = []
    i = 0
    C = pred.shape[1]
    while i < C:
        a = acts[i]
        if a == "linear" or (training and a in ("ce_sigmoid", "ce_softmax")):
            outs.append(pred[:, i:i + 1])
            i += 1
        elif a == "ce_softmax":
            j = i
            while j < C and acts[j] == "ce_softmax":
                j += 1
            outs.append(torch.softmax(pred[:, i:j], dim=1))
            i = j
        else:
            outs.append(_act("sigmoid" if a == "ce_sigmoid" else a, pred[:, i:i + 1]))
            i += 1
    return torch.cat(outs, 1)


def bce_with_logits_loss(pred, target):
    """CrossEntropyLoss_wrapper for N_CLASSES <= 2 (metrics.py:544-546, 577-580): BCEWithLogits, mean."""
    return F.binary_cross_entropy_with_logits(pred, target)


def ce_loss(pred, target):
    """CrossEntropyLoss_wrapper for N_CLASSES > 2 (metrics.py:581-586): CE(pred, target[:,0].long())."""
    return F.cross_entropy(pred, target[:, 0].long())


def n2v_loss_mse(y_pred, y_true):
    """metrics.py:2265-2286: target = y_true[:, :C], mask = y_true[:, C:]; sum((t - y*m)^2) / sum(m)."""
    C = y_pred.shape[1]
    target, mask = y_true[:, :C], y_true[:, C:]
    return torch.sum(torch.square(target - y_pred * mask)) / torch.sum(mask)

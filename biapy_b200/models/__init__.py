"""Model registry of the B200 engine: the hot-path subset of ``biapy/models/__init__.py``.

``build_model`` keeps the reference's 5-argument signature and 7-tuple return (``models/__init__.py:44-50, 487``)
for the architectures the hot path covers (``unet``, ``resunet``, ``attention_unet``); other architectures are
not re-implemented and raise ``NotImplementedError``.
"""
from __future__ import annotations

import importlib
from typing import Any, Dict, List, Tuple

_CLASS_BY_ARCH = {"unet": "U_Net", "resunet": "ResUNet", "attention_unet": "Attention_U_Net"}


def _get(node, dotted: str, default=None):
    cur = node
    for part in dotted.split("."):
        if isinstance(cur, dict):
            if part not in cur:
                return default
            cur = cur[part]
        else:
            if not hasattr(cur, part):
                return default
            cur = getattr(cur, part)
    return cur


def model_kwargs_from_cfg(cfg, output_channels, output_channel_info, head_activations) -> Dict[str, Any]:
    """The kwargs ``build_model`` hands to the U-Net classes (reference ``models/__init__.py:120-143, 172-174``).
    `cfg` may be a yacs CfgNode or a plain nested dict with the same keys."""
    ndim = 3 if _get(cfg, "PROBLEM.NDIM", "2D") == "3D" else 2
    fm = list(_get(cfg, "MODEL.FEATURE_MAPS", [16, 32, 64, 128, 256]))
    depth = len(fm) - 1
    z_down = list(_get(cfg, "MODEL.Z_DOWN", [0] * depth))
    yx_down = list(_get(cfg, "MODEL.YX_DOWN", [0] * depth))
    # derived defaults of check_configuration.py:2685-2690, 2737-2742: all-zero -> 2 per level
    if all(v == 0 for v in z_down):
        z_down = [2] * depth
    if all(v == 0 for v in yx_down):
        yx_down = [2] * depth
    for key, v in (("MODEL.Z_DOWN", z_down), ("MODEL.YX_DOWN", yx_down)):
        if len(v) != depth:                                                   # check_configuration.py:2706-2707, 2758-2759
            raise ValueError(f"'MODEL.FEATURE_MAPS' length minus one and '{key}' length must be equal")
    # every level must divide the patch evenly and leave more than two voxels (check_configuration.py:3140-3205)
    patch = tuple(_get(cfg, "DATA.PATCH_SIZE"))
    arch = str(_get(cfg, "MODEL.ARCHITECTURE", "unet")).lower()
    cur_z = patch[0] if ndim == 3 else 1
    cur_yx = list(patch[1:-1]) if ndim == 3 else list(patch[:-1])
    for i in range(depth):
        yf, zf = yx_down[i], (z_down[i] if ndim == 3 else 1)
        if any(d % yf != 0 or d <= 2 for d in cur_yx) or (ndim == 3 and (cur_z % zf != 0 or cur_z <= 2)):
            m = (f"The 'DATA.PATCH_SIZE' provided is not divisible by the downsampling factor at level {i} of the {arch}. "
                 "You can:\n 1) Reduce the number of levels (by reducing 'cfg.MODEL.FEATURE_MAPS' array length)\n"
                 " 2) Increase 'DATA.PATCH_SIZE'")
            if ndim == 3:
                m += ("\n 3) If the Z axis is the problem (often smaller due to resolution), you can tune 'MODEL.Z_DOWN' to not "
                      "downsample the Z axis in all levels.")
            raise ValueError(m)
        cur_yx = [d // yf for d in cur_yx]
        cur_z //= zf
    iso = _get(cfg, "MODEL.ISOTROPY", [True] * len(fm))
    if not isinstance(iso, bool) and all(x is True or x == 1 for x in iso):
        iso = [True] * len(fm)                                                # :2761-2763: all-True follows the feature maps
    # MODEL.DROPOUT_VALUES (:2677-2683): an all-zero list follows the feature maps, anything else must match them
    drop = list(_get(cfg, "MODEL.DROPOUT_VALUES", [0.0] * len(fm)))
    if len(drop) != len(fm):
        if all(x == 0 for x in drop):
            drop = [0.0] * len(fm)
        elif any(not (0 <= x <= 1) for x in drop):
            raise ValueError("'MODEL.DROPOUT_VALUES' not in [0, 1] range")
        else:
            raise ValueError("'MODEL.FEATURE_MAPS' and 'MODEL.DROPOUT_VALUES' lengths must be equal")
    # MODEL.CONV_LAYERS (:2776-2790): empty -> 2 per level, one value or a uniform list -> broadcast to the levels
    conv_layers = list(_get(cfg, "MODEL.CONV_LAYERS", [2] * len(fm)))
    if len(conv_layers) == 0:
        conv_layers = [2] * len(fm)
    elif len(conv_layers) != len(fm):
        if len(set(conv_layers)) != 1:
            raise ValueError("'MODEL.FEATURE_MAPS' and 'MODEL.CONV_LAYERS' lengths must be equal")
        conv_layers = [conv_layers[0]] * len(fm)
    if any(x < 1 for x in conv_layers):
        raise ValueError("'MODEL.CONV_LAYERS' values must be greater than or equal to 1")
    return dict(
        image_shape=tuple(_get(cfg, "DATA.PATCH_SIZE")),
        activation=str(_get(cfg, "MODEL.ACTIVATION", "elu")).lower(),
        feature_maps=fm,
        drop_values=drop,
        normalization=_get(cfg, "MODEL.NORMALIZATION", "in"),
        k_size=_get(cfg, "MODEL.KERNEL_SIZE", 3),
        upsample_layer=_get(cfg, "MODEL.UPSAMPLE_LAYER", "convtranspose"),
        yx_down=yx_down,
        z_down=z_down if ndim == 3 else [2] * depth,
        output_channels=list(output_channels),
        output_channel_info=list(output_channel_info),
        head_activations=list(head_activations),
        explicit_activations=False,
        contrast=bool(_get(cfg, "LOSS.CONTRAST.ENABLE", False)),
        contrast_proj_dim=_get(cfg, "LOSS.CONTRAST.PROJ_DIM", 256),
        separated_decoders=bool(_get(cfg, "MODEL.SEPARATED_DECODERS", False)),
        divide_decoder_feature_maps=bool(_get(cfg, "MODEL.DIVIDE_DECODER_FEATURE_MAPS", False)),
        isotropy=list(iso) if not isinstance(iso, bool) else iso,
        larger_io=bool(_get(cfg, "MODEL.LARGER_IO", False)),
        conv_layers=conv_layers,
        conv_block_order=_get(cfg, "MODEL.CONV_BLOCK_ORDER", "conv_norm_act"),
    )


def build_model(cfg, output_channels: List[int], output_channel_info: List[str], head_activations: List[str], device) -> Tuple:
    """Drop-in for ``biapy.models.build_model`` on the hot-path architectures.

    Returns ``(model, class_name, collected_sources, import_lines, scanned_files, args, network_stride)`` like the
    reference (``models/__init__.py:487``); the three source-extraction fields are empty (they only feed the
    BioImage Model Zoo exporter, out of scope).
    """
    arch = str(_get(cfg, "MODEL.ARCHITECTURE", "unet")).lower()
    if arch not in _CLASS_BY_ARCH:
        raise NotImplementedError(f"MODEL.ARCHITECTURE={arch!r} is outside the B200 hot path "
                                  f"(supported: {sorted(_CLASS_BY_ARCH)})")
    mdl = importlib.import_module("biapy_b200.models." + arch)
    cls_name = _CLASS_BY_ARCH[arch]
    args = model_kwargs_from_cfg(cfg, output_channels, output_channel_info, head_activations)
    model = getattr(mdl, cls_name)(**args).to(device)
    ndim = 3 if len(args["image_shape"]) == 4 else 2
    return model, cls_name, {}, [], [], args, [1] * ndim

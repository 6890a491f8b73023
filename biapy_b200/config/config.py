"""The slice of BiaPy's YAML configuration the hot path reads (``biapy/config/config.py``; defaults copied value for value
from the lines cited).  BiaPy builds a yacs ``CfgNode`` with ~600 keys and validates it in ``check_configuration.py``; the
engine only needs the keys below, so a ``Config`` is a nested attribute dictionary: defaults, overridden by a YAML file /
string / dict.  Unknown keys are kept (a full BiaPy YAML loads unchanged) but not interpreted."""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Dict

import yaml

DEFAULTS: Dict[str, Any] = {
    "SYSTEM": {"NUM_GPUS": 1, "SEED": 0},                                                    # config.py:40-70
    "PROBLEM": {"TYPE": "SEMANTIC_SEG", "NDIM": "2D"},                                      # :83-85
    "DATA": {
        "PATCH_SIZE": (256, 256, 1), "N_CLASSES": 2,                                         # :797-799
        "NORMALIZATION": {                                                                   # :829-849
            "PERC_CLIP": {"ENABLE": False, "LOWER_PERC": -1.0, "UPPER_PERC": -1.0, "LOWER_VALUE": [-1.0], "UPPER_VALUE": [-1.0]},
            "TYPE": "zero_mean_unit_variance",
            "ZERO_MEAN_UNIT_VAR": {"MEAN_VAL": [-1.0], "STD_VAL": [-1.0]},
        },
        "TEST": {"OVERLAP": (0, 0), "PADDING": (0, 0), "MEDIAN_PADDING": False},             # :1111-1115
    },
    "MODEL": {                                                                               # :1507-1553
        "SOURCE": "biapy", "ARCHITECTURE": "unet", "FEATURE_MAPS": [16, 32, 64, 128, 256], "DROPOUT_VALUES": [0.0] * 5,
        "NORMALIZATION": "in", "KERNEL_SIZE": 3, "UPSAMPLE_LAYER": "convtranspose", "ACTIVATION": "elu", "Z_DOWN": [0, 0, 0, 0],
        "YX_DOWN": [0, 0, 0, 0], "ISOTROPY": [True] * 5, "LARGER_IO": False, "CONV_LAYERS": [2] * 5,
        "CONV_BLOCK_ORDER": "conv_norm_act", "LOAD_CHECKPOINT": False, "LOAD_CHECKPOINT_EPOCH": "best_on_val",
        "ITEMS_TO_LOAD_FROM_CHECKPOINT": ["model"], "SAVE_CKPT_FREQ": -1,
    },
    "LOSS": {"TYPE": "", "CONTRAST": {"ENABLE": False, "PROJ_DIM": 256},                     # :1916
             "CLASS_REBALANCE": "none", "CLASS_WEIGHTS": [], "IGNORE_INDEX": -1},            # :1925-1931
    "TRAIN": {"ENABLE": False, "OPTIMIZER": ["SGD"], "LR": [1.0e-4], "W_DECAY": 0.02, "OPT_BETAS": [[0.9, 0.999]],   # :1964-1990
              "BATCH_SIZE": 2, "GRADIENT_CLIP_NORM": 0.0, "EPOCHS": 360, "PATIENCE": -1, "VERBOSE": False,
              "LR_SCHEDULER": {"NAME": "", "MIN_LR": [-1.0], "REDUCEONPLATEAU_FACTOR": 0.5,                           # :2005-2031
                               "REDUCEONPLATEAU_PATIENCE": -1, "WARMUP_COSINE_DECAY_EPOCHS": -1}},
    "TEST": {"ENABLE": False, "AUGMENTATION": False, "AUGMENTATION_MODE": "mean", "AUGMENTATION_GROUP": "auto",     # :2049-2138
             "REDUCE_MEMORY": False, "BY_CHUNKS": {"ENABLE": False}, "FULL_IMG": False},
    "PATHS": {"CHECKPOINT": "checkpoints", "CHECKPOINT_FILE": ""},
}


class Config(dict):
    """Nested dict with attribute access (``cfg.DATA.TEST.PADDING``), the way BiaPy code reads its ``CfgNode``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(d):
    if isinstance(d, dict):
        return Config({k: _wrap(v) for k, v in d.items()})
    return d


def _literal(v):
    """yacs evaluates string values that look like Python literals (``"(10,10,10)"`` -> tuple); do the same."""
    if isinstance(v, str):
        s = v.strip()
        if s[:1] in "([" and s[-1:] in ")]":
            try:
                return ast.literal_eval(s)
            except (ValueError, SyntaxError):
                return v
    return v


def _merge(base: dict, over: dict):
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = {kk: _literal(vv) for kk, vv in v.items()} if isinstance(v, dict) and k not in base else _literal(v)


def load_config(source=None) -> Config:
    """`source`: path of a BiaPy YAML file, a YAML string, a (nested) dict / Config, or None for the defaults."""
    cfg = copy.deepcopy(DEFAULTS)
    if source is None:
        return _wrap(cfg)
    if isinstance(source, dict):
        over = copy.deepcopy(dict(source))
    elif isinstance(source, str) and os.path.exists(source):
        with open(source) as f:
            over = yaml.safe_load(f) or {}
    elif isinstance(source, str):
        over = yaml.safe_load(source) or {}
        if not isinstance(over, dict):
            raise FileNotFoundError(f"configuration file {source!r} not found")
    else:
        raise TypeError(f"cannot build a configuration from {type(source).__name__}")
    _merge(cfg, over)
    c = _wrap(cfg)
    # the defaults of OVERLAP / PADDING are 2D; check_configuration.py extends them for 3D problems
    nd = 3 if c.PROBLEM.NDIM == "3D" else 2
    for key in ("OVERLAP", "PADDING"):
        v = tuple(c.DATA.TEST[key])
        c.DATA.TEST[key] = v if len(v) == nd else (0,) * nd if all(x == 0 for x in v) else v
        if len(c.DATA.TEST[key]) != nd:
            raise ValueError(f"DATA.TEST.{key} must have {nd} values for a {c.PROBLEM.NDIM} problem")
    if len(tuple(c.DATA.PATCH_SIZE)) != nd + 1:
        raise ValueError(f"DATA.PATCH_SIZE must be (spatial..., channels) with {nd} spatial axes for a {c.PROBLEM.NDIM} problem")
    return c


def first(v):
    """`TRAIN.OPTIMIZER` / `TRAIN.LR` / `TRAIN.OPT_BETAS` are per-optimiser lists since BiaPy 3.6 (scalars in older YAMLs):
    ``["ADAMW"] -> "ADAMW"``, ``[[0.9, 0.999]] -> [0.9, 0.999]``, scalars and plain ``[0.9, 0.999]`` pass through."""
    if isinstance(v, (list, tuple)) and v and (len(v) == 1 or isinstance(v[0], (list, tuple))):
        return v[0]
    return v

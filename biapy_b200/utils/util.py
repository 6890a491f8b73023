"""The helper of ``biapy/utils/util.py`` the inference path uses."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch


def check_downsample_division(X, d_levels: int) -> Tuple[object, Tuple[int, ...]]:
    """Zero-pad ``(num_images, height, width, channels)`` at the bottom / right so that both extents are multiples of
    ``2 ** d_levels`` (reference ``util.py:637-674``); returns the padded data and the original shape.  numpy arrays stay
    numpy, tensors stay tensors (on their device)."""
    d_val = pow(2, d_levels)
    dy = math.ceil(X.shape[1] / d_val)
    dx = math.ceil(X.shape[2] / d_val)
    o_shape = tuple(X.shape)
    py, px = dy * d_val - X.shape[1], dx * d_val - X.shape[2]
    if py or px:
        if isinstance(X, torch.Tensor):
            X = torch.nn.functional.pad(X, (0, 0, 0, px, 0, py))
        else:
            X = np.pad(X, ((0, 0), (0, py), (0, px), (0, 0)))
        print("Data has been padded to be downsampled {} times. Its shape now is: {}".format(d_levels, tuple(X.shape)))
    return X, o_shape

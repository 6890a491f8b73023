"""Host <-> device layout helpers with the reference's signatures (``biapy/utils/misc.py:689-733``).

BiaPy keeps images as ``(N, [Z,] Y, X, C)`` numpy arrays and only *permutes* them into ``(N, C, [Z,] Y, X)``
tensors, so the tensor a model receives already has channels-last strides -- exactly the layout the B200 engine
computes in.  No data is re-ordered at the boundary."""
from __future__ import annotations

from typing import Tuple

import torch


def to_pytorch_format(x, axes_order: Tuple, device, dtype=torch.float32) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(dtype).permute(axes_order).to(device, non_blocking=True)
    return torch.from_numpy(x).to(dtype).permute(axes_order).to(device, non_blocking=True)


def to_numpy_format(x: torch.Tensor, axes_order_back: Tuple):
    return x.permute(axes_order_back).cpu().numpy()

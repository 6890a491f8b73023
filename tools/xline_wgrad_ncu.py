"""The x-line weight-gradient launch of BASELINE config[1]'s first level (16 -> 16, 128^3 x batch 4, fp16) twice, for
`ncu --set full --import-source on -k regex:conv_wgrad_xline`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biapy_b200 import _lib, ops  # noqa: E402

if __name__ == "__main__":
    dev, dtype = "cuda", torch.float16
    x = torch.randn((4, 128, 128, 128, 16), device=dev).to(dtype)
    dy = torch.randn((4, 128, 128, 128, 16), device=dev).to(dtype)
    packed = torch.zeros(16 * 27 * 16, dtype=torch.float32, device=dev)
    dbias = torch.zeros(16, dtype=torch.float32, device=dev)
    for _ in range(2):
        _lib.call("b200_conv_wgrad_xline", ops._ref(x), ops._ref(dy), ops._ptr(packed), ops._ptr(dbias), ops.stream_ptr())
    torch.cuda.synchronize()

"""Epilogue A/B micro-benchmark of the x-slab convolution kernels (run once per B200_EPI_TMA / B200_EPI_SWZ setting).

    B200_EPI_TMA=0 python tools/epi_micro.py ; B200_EPI_TMA=1 python tools/epi_micro.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biapy_b200 import _lib, ops

dt = torch.bfloat16
tag = f"VARSLOT={os.environ.get('B200_XSLAB_VARSLOT', '1')}"
SHAPES = [(16, 16, 128, 4), (2, 16, 128, 4), (48, 16, 128, 4), (16, 48, 128, 4), (32, 32, 64, 4), (96, 32, 64, 4), (16, 64, 64, 4)]
if os.environ.get('EPI_MICRO_SHAPES') == 'wres':
    SHAPES = SHAPES[:2]
for (cin, cout, size, batch) in SHAPES:
    x = torch.randn(batch, size, size, size, cin, device="cuda").to(dt)
    w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    y = torch.zeros(batch, size, size, size, cout, device="cuda", dtype=dt)
    sums = torch.zeros(batch * cout * 2, dtype=torch.float64, device="cuda")
    wp = ops.pack_conv_weight_xfold(w, dt, False)
    flops = 2.0 * batch * size ** 3 * cin * cout * 27
    variants = {
        "plain": lambda: ops.conv_fprop(x, wp, b, y, (3, 3, 3), impl=_lib.IMPL_XFOLD),
        "accumulate": lambda: ops.conv_fprop(x, wp, b, y, (3, 3, 3), accumulate=True, impl=_lib.IMPL_XFOLD),
    }
    if cout == 16:
        variants["stats"] = lambda: ops.conv_fprop_stats(x, wp, b, y, (3, 3, 3), sums)
    for name, fn in variants.items():
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"[{tag}] {cin}->{cout} @{size}^3 x{batch} {name}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    y.zero_()

"""Micro-benchmark of one convolution layer through the C ABI (used for the ncu captures in profiles/).

    python profiles/conv_micro.py --cin 16 --cout 16 --size 128 --batch 4 --op fprop --iters 5
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biapy_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=16)
ap.add_argument("--cout", type=int, default=16)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--op", default="fprop", choices=["fprop", "wgrad"])
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--impl", default="umma")
a = ap.parse_args()
impl = {"umma": _lib.IMPL_UMMA, "simt": _lib.IMPL_SIMT, "xfold": _lib.IMPL_XFOLD}[a.impl]
dt = torch.bfloat16
x = torch.randn(a.batch, a.size, a.size, a.size, a.cin, device="cuda").to(dt)
w = torch.randn(a.cout, a.cin, a.k, a.k, a.k, device="cuda") * 0.05
b = torch.zeros(a.cout, device="cuda")
y = torch.empty(a.batch, a.size, a.size, a.size, a.cout, device="cuda", dtype=dt)
k = (a.k,) * 3
wp = ops.pack_conv_weight_xfold(w, dt, False) if a.impl == "xfold" else ops.pack_conv_weight(w, dt, False)
flops = 2.0 * a.batch * a.size ** 3 * a.cin * a.cout * a.k ** 3


def run():
    if a.op == "fprop":
        ops.conv_fprop(x, wp, b, y, k, impl=impl)
    else:
        dw = torch.empty_like(w)
        ops.conv_wgrad(x, y, a.cout, a.cin, k, dw, b, impl=impl)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print(f"{a.op} {a.cin}->{a.cout} k{a.k} @{a.size}^3 x{a.batch} impl={a.impl}: "
      f"{ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")

"""Micro-benchmark of the HBM-bound normalisation kernels through the C ABI: GB/s against the measured copy peak.

    python profiles/rows_micro.py [--c 48] [--size 128] [--batch 4]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from biapy_b200 import ops
from biapy_b200.ops import _launch, _ref, _ptr, stream_ptr
from biapy_b200._lib import ACT

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=48)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--act", default="silu")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dt = torch.bfloat16
n, s, c = a.batch, a.size, a.c
groups = 8
x = torch.randn(n, s, s, s, c, device="cuda").to(dt)
dy = torch.randn(n, s, s, s, c, device="cuda").to(dt)
y = torch.empty_like(x)
dx = torch.empty_like(x)
gamma = torch.ones(c, device="cuda")
beta = torch.zeros(c, device="cuda")
nbytes = x.numel() * 2
st = ops.norm_stats(x, groups, gamma, beta)
red = torch.zeros(n * c * 2, dtype=torch.float64, device="cuda")
coef = torch.randn(n * c * 4, device="cuda") * 0.01
sums = torch.zeros(n * c * 2, dtype=torch.float64, device="cuda")

cases = {
    "channel_sums (1 read)": (1, lambda: _launch("b200_channel_sums", _ref(x), _ptr(sums), stream_ptr())),
    "scale_shift_act (1 read + 1 write)": (2, lambda: ops.scale_shift_act(x, st.scale, st.shift, a.act, y)),
    "norm_act_bwd_reduce (2 reads)": (2, lambda: _launch("b200_norm_act_bwd_reduce", _ref(x), _ref(dy), _ptr(st.mean), _ptr(st.rstd),
                                                        st.groups, _ptr(gamma), _ptr(beta), ACT[a.act], _ptr(red), stream_ptr())),
    "norm_act_bwd_apply (2 reads + 1 write)": (3, lambda: _launch("b200_norm_act_bwd_apply", _ref(x), _ref(dy), ACT[a.act], _ptr(coef),
                                                                  _ref(dx), 0, stream_ptr())),
    "torch copy (1 read + 1 write)": (2, lambda: y.copy_(x)),
}
for name, (passes, fn) in cases.items():
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"c{c} @{s}^3 x{n} {name}: {ms:.3f} ms  {passes * nbytes / ms / 1e6:.0f} GB/s  variant={os.environ.get('B200_ROWS_VARIANT', '0')}")

"""Device probe of the x-line convolution (csrc/conv_xline.cu): the tensor-memory operand self-test, parity against ATen fp32 on the
rounded operands for plain / fused / accumulating launches, and conv-only timings next to the x-folded kernels.

    python tools/xline_probe.py [--quick] [--time]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biapy_b200 import _lib, ops  # noqa: E402


def ref_conv(x, w, b):
    """x: (N, D, H, W, C) any dtype -> fp32 conv3d 'same' on the values as stored."""
    y = torch.nn.functional.conv3d(x.float().permute(0, 4, 1, 2, 3), w, b, padding=1)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def nmax(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def case(n, d, h, cin, dtype, fuse=0, accumulate=False, stats=False, a_out=False, flip=False, seed=0, ld_y=None):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    x = torch.randn((n, d, h, 128, cin), device=dev, generator=g).to(dtype)
    cout = 16
    if flip:   # dgrad operand: the parameter is (Cout_orig = cin, Cin_orig = 16, 3, 3, 3) and the launch computes conv(x, W')
        w = torch.randn((cin, 16, 3, 3, 3), device=dev, generator=g) * 0.1
        w_eff = w.permute(1, 0, 2, 3, 4).flip(2, 3, 4).contiguous()
    else:
        w = torch.randn((16, cin, 3, 3, 3), device=dev, generator=g) * 0.1
        w_eff = w
    b = torch.randn(16, device=dev, generator=g)
    wp = ops.pack_conv_weight_xline(w, dtype, flip)
    w_r = w_eff.to(dtype).float()
    scale = shift = None
    xa = x
    if fuse:
        scale = (torch.rand((n, cin), device=dev, generator=g) + 0.5).contiguous()
        shift = (torch.randn((n, cin), device=dev, generator=g) * 0.3).contiguous()
        xa = torch.empty_like(x)
        if fuse == 2:
            _lib.call("b200_scale_shift_silu_fast", ops._ref(x), ops._ptr(scale), ops._ptr(shift), ops._ref(xa), ops.stream_ptr())
        else:
            _lib.call("b200_scale_shift_act", ops._ref(x), ops._ptr(scale), ops._ptr(shift), _lib.ACT["silu"], ops._ref(xa),
                      ops.stream_ptr())
    if ld_y:
        ybuf = torch.zeros((n, d, h, 128, ld_y), device=dev, dtype=dtype)
        y = ybuf[..., 8:8 + 16] if ld_y >= 32 else ybuf[..., :16]
    else:
        y = torch.empty((n, d, h, 128, cout), device=dev, dtype=dtype)
    old = None
    if accumulate:
        old = torch.randn(y.shape, device=dev, generator=g).to(dtype)
        y.copy_(old)
    ao = torch.empty_like(x) if a_out else None
    sums = torch.zeros(n * 16 * 2, dtype=torch.float64, device=dev) if stats else None
    ops.conv_fprop_xline(x, wp, b, y, accumulate=accumulate, scale=scale, shift=shift, fuse=fuse, a_out=ao, sums=sums)
    torch.cuda.synchronize()
    want = ref_conv(xa, w_r, b)
    if accumulate:
        want = want + old.float()
    e = nmax(y, want)
    msg = f"n{n} d{d} h{h} cin{cin} {str(dtype)[6:]} fuse{fuse} acc{int(accumulate)} flip{int(flip)}: nmax {e:.2e}"
    ok = e < (2e-2 if dtype == torch.bfloat16 else 3e-3)
    if a_out:
        same = torch.equal(ao, xa)
        msg += f" a_out==unfused {same}"
        ok = ok and same
    if stats:
        yf = y.float()
        s = torch.stack([yf.sum((1, 2, 3)), (yf * yf).sum((1, 2, 3))], -1).double().reshape(-1)
        es = ((sums - s).abs().max() / s.abs().max()).item()
        msg += f" stats rel {es:.1e}"
        ok = ok and es < 1e-4
    print(("ok   " if ok else "FAIL ") + msg, flush=True)
    return ok


def time_ms(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timings(dtype):
    dev = "cuda"
    for cin in (16, 48):
        x = torch.randn((4, 128, 128, 128, cin), device=dev).to(dtype)
        w = torch.randn((16, cin, 3, 3, 3), device=dev) * 0.1
        b = torch.randn(16, device=dev)
        y = torch.empty((4, 128, 128, 128, 16), device=dev, dtype=dtype)
        xa = torch.empty_like(x)
        scale = torch.rand((4, cin), device=dev) + 0.5
        shift = torch.randn((4, cin), device=dev)
        wl = ops.pack_conv_weight_xline(w, dtype, False)
        wf = ops.pack_conv_weight_xfold(w, dtype, False)
        gf = 2.0 * 4 * 128 ** 3 * cin * 16 * 27 / 1e9
        sums = torch.zeros(4 * 16 * 2, dtype=torch.float64, device=dev)
        fz = 2 if dtype == torch.bfloat16 else 1
        rows = [
            ("xfold", lambda: ops.conv_fprop(x, wf, b, y, (3, 3, 3), impl=_lib.IMPL_XFOLD)),
            ("xfold +acc", lambda: ops.conv_fprop(x, wf, b, y, (3, 3, 3), accumulate=True, impl=_lib.IMPL_XFOLD)),
            ("xline", lambda: ops.conv_fprop_xline(x, wl, b, y)),
            ("xline +acc", lambda: ops.conv_fprop_xline(x, wl, b, y, accumulate=True)),
            ("xline +stats", lambda: ops.conv_fprop_xline(x, wl, b, y, sums=sums)),
            (f"xline fuse{fz}", lambda: ops.conv_fprop_xline(x, wl, b, y, scale=scale, shift=shift, fuse=fz)),
            (f"xline fuse{fz} +a_out", lambda: ops.conv_fprop_xline(x, wl, b, y, scale=scale, shift=shift, fuse=fz, a_out=xa)),
            ("xline fuse1", lambda: ops.conv_fprop_xline(x, wl, b, y, scale=scale, shift=shift, fuse=1)),
            ("xline fuse2", lambda: ops.conv_fprop_xline(x, wl, b, y, scale=scale, shift=shift, fuse=2)),
            ("scale_shift_act alone", lambda: ops.scale_shift_act(x, scale, shift, "silu", xa)),
        ]
        for name, fn in rows:
            ms = time_ms(fn)
            tf = gf / ms if "alone" not in name else 0.0
            print(f"time {str(dtype)[6:]} {cin}->16 @128^3 x4 {name}: {ms:.4f} ms  {tf:.0f} TFLOP/s", flush=True)


def timings48(dtype):
    """Cin = 16 -> Cout = 48 (the input gradient of the 48 -> 16 layer): x-folded against x-line."""
    dev = "cuda"
    x = torch.randn((4, 128, 128, 128, 16), device=dev).to(dtype)
    w = torch.randn((48, 16, 3, 3, 3), device=dev) * 0.1
    y = torch.empty((4, 128, 128, 128, 48), device=dev, dtype=dtype)
    wl = ops.pack_conv_weight_xline(w, dtype, False)
    wf = ops.pack_conv_weight_xfold(w, dtype, False)
    gf = 2.0 * 4 * 128 ** 3 * 16 * 48 * 27 / 1e9
    for name, fn in (("xfold", lambda: ops.conv_fprop(x, wf, None, y, (3, 3, 3), impl=_lib.IMPL_XFOLD)),
                     ("xline", lambda: ops.conv_fprop_xline(x, wl, None, y)),
                     ("xfold", lambda: ops.conv_fprop(x, wf, None, y, (3, 3, 3), impl=_lib.IMPL_XFOLD)),
                     ("xline", lambda: ops.conv_fprop_xline(x, wl, None, y))):
        ms = time_ms(fn)
        print(f"time {str(dtype)[6:]} 16->48 @128^3 x4 {name}: {ms:.4f} ms  {gf / ms:.0f} TFLOP/s", flush=True)


def timings_wgrad(dtype):
    """Weight gradient of the 16-output-channel layers at 128^3 x 4: x-folded kernels against the x-line form."""
    dev = "cuda"
    for cin in (16, 48):
        x = torch.randn((4, 128, 128, 128, cin), device=dev).to(dtype)
        dy = torch.randn((4, 128, 128, 128, 16), device=dev).to(dtype)
        packed = torch.zeros(16 * 27 * cin, dtype=torch.float32, device=dev)
        dbias = torch.zeros(16, dtype=torch.float32, device=dev)
        gf = 2.0 * 4 * 128 ** 3 * cin * 16 * 27 / 1e9
        rows = [("xfold", lambda: _lib.call("b200_conv_wgrad", ops._ref(x), ops._ref(dy), ops._ptr(packed), ops._ptr(dbias), 3, 3, 3,
                                            _lib.IMPL_AUTO, ops.stream_ptr())),
                ("xline", lambda: _lib.call("b200_conv_wgrad_xline", ops._ref(x), ops._ref(dy), ops._ptr(packed), ops._ptr(dbias),
                                            ops.stream_ptr()))]
        for name, fn in rows + rows:
            ms = time_ms(fn, reps=10)
            print(f"time {str(dtype)[6:]} wgrad {cin}->16 @128^3 x4 {name}: {ms:.4f} ms  {gf / ms:.0f} TFLOP/s", flush=True)


def diagnose(dtype, sweep=True):
    """Per-role cycle counters (B200_XL_DBG) and ablation timings (B200_XL_ABLATE) of the full-size launches."""
    dev = "cuda"
    for cin in (16, 48):
        x = torch.randn((4, 128, 128, 128, cin), device=dev).to(dtype)
        w = torch.randn((16, cin, 3, 3, 3), device=dev) * 0.1
        b = torch.randn(16, device=dev)
        y = torch.zeros((4, 128, 128, 128, 16), device=dev, dtype=dtype)
        scale = torch.rand((4, cin), device=dev) + 0.5
        shift = torch.randn((4, cin), device=dev)
        wl = ops.pack_conv_weight_xline(w, dtype, False)
        fz = 2 if dtype == torch.bfloat16 else 1
        variants = [("plain", dict()), ("acc", dict(accumulate=True)), (f"fuse{fz}", dict(scale=scale, shift=shift, fuse=fz))]
        os.environ["B200_XL_DBG"] = "1"
        for name, kw in variants:
            print(f"dbg {str(dtype)[6:]} {cin}->16 {name}:", flush=True)
            ops.conv_fprop_xline(x, wl, b, y, **kw)
            torch.cuda.synchronize()
        os.environ["B200_XL_DBG"] = "0"
        for ab in ((0, 1, 4, 8, 5, 9, 12, 13, 2) if sweep else (0, 512, 1024 + 512, 125, 125 + 512, 125 + 512 + 1024)):
            os.environ["B200_XL_ABLATE"] = str(ab)
            for name, kw in variants:
                if ab == 2 and not name.startswith("fuse"):
                    continue
                ms = time_ms(lambda: ops.conv_fprop_xline(x, wl, b, y, **kw), reps=10)
                print(f"ablate {ab:2d} {str(dtype)[6:]} {cin}->16 {name}: {ms:.4f} ms", flush=True)
        os.environ["B200_XL_ABLATE"] = "0"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--diag", action="store_true")
    ap.add_argument("--wgrad", action="store_true")
    ap.add_argument("--diag-quick", action="store_true")
    a = ap.parse_args()
    err = ops.xline_selftest(verbose=3 if (a.diag or a.diag_quick) else 2)
    print(f"selftest max abs error {err}", flush=True)
    ok = err == 0.0
    bf, hf = torch.bfloat16, torch.float16
    ok &= case(1, 3, 8, 16, hf)
    ok &= case(1, 5, 16, 16, hf, seed=1)
    ok &= case(2, 9, 20, 16, bf, seed=2)
    ok &= case(1, 4, 8, 48, hf, seed=3)
    if not a.quick:
        ok &= case(2, 7, 10, 48, bf, seed=4)
        ok &= case(1, 6, 16, 16, hf, fuse=1, a_out=True, seed=5)
        ok &= case(1, 6, 16, 16, bf, fuse=2, a_out=True, stats=True, seed=6)
        ok &= case(1, 6, 12, 48, hf, fuse=1, a_out=True, seed=7)
        ok &= case(1, 6, 12, 48, bf, fuse=2, seed=8)
        ok &= case(2, 5, 16, 16, hf, accumulate=True, stats=True, seed=9)
        ok &= case(1, 5, 16, 16, hf, flip=True, seed=10)
        ok &= case(1, 5, 8, 48, hf, flip=True, seed=11)
        ok &= case(1, 5, 16, 16, hf, ld_y=48, seed=12)
        ok &= case(2, 40, 128, 16, hf, stats=True, seed=13)
        ok &= case(1, 128, 128, 16, bf, fuse=2, seed=14)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if a.wgrad:
        timings_wgrad(torch.float16)
        timings_wgrad(torch.bfloat16)
    if a.time:
        timings48(torch.float16)
        timings48(torch.bfloat16)
        timings(torch.float16)
        timings(torch.bfloat16)
    if a.diag or a.diag_quick:
        diagnose(torch.float16, sweep=not a.diag_quick)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())

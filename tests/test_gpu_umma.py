"""GPU parity of the tcgen05/TMA convolution kernels (forced with impl=UMMA through the C ABI) against the CPU
oracle arithmetic (ATen fp32 on bf16-rounded operands).  Covers every shared-memory swizzle mode the kernel uses
(32/64/128 B <-> Cin chunk 16/32/64), partial tiles, N tiles up to 256, channel-slice outputs, the accumulate
epilogue, dgrad (flipped packing) and fp16."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def cl(x):
    return x.permute(0, 2, 3, 4, 1).contiguous().cuda()


def ncdhw(x):
    return x.float().permute(0, 4, 1, 2, 3).cpu()


def nerr(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


CASES = [
    # n, d, h, w, cin, cout, k
    (1, 8, 8, 8, 16, 16, (3, 3, 3)),       # CK=16 / SWIZZLE_32B
    (2, 8, 16, 16, 32, 32, (3, 3, 3)),     # CK=32 / SWIZZLE_64B
    (1, 8, 8, 16, 64, 64, (3, 3, 3)),      # CK=64 / SWIZZLE_128B
    (1, 4, 8, 8, 48, 16, (3, 3, 3)),       # 3 chunks of 16 (decoder concat)
    (1, 8, 8, 8, 128, 256, (3, 3, 3)),     # N tile 256, 2 chunks of 64
    (1, 8, 8, 8, 16, 48, (3, 3, 3)),       # N = 48
    (1, 5, 9, 11, 16, 16, (3, 3, 3)),      # partial tiles in every axis
    (2, 1, 24, 40, 32, 16, (1, 3, 3)),     # 2D
    (1, 4, 8, 8, 64, 32, (1, 1, 1)),       # pointwise
    (3, 16, 16, 16, 16, 32, (3, 3, 3)),    # several tiles per CTA (persistent loop + TMEM double buffering)
    (1, 32, 32, 32, 96, 32, (3, 3, 3)),    # more tiles than SMs
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_fprop_umma(case, dtype):
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float()
    wt = (torch.randn(cout, cin, *k, generator=g) * 0.1).to(dtype).float()
    b = torch.randn(cout, generator=g)
    yr = F.conv3d(x, wt, b, padding=[kk // 2 for kk in k])
    xd = cl(x).to(dtype)
    wp = ops.pack_conv_weight(wt.cuda(), dtype, False)
    ybuf = torch.zeros(n, d, h, w, cout + 16, dtype=dtype, device="cuda")
    yv = ybuf[..., 8:8 + cout]
    ops.conv_fprop(xd, wp, b.cuda(), yv, k, impl=_lib.IMPL_UMMA)
    torch.cuda.synchronize()
    assert nerr(ncdhw(yv), yr) < 1.5e-2
    assert ybuf[..., :8].abs().max().item() == 0 and ybuf[..., 8 + cout:].abs().max().item() == 0
    # agreement with the CUDA-core kernel on identical operands is much tighter (both accumulate in fp32)
    y2 = torch.empty(n, d, h, w, cout, dtype=dtype, device="cuda")
    ops.conv_fprop(xd, wp, b.cuda(), y2, k, impl=_lib.IMPL_SIMT)
    assert nerr(yv.float().cpu(), y2.float().cpu()) < 8e-3
    before = ncdhw(yv)
    ops.conv_fprop(xd, wp, b.cuda(), yv, k, accumulate=True, impl=_lib.IMPL_UMMA)
    assert nerr(ncdhw(yv), before + yr) < 3e-2


@pytest.mark.parametrize("case", CASES[:5])
def test_conv_dgrad_umma(case):
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, d, h, w, generator=g, requires_grad=True)
    wt = (torch.randn(cout, cin, *k, generator=g) * 0.1).to(dtype).float()
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    F.conv3d(x, wt, None, padding=[kk // 2 for kk in k]).backward(gy)
    wpf = ops.pack_conv_weight(wt.cuda(), dtype, True)
    dx = torch.empty(n, d, h, w, cin, dtype=dtype, device="cuda")
    ops.conv_fprop(cl(gy).to(dtype), wpf, None, dx, k, impl=_lib.IMPL_UMMA)
    assert nerr(ncdhw(dx), x.grad) < 1.5e-2


WGRAD_CASES = [
    (1, 8, 8, 8, 16, 16, (3, 3, 3)),       # B: SWIZZLE_32B, one M-group of 4 blocks (27 chunks -> 5 dummy slots)
    (2, 8, 16, 16, 32, 32, (3, 3, 3)),     # B: SWIZZLE_64B
    (1, 8, 8, 16, 64, 64, (3, 3, 3)),      # B: SWIZZLE_128B, several M-groups
    (1, 4, 8, 8, 48, 16, (3, 3, 3)),       # decoder concat shape
    (1, 8, 8, 8, 128, 256, (3, 3, 3)),     # N = 256 (4 boxes), g = 2
    (1, 8, 8, 8, 32, 128, (3, 3, 3)),      # N = 128 (2 boxes)
    (1, 5, 9, 11, 16, 16, (3, 3, 3)),      # partial tiles
    (2, 1, 24, 40, 32, 16, (1, 3, 3)),     # 2D
    (1, 4, 8, 8, 64, 32, (1, 1, 1)),       # pointwise
    (3, 16, 16, 16, 16, 32, (3, 3, 3)),    # many voxel tiles per CTA
    (1, 32, 32, 32, 96, 32, (3, 3, 3)),
    (1, 8, 16, 16, 48, 16, (1, 1, 1)),     # x-folded wgrad, pointwise
    (2, 9, 20, 12, 16, 64, (3, 3, 3)),     # x-folded wgrad, N = 256, partial tiles
    (1, 16, 16, 8, 48, 16, (3, 3, 3)),     # x-folded wgrad, 3 M-groups
    (1, 9, 20, 12, 16, 16, (3, 3, 3)),     # z-slab wgrad: partial tiles in z and y, 9 slab atoms -> 3 groups
    (2, 24, 32, 16, 16, 16, (3, 3, 3)),    # z-slab wgrad: several voxel tiles per CTA (ring wrap-around)
    (1, 16, 16, 16, 96, 16, (3, 3, 3)),    # z-slab wgrad: 54 slab atoms
]
# Cout outside the kernels' N tiles: ops.conv_wgrad covers it with slices of dy (512 = 256 + 256, 384 = 256 + 128, 48 = 32 + 16)
WGRAD_SPLIT_CASES = [(2, 1, 16, 16, 256, 512, (1, 3, 3)), (1, 4, 8, 8, 64, 384, (3, 3, 3)), (1, 8, 8, 8, 16, 48, (3, 3, 3))]


@pytest.mark.parametrize("case", WGRAD_SPLIT_CASES)
def test_conv_wgrad_cout_slices(case):
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float()
    wt = torch.zeros(cout, cin, *k, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    F.conv3d(x, wt, b, padding=[kk // 2 for kk in k]).backward(gy)
    assert len(ops._wgrad_cuts(cout)) > 1
    dw = torch.empty(cout, cin, *k, device="cuda")
    db = torch.zeros(cout, device="cuda")
    n0 = ops.LAUNCHES
    ops.conv_wgrad(cl(x).to(dtype), cl(gy).to(dtype), cout, cin, k, dw, db)          # AUTO: must not fall back to the CUDA cores
    assert ops.LAUNCHES - n0 == len(ops._wgrad_cuts(cout)) + 1
    torch.cuda.synchronize()
    assert nerr(dw.cpu(), wt.grad) < 2e-3 and nerr(db.cpu(), b.grad) < 2e-3


@pytest.mark.parametrize("case", WGRAD_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_wgrad_umma(case, dtype):
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float()
    wt = torch.zeros(cout, cin, *k, requires_grad=True)
    b = torch.zeros(cout, requires_grad=True)
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    F.conv3d(x, wt, b, padding=[kk // 2 for kk in k]).backward(gy)
    dw = torch.empty(cout, cin, *k, device="cuda")
    db = torch.zeros(cout, device="cuda")
    ops.conv_wgrad(cl(x).to(dtype), cl(gy).to(dtype), cout, cin, k, dw, db, impl=_lib.IMPL_UMMA)
    torch.cuda.synchronize()
    assert nerr(dw.cpu(), wt.grad) < 2e-3          # same bf16 operands, fp32 accumulation on both sides
    assert nerr(db.cpu(), b.grad) < 2e-3


@pytest.mark.parametrize("shape", [(2, 4, 8, 8, 32, 32), (1, 8, 8, 8, 256, 256), (1, 4, 4, 16, 64, 64), (1, 3, 5, 12, 32, 16),
                                   (2, 1, 16, 16, 512, 256)])       # Cin = 512: the bottleneck up-sampling of BASELINE config[4]
@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2)])
def test_convT_tensor_core_route(shape, stride):
    """ConvTranspose(k = s): fprop / dgrad / wgrad through the tcgen05 kernels on strided sub-lattice views."""
    from biapy_b200 import ops
    n, d, h, w, cin, cout = shape
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float().requires_grad_(True)
    wt = (torch.randn(cin, cout, *stride, generator=g) * 0.2).to(dtype).float().requires_grad_(True)
    b = torch.randn(cout, generator=g).requires_grad_(True)
    gy = torch.randn(n, cout, d * stride[0], h * stride[1], w * stride[2], generator=g).to(dtype).float()
    yr = F.conv_transpose3d(x, wt, b, stride=stride)
    yr.backward(gy)
    xd, gyd = cl(x.detach()).to(dtype), cl(gy).to(dtype)
    ybuf = torch.zeros(n, d * stride[0], h * stride[1], w * stride[2], cout + 16, dtype=dtype, device="cuda")
    y = ybuf[..., :cout]                       # first slice of a concat buffer, as the decoder uses it
    assert ops.convT_tc_supported(xd, y, stride)
    wp = ops.pack_convT_weight(wt.detach().cuda(), dtype, False)
    ops.convT_fprop_tc(xd, wp, b.detach().cuda(), y, stride)
    torch.cuda.synchronize()
    assert nerr(ncdhw(y), yr.detach()) < 1.5e-2
    assert ybuf[..., cout:].abs().max().item() == 0
    wpt = ops.pack_convT_weight(wt.detach().cuda(), dtype, True)
    dx = torch.empty_like(xd)
    ops.convT_dgrad_tc(gyd, wpt, dx, stride)
    assert nerr(ncdhw(dx), x.grad) < 1.5e-2
    dw, db = torch.zeros_like(wt).cuda(), torch.zeros(cout, device="cuda")
    ops.convT_wgrad_tc(xd, gyd, dw, db, stride)
    assert nerr(dw.cpu(), wt.grad) < 2e-3 and nerr(db.cpu(), b.grad) < 2e-3


XFOLD_CASES = [
    # n, d, h, w, cin, cout, k
    (1, 8, 16, 16, 16, 16, (3, 3, 3)),     # K row = 96 el: one 64-box + one 32-box, N = 64
    (2, 16, 16, 32, 32, 32, (3, 3, 3)),    # 192 el: three 64-boxes, N = 128
    (1, 8, 16, 8, 48, 16, (3, 3, 3)),      # decoder: 288 el = 4 x 64 + 32
    (1, 8, 16, 16, 16, 48, (3, 3, 3)),     # dgrad of the decoder conv: N = 192
    (1, 16, 8, 12, 96, 32, (3, 3, 3)),     # 576 el = 9 boxes
    (1, 9, 20, 12, 16, 16, (3, 3, 3)),     # partial tiles in y and z
    (3, 1, 128, 16, 32, 16, (1, 3, 3)),    # 2D (kd = 1)
    (1, 32, 32, 32, 16, 64, (3, 3, 3)),    # N = 256, more tiles than SMs
    (1, 24, 24, 8, 48, 16, (3, 3, 3)),     # z-slab variant: two row tiles per CTA, partial tile in z and y
    (1, 20, 16, 8, 16, 48, (3, 3, 3)),     # z-slab, N = 192: single TMEM buffer
    (2, 40, 64, 32, 16, 16, (3, 3, 3)),    # z-slab, several tiles per CTA (ring wrap-around across tiles)
    (1, 8, 16, 16, 48, 16, (1, 1, 1)),     # pointwise shortcut: window of 4 voxels, block-diagonal weights
    (2, 16, 8, 8, 16, 32, (1, 1, 1)),
    (1, 8, 16, 32, 96, 32, (1, 1, 1)),
]


@pytest.mark.parametrize("case", XFOLD_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_fprop_xfold(case, dtype):
    """x-folded kernel (4 x-voxels per GEMM row, block-Toeplitz weights) vs the CPU oracle arithmetic, fprop + dgrad."""
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(4)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float().requires_grad_(True)
    wt = (torch.randn(cout, cin, *k, generator=g) * 0.1).to(dtype).float()
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    yr = F.conv3d(x, wt, b, padding=[kk // 2 for kk in k])
    yr.backward(gy)
    xd = cl(x.detach()).to(dtype)
    ybuf = torch.zeros(n, d, h, w, cout + 16, dtype=dtype, device="cuda")
    yv = ybuf[..., 8:8 + cout]
    assert ops.conv_impl_query(xd, yv, k) == _lib.IMPL_XFOLD
    wp = ops.pack_conv_weight_xfold(wt.cuda(), dtype, False)
    ops.conv_fprop(xd, wp, b.cuda(), yv, k, impl=_lib.IMPL_XFOLD)
    torch.cuda.synchronize()
    assert nerr(ncdhw(yv), yr.detach()) < 1.5e-2
    assert ybuf[..., :8].abs().max().item() == 0 and ybuf[..., 8 + cout:].abs().max().item() == 0
    before = ncdhw(yv)
    ops.conv_fprop(xd, wp, b.cuda(), yv, k, accumulate=True, impl=_lib.IMPL_XFOLD)
    assert nerr(ncdhw(yv), before + yr.detach()) < 3e-2
    # dgrad through the same kernel with the flipped / transposed Toeplitz packing (when the roles fit: Cout <= 96, Cin <= 64)
    gyd = cl(gy).to(dtype)
    dx = torch.empty(n, d, h, w, cin, dtype=dtype, device="cuda")
    if ops.conv_impl_query(gyd, dx, k) == _lib.IMPL_XFOLD:
        wpf = ops.pack_conv_weight_xfold(wt.cuda(), dtype, True)
        ops.conv_fprop(gyd, wpf, None, dx, k, impl=_lib.IMPL_XFOLD)
        assert nerr(ncdhw(dx), x.grad) < 1.5e-2


IMAGE_CASES = [
    # n, d, h, w, cin, cout, k        image-fed layers: narrow x-folded window (3x3x3) / CUDA-core stream (1x1x1)
    (1, 16, 16, 16, 2, 16, (3, 3, 3)),
    (2, 9, 20, 12, 2, 16, (3, 3, 3)),      # partial tiles in z and y
    (1, 8, 32, 16, 4, 16, (3, 3, 3)),
    (1, 24, 16, 8, 8, 16, (3, 3, 3)),
    (1, 8, 16, 16, 2, 16, (1, 1, 1)),
    (2, 4, 8, 8, 1, 16, (1, 1, 1)),
]


@pytest.mark.parametrize("case", IMAGE_CASES)
def test_conv_image_fed(case):
    """Cin = 1..8 layers without channel padding: fprop (+ accumulate) and wgrad through the automatic dispatch."""
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    dtype = torch.bfloat16
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float()
    wt = (torch.randn(cout, cin, *k, generator=g) * 0.2).to(dtype).float().requires_grad_(True)
    b = torch.randn(cout, generator=g).requires_grad_(True)
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    yr = F.conv3d(x, wt, b, padding=[kk // 2 for kk in k])
    yr.backward(gy)
    xd = cl(x).to(dtype)
    # output and its gradient live in a channel slice of a wider buffer, as the first encoder block's do
    ybuf = torch.zeros(n, d, h, w, cout + 32, dtype=dtype, device="cuda")
    yv = ybuf[..., 16:16 + cout]
    impl = ops.conv_impl_query(xd, yv, k)
    if k == (3, 3, 3) and cin in (2, 4, 8):
        assert impl == _lib.IMPL_XFOLD
        wp = ops.pack_conv_weight_xfold(wt.detach().cuda(), dtype, False)
    else:
        impl = _lib.IMPL_AUTO
        wp = ops.pack_conv_weight(wt.detach().cuda(), dtype, False)
    ops.conv_fprop(xd, wp, b.detach().cuda(), yv, k, impl=impl)
    torch.cuda.synchronize()
    assert nerr(ncdhw(yv), yr.detach()) < 1.5e-2
    assert ybuf[..., :16].abs().max().item() == 0 and ybuf[..., 16 + cout:].abs().max().item() == 0
    before = ncdhw(yv)
    ops.conv_fprop(xd, wp, b.detach().cuda(), yv, k, accumulate=True, impl=impl)
    assert nerr(ncdhw(yv), before + yr.detach()) < 3e-2
    gyd = cl(gy).to(dtype)
    dw = torch.empty(cout, cin, *k, device="cuda")
    db = torch.zeros(cout, device="cuda")
    ops.conv_wgrad(xd, gyd, cout, cin, k, dw, db)
    torch.cuda.synchronize()
    assert nerr(dw.cpu(), wt.grad) < 2e-3
    assert nerr(db.cpu(), b.grad) < 2e-3


STATS_CASES = [
    # n, d, h, w, cin, cout      x-slab shapes with fused channel statistics (Cout = 16)
    (2, 16, 16, 16, 16, 16),
    (1, 20, 24, 12, 48, 16),       # partial tiles in z and y
    (2, 16, 32, 8, 32, 16),
    (1, 9, 20, 12, 16, 16),        # thin volume: one row tile per CTA tile
    (3, 40, 32, 16, 2, 16),        # image-fed layer, several tiles and several samples per CTA
]


@pytest.mark.parametrize("case", STATS_CASES)
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_conv_epilogue_statistics(case, dtype):
    """b200_conv_fprop_stats: the channel sums that leave the convolution epilogue equal b200_channel_sums run on the stored
    output, and the output itself is bit-identical to the plain launch."""
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout = case
    k = (3, 3, 3)
    g = torch.Generator().manual_seed(11)
    x = cl(torch.randn(n, cin, d, h, w, generator=g)).to(dtype)
    wt = (torch.randn(cout, cin, *k, generator=g) * 0.1).cuda()
    b = torch.randn(cout, generator=g).cuda()
    wp = ops.pack_conv_weight_xfold(wt, dtype, False)
    y_ref = torch.empty(n, d, h, w, cout, dtype=dtype, device="cuda")
    assert ops.conv_impl_query(x, y_ref, k) == _lib.IMPL_XFOLD
    ops.conv_fprop(x, wp, b, y_ref, k, impl=_lib.IMPL_XFOLD)
    y = torch.empty_like(y_ref)
    sums = torch.zeros(n * cout * 2, dtype=torch.float64, device="cuda")
    assert ops.conv_fprop_stats(x, wp, b, y, k, sums)
    torch.cuda.synchronize()
    assert torch.equal(y, y_ref)
    ref = torch.zeros_like(sums)
    _lib.call("b200_channel_sums", ops._ref(y_ref), ops._ptr(ref), _lib.stream_ptr())
    yf = y_ref.double()
    exact = torch.stack([yf.sum((1, 2, 3)), (yf * yf).sum((1, 2, 3))], -1).reshape(-1)
    scale = exact.abs().max().item()
    assert (ref - exact).abs().max().item() < 1e-4 * scale
    assert (sums - exact).abs().max().item() < 1e-4 * scale         # fp32 partials per thread, fp64 across threads
    # an accumulating launch cannot see the final value in the TMA-store epilogue: it reports "not applied" and leaves sums alone
    before = sums.clone()
    applied = ops.conv_fprop_stats(x, wp, b, y, k, sums, accumulate=True)
    torch.cuda.synchronize()
    if not applied:
        assert torch.equal(sums, before)


@pytest.mark.parametrize("swz", ["32", "0"])
def test_xslab_tma_store_epilogue_matches_direct_stores(swz):
    """The TMA-store epilogue (staged boxes, bulk tensor store / element-wise add) against the direct-store epilogue of the same
    kernel, run in a subprocess per staging-buffer swizzle because the mode is latched from the environment."""
    import os
    import subprocess
    import sys
    code = r'''
import torch
from biapy_b200 import _lib, ops
torch.manual_seed(0)
for (n, d, h, w, cin, cout) in [(2, 16, 16, 16, 16, 16), (1, 20, 24, 12, 48, 16), (1, 24, 16, 8, 16, 48), (1, 16, 32, 16, 32, 32),
                                (1, 32, 32, 32, 16, 64), (2, 9, 20, 12, 2, 16)]:
    for dtype in (torch.bfloat16, torch.float16):
        x = torch.randn(n, d, h, w, cin, device="cuda").to(dtype)
        wt = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.1
        b = torch.randn(cout, device="cuda")
        wp = ops.pack_conv_weight_xfold(wt, dtype, False)
        ybuf = torch.zeros(n, d, h, w, cout + 16, dtype=dtype, device="cuda")
        y = ybuf[..., 8:8 + cout]
        ops.conv_fprop(x, wp, b, y, (3, 3, 3), impl=_lib.IMPL_XFOLD)
        y1 = y.float().clone()
        ops.conv_fprop(x, wp, b, y, (3, 3, 3), accumulate=True, impl=_lib.IMPL_XFOLD)
        torch.cuda.synchronize()
        print("RES", n, d, h, w, cin, cout, str(dtype), y1.double().sum().item(), y1.abs().max().item(),
              (y.float() - 2 * y1).abs().max().item(), ybuf[..., :8].abs().max().item() + ybuf[..., 8 + cout:].abs().max().item())
        torch.save(y1.cpu(), f"{OUT}/y_{n}_{d}_{h}_{w}_{cin}_{cout}_{dtype}.pt")
'''
    import tempfile
    outs = {}
    for tma in ("0", "1"):
        tmp = tempfile.mkdtemp()
        # the main loop is pinned: the operand-ring budget, hence the box cut / slot layout, depends on the epilogue's staging
        # buffers, and a different K order changes the fp32 rounding (1 ulp on ~3e-5 of the outputs), not the epilogue under test
        env = dict(os.environ, B200_EPI_TMA=tma, B200_EPI_SWZ=swz, B200_XSLAB_ALL32="0", B200_XSLAB_VARSLOT="0")
        r = subprocess.run([sys.executable, "-c", f"OUT={tmp!r}\n" + code], env=env, capture_output=True, text=True, timeout=600,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tma] = (tmp, [ln.split()[1:] for ln in r.stdout.splitlines() if ln.startswith("RES")])
    assert len(outs["0"][1]) == len(outs["1"][1]) == 12
    for a, b in zip(outs["0"][1], outs["1"][1]):
        assert a[:7] == b[:7]
        ya = torch.load(f"{outs['0'][0]}/y_{'_'.join(a[:6])}_{a[6]}.pt")
        yb = torch.load(f"{outs['1'][0]}/y_{'_'.join(b[:6])}_{b[6]}.pt")
        assert torch.equal(ya, yb), (a[:7], (ya - yb).abs().max().item(), int((ya != yb).sum()), ya.abs().max().item())   # plain stores: bit-identical
        assert float(b[10]) == 0.0                             # neighbouring channels of the slice untouched
        # accumulate: y + y, rounded once more by the element-wise add of the TMA unit
        assert float(b[9]) <= 2 ** -7 * 2 * float(b[8]) + 1e-6, b

"""GPU parity tests, one CUDA entry point at a time, through the C ABI (biapy_b200.ops -> ctypes -> kernels).

The checker is the oracle arithmetic of the reference path: PyTorch ATen on CPU in fp32 (what BiaPy itself runs,
SURVEY 8c), evaluated on the same seeded inputs.  fp32 kernels must agree to 1e-4 (normalised max error); the
bf16 storage path is compared with the same CPU computation on bf16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def cl(x):      # (N,C,D,H,W) cpu -> (N,D,H,W,C) cuda contiguous
    return x.permute(0, 2, 3, 4, 1).contiguous().cuda()


def ncdhw(x):   # (N,D,H,W,C) cuda -> (N,C,D,H,W) cpu fp32
    return x.float().permute(0, 4, 1, 2, 3).cpu()


def nerr(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-6)


CONV_CASES = [
    # n, d, h, w, cin, cout, k
    (2, 5, 9, 37, 3, 5, (3, 3, 3)),
    (1, 1, 20, 33, 16, 32, (1, 3, 3)),
    (1, 4, 8, 8, 8, 8, (1, 1, 1)),
    (1, 6, 7, 6, 2, 4, (5, 5, 5)),
    (2, 8, 8, 8, 16, 16, (3, 3, 3)),
    (1, 3, 5, 40, 48, 16, (3, 3, 3)),
    (1, 2, 4, 4, 20, 1, (1, 1, 1)),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_fprop_dgrad_wgrad_simt(case, dtype):
    from biapy_b200 import _lib, ops
    n, d, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, cin, d, h, w, generator=g)
    wt = torch.randn(cout, cin, *k, generator=g) * 0.2
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, d, h, w, generator=g)
    if dtype != torch.float32:
        x, wt, gy = x.to(dtype).float(), wt.to(dtype).float(), gy.to(dtype).float()
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br, padding=[kk // 2 for kk in k])
    (yr * gy).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 1.5e-2

    xd = cl(x).to(dtype)
    wp = ops.pack_conv_weight(wt.cuda(), dtype, False)
    # output written into a channel slice of a wider buffer (concat fusion), input read from a slice as well
    ybuf = torch.zeros(n, d, h, w, cout + 3, dtype=dtype, device="cuda")
    yv = ybuf[..., 2:2 + cout]
    xbuf = torch.zeros(n, d, h, w, cin + 5, dtype=dtype, device="cuda")
    xbuf[..., 1:1 + cin] = xd
    ops.conv_fprop(xbuf[..., 1:1 + cin], wp, b.cuda(), yv, k, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(yv), yr.detach()) < tol
    assert ybuf[..., :2].abs().max().item() == 0 and ybuf[..., 2 + cout:].abs().max().item() == 0
    # accumulate epilogue: y += conv(x) + bias
    before = ncdhw(yv)
    ops.conv_fprop(xd, wp, b.cuda(), yv, k, accumulate=True, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(yv), before + yr.detach()) < 2 * tol
    # dgrad = fprop with the flipped/transposed packing
    wpf = ops.pack_conv_weight(wt.cuda(), dtype, True)
    gyd = cl(gy).to(dtype)
    dx = torch.empty(n, d, h, w, cin, dtype=dtype, device="cuda")
    ops.conv_fprop(gyd, wpf, None, dx, k, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(dx), xr.grad) < tol
    # wgrad + bias grad
    dw = torch.empty_like(wt).cuda()
    db = torch.zeros(cout, device="cuda")
    ops.conv_wgrad(xd, gyd, cout, cin, k, dw, db, impl=_lib.IMPL_SIMT)
    assert nerr(dw.cpu(), wr.grad) < tol
    assert nerr(db.cpu(), br.grad) < tol


@pytest.mark.parametrize("cin,cout", [(16, 1), (16, 2), (32, 2), (8, 1), (8, 8), (32, 4), (1, 16), (2, 16), (2, 8), (4, 32), (3, 16),
                                      (8, 16), (1, 128)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_pointwise_narrow_layers_coalesced(cin, cout, dtype):
    """The 1-2 channel layers of the step (segmentation head 16->1 and its dgrad 1->16 / wgrad, the image-fed shortcut 2->16 and
    its wgrad) run on the coalesced CUDA-core kernels (thread = voxel x 16-byte vector): dense tensors and 16-byte-aligned
    channel slices, ragged voxel counts, plain and accumulate epilogues, against ATen on the same 16-bit-rounded operands."""
    from biapy_b200 import _lib, ops
    n, d, h, w, k = 2, 3, 7, 11, (1, 1, 1)
    g = torch.Generator().manual_seed(cin * 131 + cout)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dtype).float()
    wt = (torch.randn(cout, cin, 1, 1, 1, generator=g) * 0.3).to(dtype).float()
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, d, h, w, generator=g).to(dtype).float()
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv3d(xr, wr, br)
    (yr * gy).sum().backward()
    tol = 1.5e-2 if dtype == torch.bfloat16 else 3e-3
    padx, pady = (8 if cin % 8 == 0 else 0), (8 if cout % 8 == 0 else 0)      # aligned slices of wider (concat) buffers
    xbuf = torch.zeros(n, d, h, w, cin + 2 * padx, dtype=dtype, device="cuda")
    xv = xbuf[..., padx:padx + cin]
    xv.copy_(cl(x).to(dtype))
    ybuf = torch.zeros(n, d, h, w, cout + 2 * pady, dtype=dtype, device="cuda")
    yv = ybuf[..., pady:pady + cout]
    wp = ops.pack_conv_weight(wt.cuda(), dtype, False)
    ops.conv_fprop(xv, wp, b.cuda(), yv, k, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(yv), yr.detach()) < tol
    if pady:
        assert ybuf[..., :pady].abs().max().item() == 0 and ybuf[..., pady + cout:].abs().max().item() == 0
    before = ncdhw(yv)
    ops.conv_fprop(xv, wp, b.cuda(), yv, k, accumulate=True, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(yv), before + yr.detach()) < 2 * tol
    # dgrad: the same kernels with the roles of the channel counts swapped
    gbuf = torch.zeros(n, d, h, w, cout + 2 * pady, dtype=dtype, device="cuda")
    gv = gbuf[..., pady:pady + cout]
    gv.copy_(cl(gy).to(dtype))
    wpf = ops.pack_conv_weight(wt.cuda(), dtype, True)
    dx = torch.empty(n, d, h, w, cin, dtype=dtype, device="cuda")
    ops.conv_fprop(gv, wpf, None, dx, k, impl=_lib.IMPL_SIMT)
    assert nerr(ncdhw(dx), xr.grad) < tol
    dw = torch.empty_like(wt).cuda()
    db = torch.zeros(cout, device="cuda")
    ops.conv_wgrad(xv, gv, cout, cin, k, dw, db, impl=_lib.IMPL_SIMT)
    assert nerr(dw.cpu(), wr.grad) < tol and nerr(db.cpu(), br.grad) < tol


@pytest.mark.parametrize("window", [(2, 2, 2), (1, 2, 2)])
@pytest.mark.parametrize("dtype,c", [(torch.float32, 8), (torch.bfloat16, 16), (torch.float16, 24)])
def test_maxpool_vector_kernels(window, dtype, c):
    """The compile-time-window kernels (16-byte channel vectors, 2x2x2 / 1x2x2): ties go to the first maximum in scan order,
    plain and accumulate gradient epilogues, tensors that are channel slices of wider buffers, a NaN in one window."""
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, d, h, w = 2, 4, 6, 10
    x = torch.randint(-2, 3, (n, c, d, h, w), generator=g).float()           # exactly representable, many ties
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool3d(xr, window)
    gy = torch.randint(-4, 5, yr.shape, generator=g).float()
    (yr * gy).sum().backward()
    xbuf = torch.zeros(n, d, h, w, c + 16, dtype=dtype, device="cuda")
    xv = xbuf[..., 8:8 + c]
    xv.copy_(cl(x).to(dtype))
    od, oh, ow = d // window[0], h // window[1], w // window[2]
    ybuf = torch.zeros(n, od, oh, ow, c + 8, dtype=dtype, device="cuda")
    yv = ybuf[..., :c]
    ops.maxpool_fwd(xv, yv, window)
    assert torch.equal(ncdhw(yv), yr.detach()) and ybuf[..., c:].abs().max().item() == 0
    dxbuf = torch.full((n, d, h, w, c + 8), 3.0, dtype=dtype, device="cuda")
    dxv = dxbuf[..., 8:8 + c]
    ops.maxpool_bwd(xv, yv, cl(gy).to(dtype), dxv, window)
    assert torch.equal(ncdhw(dxv), xr.grad) and (dxbuf[..., :8] == 3.0).all()
    ops.maxpool_bwd(xv, yv, cl(gy).to(dtype), dxv, window, accumulate=True)
    assert torch.equal(ncdhw(dxv), 2 * xr.grad)
    # a NaN wins its window (ATen's rule)
    x2 = x.clone()
    x2[0, 1, 1, 2, 3] = float("nan")
    y2r = F.max_pool3d(x2, window)
    xd2 = cl(x2).to(dtype)
    y2 = torch.empty(n, od, oh, ow, c, dtype=dtype, device="cuda")
    ops.maxpool_fwd(xd2, y2, window)
    got, want = ncdhw(y2), y2r
    assert torch.equal(torch.isnan(got), torch.isnan(want)) and torch.equal(torch.nan_to_num(got), torch.nan_to_num(want))


@pytest.mark.parametrize("stride", [(2, 2, 2), (1, 2, 2), (2, 1, 1)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_convT(stride, dtype):
    from biapy_b200 import ops
    n, d, h, w, cin, cout = 2, 3, 5, 7, 12, 20
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, d, h, w, generator=g)
    wt = torch.randn(cin, cout, *stride, generator=g) * 0.3
    b = torch.randn(cout, generator=g)
    gy = torch.randn(n, cout, d * stride[0], h * stride[1], w * stride[2], generator=g)
    if dtype != torch.float32:
        x, gy = x.to(dtype).float(), gy.to(dtype).float()
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = F.conv_transpose3d(xr, wr, br, stride=stride)
    (yr * gy).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 1.5e-2
    xd, gyd = cl(x).to(dtype), cl(gy).to(dtype)
    y = torch.empty(n, d * stride[0], h * stride[1], w * stride[2], cout, dtype=dtype, device="cuda")
    ops.convT_fprop(xd, wt.cuda(), b.cuda(), y, stride)
    assert nerr(ncdhw(y), yr.detach()) < tol
    dx = torch.empty_like(xd)
    ops.convT_dgrad(gyd, wt.cuda(), dx, stride)
    assert nerr(ncdhw(dx), xr.grad) < tol
    ops.convT_dgrad(gyd, wt.cuda(), dx, stride, accumulate=True)
    assert nerr(ncdhw(dx), 2 * xr.grad) < 2 * tol
    dw, db = torch.zeros_like(wt).cuda(), torch.zeros(cout, device="cuda")
    ops.convT_wgrad(xd, gyd, dw, db, stride)
    assert nerr(dw.cpu(), wr.grad) < tol and nerr(db.cpu(), br.grad) < tol


@pytest.mark.parametrize("window", [(2, 2, 2), (1, 2, 2)])
def test_maxpool_with_ties(window):
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(2)
    # small integer values -> many ties; the gradient must go to the first maximum in scan order (ATen behaviour)
    x = torch.randint(0, 3, (2, 5, 4, 6, 8), generator=g).float()
    xr = x.clone().requires_grad_(True)
    yr = F.max_pool3d(xr, window)
    gy = torch.randn(yr.shape, generator=g)
    (yr * gy).sum().backward()
    xd = cl(x)
    y = torch.empty(2, 4 // window[0], 6 // window[1], 8 // window[2], 5, device="cuda")
    ops.maxpool_fwd(xd, y, window)
    assert torch.equal(ncdhw(y), yr.detach())
    dx = torch.empty_like(xd)
    ops.maxpool_bwd(xd, y, cl(gy), dx, window)
    assert torch.equal(ncdhw(dx), xr.grad)


@pytest.mark.parametrize("kind,c,groups", [("gn", 16, 8), ("gn", 48, 8), ("in", 6, 6), ("in", 1, 1)])
@pytest.mark.parametrize("act", ["silu", "elu", "relu", "none", "sigmoid"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_norm_act_fwd_bwd(kind, c, groups, act, dtype):
    from biapy_b200 import ops
    n, d, h, w = 2, 4, 6, 10
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n, c, d, h, w, generator=g) * 1.7 + 0.4
    gamma = torch.randn(c, generator=g)
    beta = torch.randn(c, generator=g)
    gy = torch.randn(n, c, d, h, w, generator=g)
    if dtype != torch.float32:
        x, gy = x.to(dtype).float(), gy.to(dtype).float()
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yn = F.group_norm(xr, groups, gr, br, 1e-5)
    fn = {"silu": F.silu, "elu": F.elu, "relu": F.relu, "none": lambda t: t, "sigmoid": torch.sigmoid}[act]
    yr = fn(yn)
    (yr * gy).sum().backward()
    tol = 2e-4 if dtype == torch.float32 else 2e-2
    xd = cl(x).to(dtype)
    st = ops.norm_stats(xd, groups, gamma.cuda(), beta.cuda())
    y = torch.empty_like(xd)
    ops.scale_shift_act(xd, st.scale, st.shift, act, y)
    assert nerr(ncdhw(y), yr.detach()) < tol
    dx = torch.empty_like(xd)
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    ops.norm_act_bwd(xd, cl(gy).to(dtype), st, gamma.cuda(), beta.cuda(), act, dx, dg, db)
    assert nerr(ncdhw(dx), xr.grad) < tol
    assert nerr(dg.cpu(), gr.grad) < tol and nerr(db.cpu(), br.grad) < tol


@pytest.mark.parametrize("c,groups,pad", [(16, 8, 0), (48, 8, 16), (96, 8, 0)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("bwd", ["g", "recompute"])
def test_norm_silu_fast_chain(c, groups, pad, dtype, bwd, monkeypatch):
    """The one-MUFU SiLU chain (norm_fast.cuh) against ATen in fp32 on operands rounded to the engine dtype: forward, dx (plain and
    accumulated into a channel slice), dgamma / dbeta; dy is overwritten by g = dy * silu'(z) as documented."""
    from biapy_b200 import ops
    monkeypatch.setattr(ops, "NORM_FAST", "1")
    monkeypatch.setattr(ops, "NORM_BWD", bwd)
    n, d, h, w = 2, 6, 6, 10
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(n, c, d, h, w, generator=g) * 1.7 + 0.4).to(dtype).float()
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    gy = torch.randn(n, c, d, h, w, generator=g).to(dtype).float()
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.silu(F.group_norm(xr, groups, gr, br, 1e-5))
    (yr * gy).sum().backward()
    xbuf = torch.zeros(n, d, h, w, c + pad, dtype=dtype, device="cuda")
    xd = xbuf[..., pad:]
    xd.copy_(cl(x).to(dtype))
    assert ops.norm_fast_ok(xd)
    st = ops.norm_stats(xd, groups, gamma.cuda(), beta.cuda())
    y = torch.empty(n, d, h, w, c, dtype=dtype, device="cuda")
    ops.scale_shift_act(xd, st.scale, st.shift, "silu", y)
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert nerr(ncdhw(y), yr.detach()) < tol
    for accumulate in (False, True):
        dy = cl(gy).to(dtype)
        dy0 = dy.clone()
        dxbuf = torch.ones(n, d, h, w, c + pad, dtype=dtype, device="cuda")
        dx = dxbuf[..., pad:]
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        ops.norm_act_bwd(xd, dy, st, gamma.cuda(), beta.cuda(), "silu", dx, dg, db, accumulate=accumulate, dy_dead=True)
        assert torch.equal(dy, dy0) == (bwd == "recompute")             # 'g': g left in place of dy; 'recompute': dy untouched
        want = xr.grad + (1.0 if accumulate else 0.0)
        assert nerr(ncdhw(dx), want) < tol
        assert nerr(dg.cpu(), gr.grad) < tol and nerr(db.cpu(), br.grad) < tol
        if pad:
            assert torch.equal(dxbuf[..., :pad], torch.ones_like(dxbuf[..., :pad]))


def test_act_only_and_elementwise():
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(2, 7, 3, 4, 5, generator=g), torch.randn(2, 7, 3, 4, 5, generator=g)
    ad, bd = cl(a), cl(b)
    y = torch.empty_like(ad)
    for op, ref in ((ops.OP_ADD, a + b), (ops.OP_MUL, a * b), (ops.OP_ADD_RELU, F.relu(a + b))):
        ops.binary(ad, bd, y, op)
        assert nerr(ncdhw(y), ref) < 1e-6
    ops.binary(ad, None, y, ops.OP_SIGMOID)
    assert nerr(ncdhw(y), torch.sigmoid(a)) < 1e-6
    # gate: out = psi * x and its backward
    psi = torch.rand(2, 1, 3, 4, 5, generator=g)
    xr, pr = a.clone().requires_grad_(True), psi.clone().requires_grad_(True)
    (pr * xr * b).sum().backward()
    ops.binary(ad, cl(psi), y, ops.OP_MUL)
    assert nerr(ncdhw(y), (psi * a)) < 1e-6
    dpsi, dx = torch.empty(2, 3, 4, 5, 1, device="cuda"), torch.empty_like(ad)
    ops.gate_bwd(ad, cl(psi), bd, dpsi, dx)
    assert nerr(ncdhw(dx), xr.grad) < 1e-6 and nerr(ncdhw(dpsi), pr.grad) < 1e-5
    for act in ("silu", "elu", "relu", "leaky_relu", "gelu", "tanh", "sigmoid", "softplus"):
        ar = a.clone().requires_grad_(True)
        fn = {"silu": F.silu, "elu": F.elu, "relu": F.relu, "leaky_relu": F.leaky_relu, "gelu": F.gelu, "tanh": torch.tanh,
              "sigmoid": torch.sigmoid, "softplus": F.softplus}[act]
        out = fn(ar)
        (out * b).sum().backward()
        ops.scale_shift_act(ad, None, None, act, y)
        assert nerr(ncdhw(y), out.detach()) < 2e-6, act
        ops.act_bwd(ad, bd, act, dx)
        assert nerr(ncdhw(dx), ar.grad) < 2e-6, act


def test_losses():
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(5)
    z = torch.randn(2, 1, 4, 6, 8, generator=g) * 3
    t = (torch.rand(2, 1, 4, 6, 8, generator=g) < 0.3).float()
    zr = z.clone().requires_grad_(True)
    loss = F.binary_cross_entropy_with_logits(zr, t)
    loss.backward()
    zd = cl(z)
    dz = torch.empty_like(zd)
    s = ops.bce_logits(zd, cl(t).contiguous(), dz, grad_scale=1.0 / z.numel())
    assert abs(s.item() / z.numel() - loss.item()) < 1e-6
    assert nerr(ncdhw(dz), zr.grad) < 1e-5
    # N2V masked MSE
    C = 2
    y = torch.randn(2, C, 3, 4, 5, generator=g)
    tgt = torch.randn(2, 2 * C, 3, 4, 5, generator=g)
    tgt[:, C:] = (torch.rand(2, C, 3, 4, 5, generator=g) < 0.2).float()
    yr = y.clone().requires_grad_(True)
    l = torch.sum(torch.square(tgt[:, :C] - yr * tgt[:, C:])) / torch.sum(tgt[:, C:])
    l.backward()
    yd, td = cl(y), cl(tgt).contiguous()
    sums = ops.n2v_mse_sums(yd, td)
    assert abs((sums[0] / sums[1]).item() - l.item()) < 1e-5
    dy = torch.empty_like(yd)
    ops.n2v_mse_bwd(yd, td, dy, 1.0 / sums[1].item())
    assert nerr(ncdhw(dy), yr.grad) < 1e-5
    # softmax cross entropy
    zc = torch.randn(2, 4, 3, 4, 5, generator=g)
    cls = torch.randint(0, 4, (2, 3, 4, 5), generator=g)
    zcr = zc.clone().requires_grad_(True)
    lc = F.cross_entropy(zcr, cls)
    lc.backward()
    zcd = cl(zc)
    dzc = torch.empty_like(zcd)
    sc = ops.softmax_ce(zcd, cls.cuda().contiguous(), dzc, grad_scale=1.0 / cls.numel())
    assert abs(sc[0].item() / cls.numel() - lc.item()) < 1e-5 and sc[1].item() == cls.numel() and sc[2].item() == 0
    assert nerr(ncdhw(dzc), zcr.grad) < 1e-5
    # CrossEntropyLoss(ignore_index) of the reference's wrapper (metrics.py:534-546): mean over the counted voxels only
    cls_i = cls.clone()
    cls_i[:, 0] = -100
    zcr = zc.clone().requires_grad_(True)
    li = F.cross_entropy(zcr, cls_i, ignore_index=-100)
    li.backward()
    si = ops.softmax_ce(zcd, cls_i.cuda().contiguous(), dzc, grad_scale=1.0, ignore_index=-100)
    cnt = si[1].item()
    assert cnt == (cls_i != -100).sum().item() and si[2].item() == 0 and abs(si[0].item() / cnt - li.item()) < 1e-5
    assert nerr(ncdhw(dzc) / cnt, zcr.grad) < 1e-5
    cls_b = cls.clone()
    cls_b[0, 0, 0, 0], cls_b[1, 2, 3, 4] = 255, -7                 # illegal labels: skipped and reported, never dereferenced
    sb = ops.softmax_ce(zcd, cls_b.cuda().contiguous(), dzc, grad_scale=1.0)
    assert sb[2].item() == 2 and sb[1].item() == cls.numel() - 2
    with pytest.raises(Exception):
        ops.softmax_ce(zcd, cls.cuda()[:1].contiguous(), dzc)      # short target buffer
    # one-pass Noise2Void form used by the Trainer
    dy2 = torch.empty_like(yd)
    s2 = ops.n2v_mse_fused(yd, td, dy2, 0.25)
    assert torch.equal(s2.cpu(), sums.cpu()) and nerr(ncdhw(dy2) / (0.25 * sums[1].item()), yr.grad) < 1e-5
    out = torch.empty_like(zcd)
    ops.softmax_channels(zcd, out, 0, 4)
    assert nerr(ncdhw(out), torch.softmax(zc, 1)) < 1e-6


def test_optimizers_match_torch():
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(6)
    p0 = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(5)]
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.adamw_step(p, gr.cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.02, i + 1)
    assert nerr(p.cpu(), pr.detach()) < 1e-6
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.9, weight_decay=1e-4)
    p = p0.clone().cuda()
    mom = torch.zeros_like(p)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.sgd_step(p, gr.cuda(), mom, 1e-2, 0.9, 1e-4, i == 0)
    assert nerr(p.cpu(), pr.detach()) < 1e-6
    # TRAIN.OPTIMIZER = 'SGD' is timm's SGD(momentum=0.9, nesterov=True); 'ADAM' is torch.optim.Adam (L2 decay)
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], lr=1e-2, momentum=0.9, weight_decay=1e-4, nesterov=True)
    p = p0.clone().cuda()
    mom = torch.zeros_like(p)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.sgd_step(p, gr.cuda(), mom, 1e-2, 0.9, 1e-4, i == 0, nesterov=True)
    assert nerr(p.cpu(), pr.detach()) < 1e-6
    pr = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step(p, gr.cuda(), m, v, 1e-3, 0.9, 0.999, 1e-8, 0.02, i + 1)
    assert nerr(p.cpu(), pr.detach()) < 1e-6
    # device-hyper-parameter form (what the Trainer launches): clipping, an overflowed step skipped, step counter on the device
    for kind, mk in (("adamw", lambda q: torch.optim.AdamW([q], lr=1e-3, weight_decay=0.02)),
                     ("adam", lambda q: torch.optim.Adam([q], lr=1e-3, weight_decay=0.02)),
                     ("sgd", lambda q: torch.optim.SGD([q], lr=1e-3, momentum=0.9, weight_decay=0.02, nesterov=True))):
        pr = p0.clone().requires_grad_(True)
        opt = mk(pr)
        p = p0.clone().cuda()
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        hp = torch.zeros(ops.HP_SIZE, device="cuda")
        state = torch.zeros(2, dtype=torch.int64, device="cuda")
        derived = torch.zeros(8, device="cuda")
        gsq = torch.zeros(1, dtype=torch.float64, device="cuda")
        ops.write_floats(hp, (1e-3, 0.9, 0.999, 1e-8, 0.02, 0.9 if kind == "sgd" else 0.0, 1.0 if kind == "sgd" else 0.0, 0.5, 2.0))
        for i, gr in enumerate(grads):
            pr.grad = 0.5 * gr.clone()
            torch.nn.utils.clip_grad_norm_([pr], 2.0)
            opt.step()
            gd = gr.cuda()
            if i == 2:                                        # a non-finite gradient is skipped (GradScaler semantics)
                bad = gd.clone()
                bad[5] = float("inf")
                before = p.clone()
                gsq.zero_()
                ops.sumsq(bad, out=gsq)
                ops.optim_step_dev(kind, p, bad, m, v, hp, state, derived, gsq=gsq)
                assert torch.equal(p, before) and state.tolist() == [2, 1]
            gsq.zero_()
            ops.sumsq(gd, out=gsq)
            ops.optim_step_dev(kind, p, gd, m, v, hp, state, derived, gsq=gsq)
        assert state.tolist() == [5, 1]
        assert nerr(p.cpu(), pr.detach()) < 1e-6, kind
    ss = ops.sumsq(grads[0].cuda())
    assert abs(ss.item() - float((grads[0].double() ** 2).sum())) < 1e-6 * ss.item()


@pytest.mark.parametrize("shape,scale", [((2, 8, 3, 4, 5), (2, 2, 2)), ((1, 3, 4, 6, 6), (1, 2, 2)), ((2, 16, 1, 7, 9), (1, 2, 2)),
                                         ((1, 4, 3, 3, 3), (3, 2, 1))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample_linear(shape, scale, dtype):
    """nn.Upsample(bi/trilinear, align_corners=False) forward and its adjoint vs ATen (reference blocks.py:605)."""
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g).to(dtype).float().requires_grad_(True)          # (N, C, D, H, W)
    yr = F.interpolate(x, scale_factor=tuple(float(s) for s in scale), mode="trilinear", align_corners=False)
    gy = torch.randn(yr.shape, generator=g).to(dtype).float()
    yr.backward(gy)
    xd = cl(x.detach()).to(dtype)
    n, c, d, h, w = shape
    ybuf = torch.zeros(n, d * scale[0], h * scale[1], w * scale[2], c + 8, dtype=dtype, device="cuda")
    yv = ybuf[..., 8:]                                                                    # a channel slice, as in the decoder
    ops.upsample_linear_fwd(xd, yv)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert nerr(ncdhw(yv), yr.detach()) < tol
    assert ybuf[..., :8].abs().max().item() == 0
    dx = torch.empty_like(xd)
    ops.upsample_linear_bwd(cl(gy).to(dtype), dx)
    assert nerr(ncdhw(dx), x.grad) < (1e-5 if dtype == torch.float32 else 2e-2)
    before = ncdhw(dx)
    ops.upsample_linear_bwd(cl(gy).to(dtype), dx, accumulate=True)
    assert nerr(ncdhw(dx), 2 * before) < 2e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_mask_statistics_and_backward(dtype):
    """nn.Dropout training semantics: values are 0 or x/(1-p), keep rate 1-p, the backward reuses the forward mask, masks
    change with the layer id and with the device-resident seed."""
    from biapy_b200 import ops
    p = 0.3
    x = torch.ones(2, 8, 16, 16, 16, dtype=dtype, device="cuda")
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    y = ops.dropout(x, torch.empty_like(x), p, seed, 0).float()
    kept = y != 0
    assert torch.allclose(y[kept], torch.full_like(y[kept], 1 / (1 - p)), rtol=1e-2)
    n = y.numel()
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 5 * (p * (1 - p) / n) ** 0.5
    y_again = ops.dropout(x, torch.empty_like(x), p, seed, 0).float()
    assert torch.equal(y, y_again)                                   # same (seed, layer) -> same mask: this is the backward
    y_other = ops.dropout(x, torch.empty_like(x), p, seed, 1).float()
    agree = ((y_other != 0) == kept).float().mean().item()
    assert abs(agree - (p * p + (1 - p) * (1 - p))) < 0.02           # independent masks
    seed.add_(1)
    y_next = ops.dropout(x, torch.empty_like(x), p, seed, 0).float()
    assert not torch.equal(y_next, y)
    dy = torch.randn_like(x)
    dx = ops.dropout(dy, torch.empty_like(x), p, seed, 0).float()
    assert torch.allclose(dx, dy.float() * (y_next != 0) / (1 - p), rtol=2e-2, atol=1e-3)


def test_pack_batch_equals_the_single_job_kernels():
    """b200_pack_batch (one launch for all weight packs / weight-gradient un-packs of a training pass) against the single-job
    kernels it replaces, bit for bit, on a mixed job list long enough to need two launches (> 48 jobs)."""
    from biapy_b200 import ops
    g = torch.Generator().manual_seed(21)
    jobs, want = [], []
    for rep in range(6):
        for cout, cin, k in ((16, 16, (3, 3, 3)), (32, 16, (3, 3, 3)), (16, 48, (1, 3, 3)), (64, 32, (1, 1, 1)), (16, 2, (3, 3, 3))):
            w = torch.randn(cout, cin, *k, generator=g).cuda()
            for flip in (False, True):
                if cin % 16 == 0 or not flip:
                    ref = ops.pack_conv_weight_xfold(w, torch.bfloat16, flip)
                    jobs.append((ops.PACK_XFOLD, w, torch.zeros_like(ref), cout, cin, k[0], k[1], k[2], int(flip)))
                    want.append(ref)
                ref = ops.pack_conv_weight(w, torch.bfloat16, flip)
                jobs.append((ops.PACK_PLAIN, w, torch.zeros_like(ref), cout, cin, k[0], k[1], k[2], int(flip)))
                want.append(ref)
        wt = torch.randn(32, 16, 2, 2, 2, generator=g).cuda()                    # transposed conv (Cin, Cout, *s)
        for for_dgrad in (False, True):
            ref = ops.pack_convT_weight(wt, torch.bfloat16, for_dgrad)
            jobs.append((ops.PACK_CONVT, wt, torch.zeros_like(ref), 32, 16, 8, 1, 1, int(for_dgrad)))
            want.append(ref)
    assert len(jobs) > 48
    n0 = ops.LAUNCHES
    ops.pack_batch(jobs, torch.bfloat16)
    assert ops.LAUNCHES - n0 == (len(jobs) + 47) // 48
    for j, ref in zip(jobs, want):
        assert torch.equal(j[2], ref), j[0]
    # un-packs: queued by conv_wgrad / convT_wgrad_tc inside a Trainer pass, here driven directly
    packed = torch.randn(16 * 27 * 32, generator=g).cuda()
    dw0 = torch.randn(16, 32, 3, 3, 3, generator=g).cuda()
    ref = dw0.clone()
    ops._launch("b200_unpack_conv_wgrad", ops._ptr(packed), ops._ptr(ref), 16, 32, 27, 1, _stream())
    got = dw0.clone()
    ops.UNPACK_QUEUE = [(ops.UNPACK_WGRAD, packed, got, 16, 32, 27, 1, 1, 1)]
    try:
        ops.flush_unpacks()
    finally:
        ops.UNPACK_QUEUE = None
    assert torch.equal(got, ref)


def _stream():
    from biapy_b200 import _lib
    return _lib.stream_ptr()


def test_memset_zero_and_sums_through_pointwise():
    """b200_memset_zero (cudaMemsetAsync behind the C ABI) and b200_sums_through_pointwise (xsum += W^T dysum)."""
    from biapy_b200 import ops
    t = torch.randn(1000, device="cuda")
    ops.zero_(t[10:900])
    assert t[10:900].abs().max().item() == 0.0 and t[:10].abs().min().item() > 0.0 and t[900:].abs().min().item() > 0.0
    w = torch.randn(48, 200, 1, 1, 1, device="cuda")
    dys = torch.randn(48, device="cuda")
    xs = torch.randn(200, device="cuda")
    want = xs.double() + w.view(48, 200).double().t() @ dys.double()
    ops.sums_through_pointwise(w, dys, xs)
    assert ((xs.double() - want).abs().max() / want.abs().max()).item() < 1e-6

"""Accuracy probe: per-parameter gradient error of an Attention U-Net on the fp32 engine vs ATen fp32, both against ATen fp64."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port_models
from biapy_b200.models.attention_unet import Attention_U_Net
ne = lambda p, q: (p.double().cpu() - q).abs().max().item() / max(q.abs().max().item(), 1e-30)
for fm, norm in (([16, 32], "none"), ([16, 32], "in"), ([16, 32, 64], "in")):
    size = 64
    kw = dict(image_shape=(size, size, size, 1), activation="elu", feature_maps=fm, drop_values=[0] * len(fm), normalization=norm, k_size=3,
              yx_down=[2] * (len(fm) - 1), z_down=[2] * (len(fm) - 1), isotropy=[True] * len(fm), larger_io=False, conv_layers=[2] * len(fm), output_channels=[1])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Attention_U_Net(**kw)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in m.parameters():
            if p.ndim == 1: p.add_(0.2 * torch.randn(p.shape, generator=g))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    x = torch.randn((2, 1, size, size, size), generator=g)
    gy = torch.randn((2, 1, size, size, size), generator=g)
    res = {}
    for dt in (torch.float32, torch.float64):
        sd_r = {k: v.clone().to(dt).requires_grad_(True) for k, v in sd.items()}
        xr = x.clone().to(dt).requires_grad_(True)
        yr = port_models.forward("attention_unet", sd_r, xr, training=True, **kw)
        (yr * gy.to(dt)).sum().backward()
        res[dt] = (xr.grad.double(), {k: v.grad.double() for k, v in sd_r.items() if v.grad is not None})
    m = m.cuda().set_engine(dtype=torch.float32)
    xc = x.cuda().requires_grad_(True)
    y = m(xc)
    (y * gy.cuda()).sum().backward()
    t = res[torch.float64]
    print(f"== fm{len(fm)} norm={norm}: dx ours {ne(xc.grad, t[0]):.2e} aten32 {ne(res[torch.float32][0], t[0]):.2e}", flush=True)
    scale = max(v.abs().max().item() for v in t[1].values())
    rows = []
    for n, p in m.named_parameters():
        eo = (p.grad.double().cpu() - t[1][n]).abs().max().item() / scale
        ea = (res[torch.float32][1][n] - t[1][n]).abs().max().item() / scale
        rows.append((eo / max(ea, 1e-12), n, eo, ea))
    for r in sorted(rows, reverse=True)[:8]:
        print(f"   {r[1]:<48s} ours {r[2]:.2e} aten32 {r[3]:.2e} ratio {r[0]:.1f}", flush=True)

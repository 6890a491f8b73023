"""Accuracy probe: gradient w.r.t. the network input of the 64^3 Attention U-Net, fp32 engine vs ATen fp32 vs ATen fp64."""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port_models
from biapy_b200.models.attention_unet import Attention_U_Net
from biapy_b200.models.unet import U_Net
ne = lambda p, q: (p - q).abs().max().item() / q.abs().max().item()
for arch, cls, fm, size in (("unet", U_Net, [16, 32], 64), ("unet", U_Net, [16, 32, 64, 128, 256], 64), ("attention_unet", Attention_U_Net, [16, 32, 64, 128, 256], 64)):
    for norm in ("in", "none"):
        kw = dict(image_shape=(size, size, size, 1), activation="elu", feature_maps=fm, drop_values=[0] * len(fm), normalization=norm, k_size=3,
                  yx_down=[2] * (len(fm) - 1), z_down=[2] * (len(fm) - 1), isotropy=[True] * len(fm), larger_io=False, conv_layers=[2] * len(fm), output_channels=[1])
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            m = cls(**kw)
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
            yr = port_models.forward(arch, sd_r, xr, training=True, **kw)
            (yr * gy.to(dt)).sum().backward()
            res[dt] = (yr.detach().double(), xr.grad.double())
        m = m.cuda().set_engine(dtype=torch.float32)
        xc = x.cuda().requires_grad_(True)
        y = m(xc)
        (y * gy.cuda()).sum().backward()
        ours = (y.detach().cpu().double(), xc.grad.cpu().double())
        print(f"{arch} fm{len(fm)} norm={norm}: fwd ours-vs-64 {ne(ours[0], res[torch.float64][0]):.2e} aten32-vs-64 {ne(res[torch.float32][0], res[torch.float64][0]):.2e} | "
              f"dx ours-vs-64 {ne(ours[1], res[torch.float64][1]):.2e} aten32-vs-64 {ne(res[torch.float32][1], res[torch.float64][1]):.2e} ours-vs-aten32 {ne(ours[1], res[torch.float32][1]):.2e}", flush=True)

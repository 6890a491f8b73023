"""Accuracy probe of the attention gate (blocks.py:1014-1116) on the fp32 engine against a float64 torch evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from biapy_b200.engine.tape import TT, Tape
from biapy_b200.models.blocks import AttentionBlock
from biapy_b200 import _lib
ne = lambda p, q: (p.double().cpu() - q.double().cpu()).abs().max().item() / max(q.abs().max().item(), 1e-30)
for norm in ("none", "in"):
    for size in (16, 64):
        torch.manual_seed(0)
        blk = AttentionBlock(nn.Conv3d, 16, 8, norm=norm).cuda()
        with torch.no_grad():
            for p in blk.parameters():
                if p.ndim == 1: p.add_(0.2 * torch.randn_like(p))
        g = torch.randn(2, size, size, size, 16, device="cuda") + 0.5
        x = torch.randn(2, size, size, size, 16, device="cuda")
        up = torch.randn(2, size, size, size, 16, device="cuda")
        tape = Tape(torch.float32, g.device, training=True)
        gt, xt = TT(g.clone()), TT(x.clone())
        out = tape.gate if False else blk.run(tape, gt, xt)
        out.grad().copy_(up); out.mark_written()
        tape.backward()
        # float64 torch
        P = {k: v.detach().double() for k, v in blk.state_dict().items()}
        g64 = g.double().permute(0, 4, 1, 2, 3).requires_grad_(True); x64 = x.double().permute(0, 4, 1, 2, 3).requires_grad_(True)
        W = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        g1 = F.conv3d(g64, W["w_g.0.weight"], W["w_g.0.bias"])
        if norm == "in": g1 = F.instance_norm(g1, weight=W["w_g.1.weight"], bias=W["w_g.1.bias"])
        x1 = F.conv3d(x64, W["w_x.0.weight"], W["w_x.0.bias"])
        s = F.relu(g1 + x1)
        p = F.conv3d(s, W["psi.0.weight"], W["psi.0.bias"])
        if norm == "in": p = F.instance_norm(p, weight=W["psi.1.weight"], bias=W["psi.1.bias"])
        o = torch.sigmoid(p) * x64
        o.backward(up.double().permute(0, 4, 1, 2, 3))
        print(f"norm={norm} size={size}: out {ne(out.data.permute(0,4,1,2,3), o.detach()):.2e}  dg {ne(gt.grad().permute(0,4,1,2,3), g64.grad):.2e}  dx {ne(xt.grad().permute(0,4,1,2,3), x64.grad):.2e}  "
              + "  ".join(f"{k} {ne(tape.param_grads[dict(blk.named_parameters())[k]], W[k].grad):.1e}" for k in W if W[k].grad is not None and k in dict(blk.named_parameters())), flush=True)

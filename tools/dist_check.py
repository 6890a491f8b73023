"""Multi-GPU checks on real NCCL (run under torchrun, 2+ ranks):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

1. training: different batches per rank, one gradient all-reduce per step -> identical parameters on every rank, and the
   averaged gradient equals the gradient a single process computes on the concatenated batch (fp32 engine);
2. SyncBatchNorm: statistics all-reduced across ranks == BatchNorm on the concatenated batch;
3. sliding-window inference (`predict_volume`, `predict_by_chunks`) with world > 1 == the single-rank result, bit for bit.
Prints one line per check from rank 0 and exits non-zero on a mismatch.
"""
import contextlib
import io
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

from biapy_b200.engine.inference import predict_by_chunks, predict_volume
from biapy_b200.engine.train import Trainer
from biapy_b200.models.resunet import ResUNet
from biapy_b200.models.unet import U_Net


def say(*a):
    if rank == 0:
        print("[dist_check]", *a, flush=True)


def build(cls, seed, **kw):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return cls(**kw)


ok = True
KW = dict(image_shape=(32, 32, 32, 2), activation="silu", feature_maps=[16, 32, 64], drop_values=[0, 0, 0], normalization="gn",
          k_size=3, yx_down=[2, 2], z_down=[2, 2], isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1])

# ---- 1. training -------------------------------------------------------------------------------------------------
g = torch.Generator().manual_seed(100)
xs = torch.randn(world, 2, 32, 32, 32, 2, generator=g)
ts = (torch.rand(world, 2, 32, 32, 32, 1, generator=g) < 0.3).float()
m = build(ResUNet, 0, **KW).cuda().set_engine(dtype=torch.float32)
tr = Trainer(m, loss="bce", optimizer="sgd", lr=0.0)
tr.step(xs[rank].cuda(), ts[rank].cuda())
torch.cuda.synchronize()
grad_dp = tr.fp.grad.clone() / world                         # the all-reduced sum, scaled like the optimiser kernel does
ref = build(ResUNet, 0, **KW).cuda().set_engine(dtype=torch.float32)
tr1 = Trainer(ref, loss="bce", optimizer="sgd", lr=0.0)
tr1.world, tr1.pg = 1, None
_saved = torch.distributed.all_reduce
torch.distributed.all_reduce = lambda *a, **k: None          # single-process reference on the concatenated batch
try:
    tr1.step(xs.reshape(-1, 32, 32, 32, 2).cuda(), ts.reshape(-1, 32, 32, 32, 1).cuda())
finally:
    torch.distributed.all_reduce = _saved
torch.cuda.synchronize()
err = ((grad_dp - tr1.fp.grad).norm() / tr1.fp.grad.norm()).item()
say(f"training: data-parallel gradient vs single process on the concatenated batch: rel-L2 {err:.2e}")
ok &= err < 1e-5
# per-rank seeds, as the reference initialises its replicas (misc.set_seed: SEED + rank): the Trainer must start every rank from
# rank 0's parameters the way DistributedDataParallel does at construction
m2 = build(ResUNet, 10 + rank, **KW).cuda().set_engine(dtype=torch.bfloat16)
tr2 = Trainer(m2, loss="bce", optimizer="adamw", lr=1e-3, weight_decay=0.02)
for _ in range(3):
    loss = tr2.step(xs[rank].numpy(), ts[rank].numpy())
flat = tr2.fp.flat.clone()
lo, hi = flat.clone(), flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same = bool(torch.equal(lo, hi))
say(f"training: replicas built from seeds 10 + rank, 3 AdamW steps (bf16 engine), parameters identical on all {world} ranks: {same}; loss {loss.item():.4f}")
ok &= same

# the captured step (CUDA graph): with more than one rank the pass is two graphs and the all-reduce of the flat gradient's tail runs
# under the second one (Trainer.enable_cuda_graph) -- the reduced gradient must be the eager step's, and replicas must stay identical
m3 = build(ResUNet, 0, **KW).cuda().set_engine(dtype=torch.float32)
tr3 = Trainer(m3, loss="bce", optimizer="sgd", lr=0.0)
tr3.enable_cuda_graph(xs[rank].cuda(), ts[rank].cuda())
for _ in range(2):
    tr3.step(xs[rank].cuda(), ts[rank].cuda())
torch.cuda.synchronize()
err3 = ((tr3.fp.grad / world - tr1.fp.grad).norm() / tr1.fp.grad.norm()).item()
say(f"training: captured step, split {tr3._split} (tape step, flat offset of {tr3.fp.grad.numel()}), two graphs: {tr3._graph_b is not None}; "
    f"gradient vs single process: rel-L2 {err3:.2e}")
ok &= err3 < 1e-5 and (tr3._graph_b is not None)
m4 = build(ResUNet, 10 + rank, **KW).cuda().set_engine(dtype=torch.bfloat16)
tr4 = Trainer(m4, loss="bce", optimizer="adamw", lr=1e-3, weight_decay=0.02)
tr4.enable_cuda_graph(xs[rank].numpy(), ts[rank].numpy())
for _ in range(3):
    loss4 = tr4.step(xs[rank].numpy(), ts[rank].numpy())
torch.cuda.synchronize()
flat4 = tr4.fp.flat.clone()
lo, hi = flat4.clone(), flat4.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same4 = bool(torch.equal(lo, hi))
d4 = ((flat4 - flat).norm() / flat.norm()).item()
say(f"training: captured + overlapped, 3 AdamW steps (bf16 engine): parameters identical on all {world} ranks: {same4}; against the eager "
    f"steps: rel-L2 {d4:.2e}; loss {loss4.item():.4f}")
ok &= same4 and d4 < 1e-3

# ---- 2. SyncBatchNorm --------------------------------------------------------------------------------------------
KB = dict(image_shape=(32, 32, 1), activation="relu", feature_maps=[16, 32], drop_values=[0, 0], normalization="sync_bn", k_size=3,
          yx_down=[2], z_down=[2], isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1])
xb = torch.randn(world, 2, 1, 32, 32, generator=g)
ms = build(U_Net, 3, **KB).cuda().set_engine(dtype=torch.float32).train()
y = ms(xb[rank].cuda())
mb = build(U_Net, 3, **dict(KB, normalization="bn")).cuda().set_engine(dtype=torch.float32).train()
mb.load_state_dict(build(U_Net, 3, **KB).state_dict())
yb = mb(xb.reshape(-1, 1, 32, 32).cuda())[2 * rank:2 * rank + 2]
e = (y - yb).abs().max().item() / yb.abs().max().item()
rm_s = dict(ms.named_buffers())["down_path.0.block.0.block.1.running_mean"]
rm_b = dict(mb.named_buffers())["down_path.0.block.0.block.1.running_mean"]
e2 = (rm_s - rm_b).abs().max().item()
emax = torch.tensor([e, e2], device="cuda")
dist.all_reduce(emax, op=dist.ReduceOp.MAX)
say(f"SyncBatchNorm over {world} ranks vs BatchNorm on the concatenated batch: output {emax[0].item():.2e}, running_mean {emax[1].item():.2e}")
ok &= emax[0].item() < 1e-5 and emax[1].item() < 1e-6

# ---- 3. inference ------------------------------------------------------------------------------------------------
mi = build(ResUNet, 1, **KW).cuda().set_engine(dtype=torch.float32).eval()
vol = np.random.default_rng(7).standard_normal((72, 64, 80, 2)).astype(np.float32)
patch, ov, pad = (32, 32, 32, 2), (0.25, 0.25, 0.25), (4, 0, 2)
one = predict_volume(mi, vol, patch, overlap=ov, padding=pad, batch_size=3, head_activations=["ce_sigmoid"])
many = predict_volume(mi, vol, patch, overlap=ov, padding=pad, batch_size=3, head_activations=["ce_sigmoid"], rank=rank, world=world)
eq = torch.tensor([float(np.array_equal(one, many))], device="cuda")
dist.all_reduce(eq, op=dist.ReduceOp.MIN)
say(f"predict_volume: world={world} result bit-identical to world=1 on every rank: {bool(eq.item())}")
ok &= bool(eq.item())
# the volume left sharded, every rank fed only the planes its patches read (what a by-chunks reader would load)
from biapy_b200.data import _stitch
from biapy_b200.engine.inference import shard_planes
a, b = shard_planes(vol.shape, patch, ov, pad, "reflect", rank, world)
st = {}
slab, (z0, z1) = predict_volume(mi, _stitch.VolumeShard(vol[a:b], a, vol.shape[0]), patch, overlap=ov, padding=pad, batch_size=3,
                                head_activations=["ce_sigmoid"], rank=rank, world=world, gather="none", stats=st)
eq = torch.tensor([float(np.array_equal(one[z0:z1], slab))], device="cuda")
dist.all_reduce(eq, op=dist.ReduceOp.MIN)
say(f"predict_volume(gather='none') from volume shards [{a},{b}) of {vol.shape[0]} planes: slab [{z0},{z1}) bit-identical: "
    f"{bool(eq.item())}; rank 0 received {st.get('exchange_bytes_received')} bytes of patch pieces")
ok &= bool(eq.item())
r0 = predict_volume(mi, vol, patch, overlap=ov, padding=pad, batch_size=3, head_activations=["ce_sigmoid"], rank=rank, world=world,
                    gather="rank0")
eq = torch.tensor([float((r0 is None) if rank else np.array_equal(one, r0))], device="cuda")
dist.all_reduce(eq, op=dist.ReduceOp.MIN)
say(f"predict_volume(gather='rank0'): full volume on rank 0 only, bit-identical: {bool(eq.item())}")
ok &= bool(eq.item())
one = predict_by_chunks(mi, vol, patch, padding=(4, 4, 4), batch_size=3, head_activations=["ce_sigmoid"])
many = predict_by_chunks(mi, vol, patch, padding=(4, 4, 4), batch_size=3, head_activations=["ce_sigmoid"], rank=rank, world=world)
eq = torch.tensor([float(np.array_equal(one, many))], device="cuda")
dist.all_reduce(eq, op=dist.ReduceOp.MIN)
say(f"predict_by_chunks: world={world} result bit-identical to world=1 on every rank: {bool(eq.item())}")
ok &= bool(eq.item())

flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
say("ALL OK" if flag.item() == 1.0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)

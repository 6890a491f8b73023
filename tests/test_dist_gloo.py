"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: gradient all-reduce factor, patch-range dealing and the slab
exchange of the sharded sliding-window inference (every rank ends with exactly the pieces its output slab needs), by-chunks tile
dealing, cross-rank loss statistics of the epoch loop and of the validation pass."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from biapy_b200.engine import dist as bd
        # 1. flat gradient all-reduce: sum in place + 1/world factor
        g = torch.full((1000,), float(rank + 1))
        f = bd.allreduce_mean_(g)
        assert f == 1.0 / world and torch.allclose(g * f, torch.full((1000,), (1 + world) / 2.0))
        # 2. contiguous patch ranges cover every patch exactly once, for any world size
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                a, b = bd.deal_patch_range(27, r, w)
                seen += list(range(a, b))
            assert seen == list(range(27))
        # 3. sharded sliding-window merge: after the slab exchange the output slab a rank owns, merged from ITS array alone,
        #    equals the same planes of the single-process merge bit for bit (patch pieces it never received stay NaN and
        #    would poison the slab if the plan missed one).  Grid: 40 x 20 x 24 volume, 16^3 patches, 25 % overlap, padding 2.
        import numpy as np
        from biapy_b200.data import _stitch
        from oracle import port_stitch
        vshape, patch, ov, pad = (40, 20, 24, 1), (16, 16, 16), (0.25, 0.25, 0.25), (2, 0, 0)
        axes = [_stitch.Axis(vshape[i], patch[i], pad[i], ov[i]) for i in range(3)]
        starts_m = [a.starts(1) for a in axes]
        n = len(starts_m[0]) * len(starts_m[1]) * len(starts_m[2])
        truth = torch.from_numpy(np.random.default_rng(0).standard_normal((n,) + patch + (1,)).astype(np.float32))
        ref = port_stitch.merge_3d(truth.numpy(), vshape, ov, pad)
        first, end = bd.deal_patch_range(n, rank, world)
        mine = torch.full_like(truth, float("nan"))
        mine[first:end] = truth[first:end]
        plan = bd.plan_slab_exchange(starts_m[0], len(starts_m[1]) * len(starts_m[2]), axes[0].core, pad[0], vshape[0], world)
        got = bd.exchange_patch_slabs(mine, plan, rank)
        assert got == sum((a1 - a0) * 16 * 16 * 4 for c, src, dst, a0, a1 in plan if dst == rank)
        assert got < (n - (end - first)) * 16 ** 3 * 4                      # less than an all-gather would deliver
        z0, z1 = bd.slab_range(vshape[0], rank, world)
        with np.errstate(invalid="ignore"):
            merged = port_stitch.merge_3d(mine.numpy(), vshape, ov, pad)
        assert np.array_equal(merged[z0:z1], ref[z0:z1]) and not np.isnan(merged[z0:z1]).any()
        # 4. by-chunks tile dealing == torch's DistributedSampler(shuffle=False); disjoint writes + one all-reduce = full volume
        from torch.utils.data import DistributedSampler
        for n_tiles in (1, 2, 7, 48):
            mine_t = bd.deal_tiles(n_tiles, rank, world)
            assert mine_t == list(DistributedSampler(list(range(n_tiles)), num_replicas=world, rank=rank, shuffle=False))
            own = torch.zeros(n_tiles)
            own[bd.deal_tiles(n_tiles, rank, world, drop_repeats=True)] = 1
            dist.all_reduce(own)
            assert torch.equal(own, torch.ones(n_tiles))      # without the repeats: exactly one owner per tile
        # 5. the epoch loop averages the loss over the ranks like MetricLogger.synchronize_between_processes
        #    (train_engine.py:204-206): sum of values and counts, not a mean of per-rank means
        import contextlib
        import io
        from biapy_b200.config.config import load_config
        from biapy_b200.engine.train_engine import train_one_epoch

        class FakeTrainer:
            def __init__(self, losses):
                self.param_groups, self.losses, self.k = [{"lr": 1e-3}], losses, 0

            def zero_grad(self):
                pass

            def step(self, batch, targets):
                self.k += 1
                return torch.tensor([self.losses[self.k - 1]], dtype=torch.float64)

        class FakeModel:
            def train(self, flag=True):
                pass

            def eval(self):
                pass

        cfg = load_config({"PROBLEM": {"NDIM": "2D"}, "DATA": {"PATCH_SIZE": (8, 8, 1)}})
        losses = [1.0, 2.0, 3.0] if rank == 0 else [10.0]                # ragged: 3 batches on rank 0, 1 on rank 1
        loader = [(torch.zeros(2, 8, 8, 1), torch.zeros(2, 8, 8, 1)) for _ in losses]
        with contextlib.redirect_stdout(io.StringIO()):
            stats, _ = train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, loader, [FakeTrainer(losses)], "cpu", 0)
        assert abs(stats["loss"] - 16.0 / 4) < 1e-12, stats
        # 6. ... and so does the validation pass (reference train_engine.py:318-321): every rank must hand the SAME loss to
        #    ReduceLROnPlateau / EarlyStopping / the best-checkpoint test
        from biapy_b200.engine.train_engine import evaluate

        class FakeEval(FakeTrainer):
            device = None

            def evaluate(self, images, targets):
                return self.step(images, targets)

        with contextlib.redirect_stdout(io.StringIO()):
            vstats = evaluate(cfg, FakeModel(), None, None, None, lambda t, b: t, 0, loader, optimizer=[FakeEval(losses)])
        assert abs(vstats["loss"] - 16.0 / 4) < 1e-12, vstats
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

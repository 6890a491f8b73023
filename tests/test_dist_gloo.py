"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: gradient all-reduce factor, patch dealing and the
prediction all-gather used by sliding-window inference, cross-rank loss statistics of the epoch loop."""
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
        # 2. patch dealing covers every patch exactly once
        n = 27
        mine = bd.deal_patches(n, rank, world)
        counts = torch.zeros(n)
        counts[mine] = 1
        dist.all_reduce(counts)
        assert torch.equal(counts, torch.ones(n))
        # 3. prediction gather: every rank ends with all rows
        pred = torch.zeros(n, 2, 3)
        for i in mine:
            pred[i] = float(i + 1)
        out = bd.gather_patch_predictions(pred, n)
        expect = torch.arange(1, n + 1, dtype=torch.float32).view(n, 1, 1).expand(n, 2, 3)
        assert torch.equal(out, expect)
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

        cfg = load_config({"PROBLEM": {"NDIM": "2D"}, "DATA": {"PATCH_SIZE": (8, 8, 1)}})
        losses = [1.0, 2.0, 3.0] if rank == 0 else [10.0]                # ragged: 3 batches on rank 0, 1 on rank 1
        loader = [(torch.zeros(2, 8, 8, 1), torch.zeros(2, 8, 8, 1)) for _ in losses]
        with contextlib.redirect_stdout(io.StringIO()):
            stats, _ = train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, loader, [FakeTrainer(losses)], "cpu", 0)
        assert abs(stats["loss"] - 16.0 / 4) < 1e-12, stats
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

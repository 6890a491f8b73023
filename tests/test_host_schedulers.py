"""Host side of the training loop: LR schedulers, ``prepare_optimizer`` plumbing, ``train_one_epoch`` /
``evaluate`` control flow, early stopping.  No GPU: the schedulers are pinned to ``torch.optim.lr_scheduler`` step by step and
to golden sequences produced by the reference's own classes (``oracle/make_golden_schedulers.py``); the epoch loops run on a
stand-in Trainer that records what the loop asks of it."""
import json
import os

import pytest
import torch

from biapy_b200.config.config import load_config
from biapy_b200.engine.schedulers import (OneCycleLR, ReduceLROnPlateau, WarmUpCosineDecayScheduler,
                                           WarmUpReduceOnPlateauScheduler)
from biapy_b200.engine.train_engine import evaluate, train_one_epoch
from biapy_b200.utils.callbacks import EarlyStopping

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "schedulers.json")


class Groups:
    """Anything with ``param_groups`` (the Trainer's surface)."""

    def __init__(self, **kw):
        self.param_groups = [dict(kw)]


def _torch_opt(kind, lr):
    p = [torch.nn.Parameter(torch.zeros(3))]
    if kind == "adamw":
        return torch.optim.AdamW(p, lr=lr, betas=(0.9, 0.999))
    return torch.optim.SGD(p, lr=lr, momentum=0.0)


@pytest.mark.parametrize("kind", ["adamw", "sgd"])
@pytest.mark.parametrize("epochs,spe,max_lr", [(3, 7, 1e-3), (10, 1, 5e-2), (2, 50, 1e-4)])
def test_onecycle_matches_torch(kind, epochs, spe, max_lr):
    ref_opt = _torch_opt(kind, 1e-4)
    ref = torch.optim.lr_scheduler.OneCycleLR(ref_opt, max_lr, epochs=epochs, steps_per_epoch=spe)
    mine_opt = Groups(lr=1e-4, betas=(0.9, 0.999)) if kind == "adamw" else Groups(lr=1e-4, momentum=0.0)
    mine = OneCycleLR(mine_opt, max_lr, epochs=epochs, steps_per_epoch=spe)
    for i in range(epochs * spe):
        g, r = mine_opt.param_groups[0], ref_opt.param_groups[0]
        assert g["lr"] == pytest.approx(r["lr"], rel=1e-12, abs=0), i
        if kind == "adamw":
            assert g["betas"][0] == pytest.approx(r["betas"][0], rel=1e-12) and g["betas"][1] == r["betas"][1]
        else:
            assert g["momentum"] == pytest.approx(r["momentum"], rel=1e-12)
        assert mine.get_last_lr() == pytest.approx(ref.get_last_lr(), rel=1e-12)
        if i < epochs * spe - 1:
            ref_opt.step()
            ref.step()
            mine.step()
    mine.step()                      # torch allows total_steps calls in all, then raises
    with pytest.raises(ValueError):
        mine.step()


@pytest.mark.parametrize("patience,factor,min_lr", [(2, 0.5, 1e-6), (0, 0.1, 0.0), (5, 0.5, 2e-4)])
def test_reduce_on_plateau_matches_torch(patience, factor, min_lr):
    ref_opt = _torch_opt("adamw", 1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(ref_opt, patience=patience, factor=factor, min_lr=min_lr)
    mine_opt = Groups(lr=1e-3)
    mine = ReduceLROnPlateau(mine_opt, patience=patience, factor=factor, min_lr=min_lr)
    g = torch.Generator().manual_seed(patience * 7 + 1)
    loss = 1.0
    for epoch in range(60):
        loss = loss * (0.97 if epoch < 8 else 1.0) + 0.01 * float(torch.rand(1, generator=g)) * (epoch % 3 == 0)
        ref.step(loss)
        mine.step(loss, epoch=epoch)
        assert mine_opt.param_groups[0]["lr"] == pytest.approx(ref_opt.param_groups[0]["lr"], rel=1e-12), epoch
        assert mine.num_bad_epochs == ref.num_bad_epochs and mine.best == pytest.approx(ref.best)
    assert mine_opt.param_groups[0]["lr"] < 1e-3       # the sequence did reduce the rate


def test_warmup_schedules_match_reference_golden():
    with open(GOLDEN) as f:
        gold = json.load(f)

    class Opt:
        def __init__(self):
            self.param_groups = [{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}]

    for case in gold["warmupcosine"]:
        s, o = WarmUpCosineDecayScheduler(case["lr"], case["min_lr"], case["warmup_epochs"], case["epochs"]), Opt()
        k = 0
        for e in range(case["epochs"]):
            for st in range(case["steps_per_epoch"]):
                r = s.adjust_learning_rate(o, st / case["steps_per_epoch"] + e)
                assert [r, o.param_groups[0]["lr"], o.param_groups[1]["lr"]] == case["seq"][k]      # bit-exact doubles
                k += 1
    for case in gold["warmupreduceonplateau"]:
        s, o = WarmUpReduceOnPlateauScheduler(case["lr"], case["epochs"]), Opt()
        assert [float(v) for v in s.LR] == case["table"]
        for e, want in enumerate(case["seq"]):
            r = s.adjust_learning_rate(o, e + 0.5)
            assert [r, o.param_groups[0]["lr"], o.param_groups[1]["lr"]] == want


def test_early_stopping_counts_epochs_without_improvement():
    msgs = []
    es = EarlyStopping(patience=3, trace_func=msgs.append)
    for v in [1.0, 0.9, 0.95, 0.91, 0.8]:
        es(v)
        assert not es.early_stop
    assert es.counter == 0 and es.val_loss_min == 0.8
    es(0.8)                                   # an equal loss is not 'worse' (strict <): the counter stays at 0
    assert es.counter == 0
    for v in [0.81, 0.85, 0.82]:
        es(v)
    assert es.early_stop and len(msgs) == 5 and msgs[-1] == "EarlyStopping counter: 3 out of 3"


# ------------------------------------------------------------------------------------------- the epoch loops
class FakeTrainer:
    """Records the learning rate of every update; the loss is a scripted sequence."""

    def __init__(self, losses, lr=1e-3):
        self.param_groups = [{"lr": lr, "betas": (0.9, 0.999), "weight_decay": 0.0}]
        self.losses = list(losses)
        self.lrs, self.betas1, self.eval_calls = [], [], 0
        self.zeroed = 0

    def zero_grad(self):
        self.zeroed += 1

    def step(self, batch, targets):
        self.lrs.append(self.param_groups[0]["lr"])
        self.betas1.append(self.param_groups[0]["betas"][0])
        return torch.tensor([self.losses[len(self.lrs) - 1]], dtype=torch.float64)

    def evaluate(self, batch, targets):
        self.eval_calls += 1
        return torch.tensor([self.losses[self.eval_calls - 1]], dtype=torch.float64)


class FakeModel:
    def __init__(self):
        self.mode = None

    def train(self, flag=True):
        self.mode = "train" if flag else "eval"

    def eval(self):
        self.mode = "eval"


def _cfg(name, **sched):
    return load_config({"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": (8, 8, 8, 1), "TEST": {"OVERLAP": (0, 0, 0), "PADDING": (0, 0, 0)}},
                        "TRAIN": {"EPOCHS": 4, "LR": [1e-3], "LR_SCHEDULER": dict(NAME=name, **sched)}})


def _loader(n, shape=(2, 8, 8, 8, 1)):
    return [(torch.zeros(shape), torch.zeros(shape)) for _ in range(n)]


def test_train_one_epoch_warmupcosine_adjusts_every_iteration():
    cfg = _cfg("warmupcosine", MIN_LR=[1e-5], WARMUP_COSINE_DECAY_EPOCHS=2)
    tr, model = FakeTrainer([0.7, 0.6, 0.5, 0.4, 0.3]), FakeModel()
    sched = WarmUpCosineDecayScheduler(1e-3, 1e-5, 2, 4)
    seen = []
    stats, step = train_one_epoch(cfg, model, None, None, None, lambda t, b: seen.append(1) or t, _loader(5), [tr], "cpu", epoch=1,
                                  lr_scheduler=[sched], loss_names=["loss"])
    assert model.mode == "train" and tr.zeroed == 1 and len(seen) == 5 and step == 4
    want = [1e-3 * (1 + k / 5) / 2 for k in range(5)]           # epoch 1 + k/5 of a 2-epoch warm-up
    assert tr.lrs == pytest.approx(want, rel=1e-12)
    assert stats["loss"] == pytest.approx(0.5) and stats["lr"] == pytest.approx(want[-1])


def test_train_one_epoch_onecycle_steps_after_every_update():
    cfg = _cfg("onecycle")
    tr = FakeTrainer([1.0] * 6)
    sched = OneCycleLR(tr, 1e-3, epochs=2, steps_per_epoch=3)
    ref_opt = _torch_opt("adamw", 1e-3)
    ref = torch.optim.lr_scheduler.OneCycleLR(ref_opt, 1e-3, epochs=2, steps_per_epoch=3)
    want_lr, want_b1 = [], []
    for _ in range(6):
        want_lr.append(ref_opt.param_groups[0]["lr"])
        want_b1.append(ref_opt.param_groups[0]["betas"][0])
        ref_opt.step()
        if len(want_lr) < 6:
            ref.step()
    for epoch in range(2):
        train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, _loader(3), [tr], "cpu", epoch=epoch, lr_scheduler=[sched])
    assert tr.lrs == pytest.approx(want_lr, rel=1e-12) and tr.betas1 == pytest.approx(want_b1, rel=1e-12)


def test_train_one_epoch_rejects_wrong_patch_shape_and_non_finite_loss():
    cfg = _cfg("")
    with pytest.raises(ValueError, match="Trying to input data with different shape than 'DATA.PATCH_SIZE'"):
        train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, _loader(2, (2, 8, 8, 4, 1)), [FakeTrainer([1, 1])], "cpu", 0)
    tr = FakeTrainer([0.5, float("nan"), 0.4, 0.3])
    with pytest.raises(SystemExit) as e:
        train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, _loader(4), [tr], "cpu", 0)
    assert e.value.code == 1 and len(tr.lrs) == 3            # detected with a lag of one iteration, as documented
    with pytest.raises(NotImplementedError):
        train_one_epoch(cfg, FakeModel(), None, None, None, lambda t, b: t, _loader(1), [tr], "cpu", 0, memory_bank=object())


def test_evaluate_averages_and_drives_reduce_on_plateau():
    cfg = _cfg("reduceonplateau", MIN_LR=[1e-6], REDUCEONPLATEAU_PATIENCE=0, REDUCEONPLATEAU_FACTOR=0.5)
    tr, model = FakeTrainer([0.2, 0.4, 0.6]), FakeModel()
    sched = ReduceLROnPlateau(tr, patience=0, factor=0.5, min_lr=1e-6)
    stats = evaluate(cfg, model, None, None, None, lambda t, b: t, 0, _loader(3), lr_scheduler=[sched], optimizer=[tr])
    assert model.mode == "eval" and tr.eval_calls == 3 and stats["loss"] == pytest.approx(0.4)
    assert tr.param_groups[0]["lr"] == 1e-3                   # first value is the best so far
    tr.eval_calls = 0
    evaluate(cfg, model, None, None, None, lambda t, b: t, 1, _loader(3), lr_scheduler=[sched], optimizer=[tr])
    assert tr.param_groups[0]["lr"] == pytest.approx(5e-4)     # no improvement with patience 0 -> halved
    with pytest.raises(ValueError):
        evaluate(cfg, model, None, None, None, lambda t, b: t, 0, _loader(1))


def test_config_carries_the_scheduler_keys_with_reference_defaults():
    c = load_config(None)
    s = c.TRAIN.LR_SCHEDULER
    assert (s.NAME, s.MIN_LR, s.REDUCEONPLATEAU_FACTOR, s.REDUCEONPLATEAU_PATIENCE, s.WARMUP_COSINE_DECAY_EPOCHS) == ("", [-1.0], 0.5, -1, -1)
    assert c.TRAIN.PATIENCE == -1 and c.TRAIN.EPOCHS == 360 and c.MODEL.SAVE_CKPT_FREQ == -1
    c = load_config("TRAIN:\n  LR_SCHEDULER:\n    NAME: onecycle\n")
    assert c.TRAIN.LR_SCHEDULER.NAME == "onecycle" and c.TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_FACTOR == 0.5


def test_scheduler_configuration_rules_of_check_configuration():
    """``prepare_optimizer`` refuses what ``check_configuration.py:3304-3351`` refuses, with its messages, before it touches the
    model (so this runs without a GPU)."""
    from biapy_b200.engine import check_lr_scheduler, prepare_optimizer

    def cfg(**sched):
        patience = sched.pop("PATIENCE", -1)
        return load_config({"TRAIN": {"EPOCHS": 10, "PATIENCE": patience, "LR_SCHEDULER": sched}})
    for c in (cfg(NAME=""), cfg(NAME="onecycle"), cfg(NAME="warmupreduceonplateau"),
              cfg(NAME="warmupcosine", MIN_LR=[1e-5], WARMUP_COSINE_DECAY_EPOCHS=2),
              cfg(NAME="reduceonplateau", MIN_LR=1e-6, REDUCEONPLATEAU_PATIENCE=2, PATIENCE=5)):
        check_lr_scheduler(c)
    bad = [(cfg(NAME="cosine"), "'TRAIN.LR_SCHEDULER.NAME' must be in"),
           (cfg(NAME="warmupcosine", WARMUP_COSINE_DECAY_EPOCHS=2), "'TRAIN.LR_SCHEDULER.MIN_LR' needs to be set"),
           (cfg(NAME="reduceonplateau", MIN_LR=[1e-6]), "'TRAIN.LR_SCHEDULER.REDUCEONPLATEAU_PATIENCE' needs to be set"),
           (cfg(NAME="reduceonplateau", MIN_LR=[1e-6], REDUCEONPLATEAU_PATIENCE=5, PATIENCE=5), "needs to be less than 'TRAIN.PATIENCE'"),
           (cfg(NAME="warmupcosine", MIN_LR=[1e-5]), "'TRAIN.LR_SCHEDULER.WARMUP_COSINE_DECAY_EPOCHS' needs to be set"),
           (cfg(NAME="warmupcosine", MIN_LR=[1e-5], WARMUP_COSINE_DECAY_EPOCHS=11), "needs to be less than 'TRAIN.EPOCHS'")]
    for c, msg in bad:
        with pytest.raises(ValueError, match=msg):
            prepare_optimizer(c, model_without_ddp=None, steps_per_epoch=3)


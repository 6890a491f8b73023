"""Learning-rate schedules of the training loop (``biapy/engine/schedulers/`` + the two ``torch.optim.lr_scheduler`` classes
``prepare_optimizer`` builds, ``biapy/engine/__init__.py:74-104``).  They are host arithmetic on ``param_groups`` -- the
:class:`biapy_b200.engine.train.Trainer` plays the optimiser's role and exposes the same list of dicts -- so they run the same
with a ``torch.optim`` optimiser, which is how ``tests/test_host_schedulers.py`` pins them to torch / the reference."""
from .lr_schedulers import OneCycleLR, ReduceLROnPlateau, WarmUpCosineDecayScheduler, WarmUpReduceOnPlateauScheduler

__all__ = ["OneCycleLR", "ReduceLROnPlateau", "WarmUpCosineDecayScheduler", "WarmUpReduceOnPlateauScheduler"]

"""Import location of the reference (``biapy.engine.schedulers.warmup_reduce_on_plateau``); the class lives in ``lr_schedulers``."""
from .lr_schedulers import WarmUpReduceOnPlateauScheduler  # noqa: F401

"""Import location of the reference (``biapy.engine.schedulers.warmup_cosine_decay``); the class lives in ``lr_schedulers``."""
from .lr_schedulers import WarmUpCosineDecayScheduler  # noqa: F401

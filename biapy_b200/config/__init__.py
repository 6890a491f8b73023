from .config import Config, load_config  # noqa: F401

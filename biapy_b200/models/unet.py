"""``U_Net`` with the constructor, state_dict layout and forward contract of ``biapy/models/unet.py:29-445``,
executed by hand-written sm_100a kernels (see ``biapy_b200/models/_base.py``)."""
from biapy_b200.models._base import UNetFamily


class U_Net(UNetFamily):
    """2D/3D U-Net: encoder ``ConvBlock`` levels, ``UpBlock`` decoder (reference ``unet.py:36-62`` for the kwargs)."""

    variant = "unet"

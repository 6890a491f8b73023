"""``ResUNet`` with the constructor, state_dict layout and forward contract of ``biapy/models/resunet.py:27-446``,
executed by hand-written sm_100a kernels (see ``biapy_b200/models/_base.py``)."""
from biapy_b200.models._base import UNetFamily


class ResUNet(UNetFamily):
    """2D/3D Residual U-Net: ``ResConvBlock`` encoder, ``ResUpBlock`` decoder (reference ``resunet.py:34-60``)."""

    variant = "resunet"

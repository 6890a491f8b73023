"""``Attention_U_Net`` with the constructor, state_dict layout and forward contract of
``biapy/models/attention_unet.py:34-459``, executed by hand-written sm_100a kernels."""
from biapy_b200.models._base import UNetFamily


class Attention_U_Net(UNetFamily):
    """U-Net whose skip connections pass through attention gates (reference ``attention_unet.py:313``)."""

    variant = "attention_unet"

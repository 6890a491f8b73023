"""Denoising (Noise2Void) workflow on the B200 engine (``biapy/engine/denoising.py``)."""
from __future__ import annotations

from ..data.norm import undo_image_norm
from .base_workflow import Base_Workflow


class Denoising_Workflow(Base_Workflow):
    loss_kind = "n2v_mse"

    def define_activations_and_channels(self):
        """As many linear output channels as the image has (reference ``:117-121``)."""
        self.model_output_channels = [int(self.cfg.DATA.PATCH_SIZE[-1])]
        self.gt_channels_expected = self.model_output_channels[0]
        self.separated_class_channel = False
        self.head_activations = ["linear"] * self.model_output_channels[0]
        self.model_output_channel_info = ["pred{}".format(i) for i in range(len(self.model_output_channels))]
        super().define_activations_and_channels()

    def after_merge_patches(self, pred):
        """Back to the image's own intensity range and dtype (reference ``:413``)."""
        info = self.current_sample.get("norm_info")
        return undo_image_norm(pred, info) if info is not None else pred

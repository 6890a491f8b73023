"""Semantic-segmentation workflow on the B200 engine (``biapy/engine/semantic_seg.py``)."""
from __future__ import annotations

from ..data.norm import binarize_prediction
from .base_workflow import Base_Workflow


class Semantic_Segmentation_Workflow(Base_Workflow):
    def define_activations_and_channels(self):
        """One sigmoid channel for binary problems, ``N_CLASSES`` softmax channels otherwise (reference ``:98-135``)."""
        n = int(self.cfg.DATA.N_CLASSES)
        self.model_output_channels = [1 if n <= 2 else n]
        self.gt_channels_expected = n
        self.separated_class_channel = False
        self.head_activations = ["ce_softmax" if n > 2 else "ce_sigmoid"] * self.model_output_channels[0]
        self.model_output_channel_info = ["pred{}".format(i) for i in range(len(self.model_output_channels))]
        self.loss_kind = "ce" if n > 2 else "bce"
        super().define_activations_and_channels()

    def after_merge_patches(self, pred, threshold: float = 0.5):
        """Binarised prediction (reference ``:409-425``: Otsu threshold for the whole image; the by-chunks path and this
        engine use the fixed 0.5 of ``:524-531`` unless a threshold is passed) -- uint8 mask or arg-max class map."""
        return binarize_prediction(pred, int(self.cfg.DATA.N_CLASSES), threshold=threshold)

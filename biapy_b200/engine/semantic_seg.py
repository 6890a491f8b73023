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
        # CrossEntropyLoss_wrapper (reference metrics.py:493-586, built at semantic_seg.py:193-200): LOSS.IGNORE_INDEX -1 means
        # torch's default -100; class re-balancing (a weight tensor in the loss) is not in the fused loss kernels
        # ('manual' is the only value the wrapper acts on, :540); refused when a Trainer is built, inference is unaffected
        loss_cfg = self.cfg.get("LOSS", {}) if hasattr(self.cfg, "get") else {}
        self.unsupported_loss_options = None
        if str(loss_cfg.get("CLASS_REBALANCE", "none")).lower() == "manual":
            self.unsupported_loss_options = ("LOSS.CLASS_REBALANCE = 'manual' (class weights in the loss) is not implemented by the "
                                             "B200 loss kernels: un-weighted BCEWithLogits / CrossEntropy only")
        ii = int(loss_cfg.get("IGNORE_INDEX", -1))
        self.ignore_index = -100 if ii == -1 else ii
        super().define_activations_and_channels()

    def after_merge_patches(self, pred, threshold=None):
        """Binarised prediction (reference ``:418-425``, ``after_full_image`` ``:444-459``): ``pred > threshold_otsu(pred)`` for
        binary problems -- the Otsu threshold of the whole merged prediction, histogram on the device -- or the arg-max class map;
        a number as `threshold` replaces Otsu."""
        return binarize_prediction(pred, int(self.cfg.DATA.N_CLASSES), threshold=threshold)

    def after_one_chunk_workflow_process(self, chunks, patch_in_data=None, added_pad=None):
        """By-chunks binarisation (reference ``:502-535``): fixed 0.5, NOT Otsu -- a chunk may hold no foreground at all."""
        return [binarize_prediction(c, int(self.cfg.DATA.N_CLASSES), threshold=0.5) for c in chunks]

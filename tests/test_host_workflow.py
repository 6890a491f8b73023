"""Host logic of the workflow plugin surface (biapy_b200.config, biapy_b200.engine.{base_workflow, semantic_seg, denoising}):
configuration defaults / merging and the hooks a BiaPy workflow defines.  No GPU needed."""
import pytest

YAML_3D = """
# the reference's 3D semantic-segmentation template, hot-path keys (templates/semantic_segmentation/3d_semantic_segmentation.yaml)
SYSTEM:
    NUM_CPUS: -1
PROBLEM:
    TYPE: SEMANTIC_SEG
    NDIM: 3D
DATA:
    PATCH_SIZE: (80, 80, 80, 1)
    TEST:
        PADDING: (10,10,10)
MODEL:
    ARCHITECTURE: resunet
    FEATURE_MAPS: [16, 32, 64, 128, 256]
    LOAD_CHECKPOINT: False
TRAIN:
    ENABLE: True
    OPTIMIZER: ADAMW
    LR: 1.E-3
    BATCH_SIZE: 4
TEST:
    ENABLE: True
    AUGMENTATION: False
"""


def test_config_defaults_and_yaml_merge(tmp_path):
    from biapy_b200.config import load_config
    from biapy_b200.config.config import first
    d = load_config()
    assert d.PROBLEM.TYPE == "SEMANTIC_SEG" and d.PROBLEM.NDIM == "2D" and d.DATA.PATCH_SIZE == (256, 256, 1)      # config.py:83-85, 797
    assert d.MODEL.NORMALIZATION == "in" and d.MODEL.ACTIVATION == "elu" and d.TRAIN.W_DECAY == 0.02                # :1520, 1529, 1968
    c = load_config(YAML_3D)
    assert c.DATA.PATCH_SIZE == (80, 80, 80, 1) and c.DATA.TEST.PADDING == (10, 10, 10) and c.DATA.TEST.OVERLAP == (0, 0, 0)
    assert c.MODEL.ARCHITECTURE == "resunet" and c.SYSTEM.NUM_CPUS == -1 and c.SYSTEM.SEED == 0
    assert first(c.TRAIN.OPTIMIZER) == "ADAMW" and abs(float(first(c.TRAIN.LR)) - 1e-3) < 1e-12
    assert first(d.TRAIN.OPTIMIZER) == "SGD" and first(d.TRAIN.OPT_BETAS) == [0.9, 0.999]
    f = tmp_path / "cfg.yaml"
    f.write_text(YAML_3D)
    assert load_config(str(f)) == c
    assert load_config(dict(c)).DATA.TEST.PADDING == (10, 10, 10)
    with pytest.raises(ValueError):
        load_config({"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": (64, 64, 1)}})
    with pytest.raises(FileNotFoundError):
        load_config("/no/such/file.yaml")


def test_workflow_hooks_and_model_kwargs():
    import contextlib
    import io
    from biapy_b200.config import load_config
    from biapy_b200.engine.denoising import Denoising_Workflow
    from biapy_b200.engine.semantic_seg import Semantic_Segmentation_Workflow
    c = load_config({"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": (32, 32, 32, 2)}, "MODEL": {"ARCHITECTURE": "resunet",
                     "FEATURE_MAPS": [8, 16], "DROPOUT_VALUES": [0, 0], "ISOTROPY": [True, True], "CONV_LAYERS": [2, 2], "Z_DOWN": [0], "YX_DOWN": [0]}})
    w = Semantic_Segmentation_Workflow(c, "job_1", "cpu", {}, None)
    assert w.model_output_channels == [1] and w.head_activations == ["ce_sigmoid"] and w.loss_kind == "bce"
    assert w.axes_order == (0, 4, 1, 2, 3) and w.axes_order_back == (0, 2, 3, 4, 1)
    assert w.norm_module["type"] == "zero_mean_unit_variance" and w.norm_module["mean"] == [-1.0] and w.test_norm_module["out_dtype"] == "float32"
    with contextlib.redirect_stdout(io.StringIO()):
        m = w.prepare_model()
    assert type(m).__name__ == "ResUNet" and w.network_stride == [1, 1, 1]
    assert "bottleneck.block.0.weight" in m.state_dict()                            # pre-norm of a residual level: reference key layout
    c3 = load_config({"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": (32, 32, 32, 1), "N_CLASSES": 4}})
    w3 = Semantic_Segmentation_Workflow(c3, "j", "cpu")
    assert w3.model_output_channels == [4] and w3.head_activations == ["ce_softmax"] * 4 and w3.loss_kind == "ce"
    d = Denoising_Workflow(load_config({"PROBLEM": {"TYPE": "DENOISING", "NDIM": "2D"}, "DATA": {"PATCH_SIZE": (64, 64, 3)}}), "j", "cpu")
    assert d.model_output_channels == [3] and d.head_activations == ["linear"] * 3 and d.loss_kind == "n2v_mse"
    assert d.axes_order == (0, 3, 1, 2)

    class Broken(Semantic_Segmentation_Workflow):
        def define_activations_and_channels(self):
            self.model_output_channels, self.head_activations = [2], ["linear"]
            super(Semantic_Segmentation_Workflow, self).define_activations_and_channels()
    with pytest.raises(ValueError):
        Broken(c, "j", "cpu")


def test_build_config_like_the_reference_api_check():
    """``build_config`` as ``tests/check_api.py:89-103`` of the reference calls it: high-level arguments -> override dict that
    loads as a configuration; same validation messages (``_biapy.py:2043-2053``)."""
    from biapy_b200._biapy import VALID_WORKFLOWS, build_config
    from biapy_b200.config.config import load_config
    cfg = build_config(workflow="semantic_seg", dims="3d", phase="both", patch_size=[32, 32, 32, 1], model={"architecture": "resunet"},
                       train_data={"path": "x", "gt_path": "y", "in_memory": True}, val_data={"split_train": 0.1},
                       test_data={"path": "tx", "in_memory": False, "overlap": (0.25, 0.25, 0.25)},
                       extra_config={"TRAIN": {"EPOCHS": 10, "PATIENCE": -1}, "MODEL": {"FEATURE_MAPS": [16, 32]}})
    assert cfg == {
        "PROBLEM": {"TYPE": "SEMANTIC_SEG", "NDIM": "3D"},
        "TRAIN": {"ENABLE": True, "EPOCHS": 10, "PATIENCE": -1},
        "TEST": {"ENABLE": True},
        "DATA": {"PATCH_SIZE": (32, 32, 32, 1), "TRAIN": {"PATH": "x", "GT_PATH": "y", "IN_MEMORY": True}, "VAL": {"SPLIT_TRAIN": 0.1},
                 "TEST": {"PATH": "tx", "IN_MEMORY": False, "OVERLAP": (0.25, 0.25, 0.25)}},
        "MODEL": {"ARCHITECTURE": "resunet", "FEATURE_MAPS": [16, 32]},
    }
    c = load_config(cfg)
    assert c.PROBLEM.NDIM == "3D" and c.DATA.TEST.OVERLAP == (0.25, 0.25, 0.25) and c.DATA.TEST.PADDING == (0, 0, 0)
    assert c.TRAIN.EPOCHS == 10 and c.MODEL.ARCHITECTURE == "resunet" and c.MODEL.KERNEL_SIZE == 3
    only_test = build_config("DENOISING", "2D", phase="test")
    assert only_test == {"PROBLEM": {"TYPE": "DENOISING", "NDIM": "2D"}, "TRAIN": {"ENABLE": False}, "TEST": {"ENABLE": True}}
    with pytest.raises(ValueError, match="'workflow' must be one of"):
        build_config("SEGMENT_ANYTHING", "2D")
    with pytest.raises(ValueError, match="'dims' must be either '2D' or '3D'. Provided: 4D"):
        build_config("SEMANTIC_SEG", "4d")
    with pytest.raises(ValueError, match="'phase' must be one of"):
        build_config("SEMANTIC_SEG", "2D", phase="validate")
    assert len(VALID_WORKFLOWS) == 8

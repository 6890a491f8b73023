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


def test_full_image_inference_pads_calls_the_model_once_and_crops():
    """``TEST.FULL_IMG`` (2D, reference ``base_workflow.py:2224-2290``): zero padding to a multiple of 2**levels at the bottom /
    right (``check_downsample_division``, ``util.py:637-674``), ONE model call on the whole image, crop back, ``after_full_image``.
    The model call is a stand-in here (the real one is the same ``model_call_func`` the patch path uses, covered on the GPU)."""
    import numpy as np
    import torch
    from biapy_b200.config import load_config
    from biapy_b200.engine.denoising import Denoising_Workflow
    from biapy_b200.utils.util import check_downsample_division
    x = np.arange(2 * 5 * 9 * 1, dtype=np.float32).reshape(2, 5, 9, 1)
    xp, o = check_downsample_division(x, 3)
    ref = np.pad(x, ((0, 0), (0, 3), (0, 7), (0, 0)))
    assert o == (2, 5, 9, 1) and xp.shape == (2, 8, 16, 1) and np.array_equal(xp, ref)
    xt, ot = check_downsample_division(torch.from_numpy(x), 3)
    assert ot == o and torch.equal(xt, torch.from_numpy(ref))
    same, _ = check_downsample_division(ref, 3)
    assert same is ref                                                   # already divisible: untouched

    cfg = load_config({"PROBLEM": {"TYPE": "DENOISING", "NDIM": "2D"}, "DATA": {"PATCH_SIZE": (16, 16, 2)}, "TEST": {"FULL_IMG": True},
                       "MODEL": {"FEATURE_MAPS": [8, 16, 32]}})
    w = Denoising_Workflow(cfg, "j", "cpu")
    calls = []

    class Model:
        def eval(self):
            calls.append("eval")
    w.model = Model()

    def fake_model_call(in_img, is_train=False, apply_act=True):
        calls.append(tuple(in_img.shape))
        t = torch.from_numpy(in_img) if isinstance(in_img, np.ndarray) else in_img
        return (t * 2 + 1).permute(0, 3, 1, 2)                            # (N, C, Y, X) like the engine
    w.model_call_func = fake_model_call
    img = np.random.default_rng(0).standard_normal((13, 22, 2)).astype(np.float32)
    pred, post = w.process_test_sample(img, norm=False)
    assert calls == ["eval", (1, 16, 24, 2)]                              # padded to multiples of 2**2, one call
    assert isinstance(pred, np.ndarray) and pred.shape == (13, 22, 2) and np.array_equal(pred, img * 2 + 1)
    assert np.array_equal(post, pred)                                      # no norm_info recorded: nothing to undo
    pred_t, _ = w.process_test_sample(torch.from_numpy(img), norm=False)
    assert isinstance(pred_t, torch.Tensor) and torch.equal(pred_t, torch.from_numpy(img * 2 + 1))
    assert load_config(None).TEST.FULL_IMG is False                         # reference default (config.py:2071)


@pytest.mark.reference
def test_reference_templates_build_the_same_networks():
    """The reference's own YAML templates for the two workflows of the hot path (``templates/semantic_segmentation``,
    ``templates/denoising``) load through ``load_config``, resolve to the workflow classes, and ``prepare_model`` builds a network
    with the same ``state_dict`` keys and shapes as the reference class constructed from the same keyword arguments (container
    only: needs /root/reference).  Templates of architectures outside the path (hrnet48) must say so."""
    import contextlib
    import glob
    import io
    from biapy_b200.config import load_config
    from biapy_b200.engine.denoising import Denoising_Workflow
    from biapy_b200.engine.semantic_seg import Semantic_Segmentation_Workflow
    from oracle import ref_loader
    R = ref_loader.load()
    ref_cls = {"U_Net": R.unet.U_Net, "ResUNet": R.resunet.ResUNet, "Attention_U_Net": R.attention_unet.Attention_U_Net}
    root = "/root/reference/templates"
    files = sorted(glob.glob(f"{root}/semantic_segmentation/*.yaml") + glob.glob(f"{root}/semantic_segmentation/*/*.yaml")
                   + glob.glob(f"{root}/denoising/*.yaml"))
    assert len(files) >= 7
    built = 0
    for f in files:
        c = load_config(f)
        cls = Denoising_Workflow if c.PROBLEM.TYPE == "DENOISING" else Semantic_Segmentation_Workflow
        w = cls(c, "job_1", "cpu")
        if str(c.MODEL.ARCHITECTURE).lower() not in ("unet", "resunet", "attention_unet"):
            with pytest.raises(NotImplementedError, match="outside the B200 hot path"):
                w.prepare_model()
            continue
        with contextlib.redirect_stdout(io.StringIO()):
            m = w.prepare_model()
            ref = ref_cls[type(m).__name__](**w.model_build_kwargs)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        theirs = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert ours == theirs, f
        assert c.TRAIN.LR_SCHEDULER.NAME in ("", "onecycle", "warmupcosine", "reduceonplateau", "warmupreduceonplateau")
        built += 1
    assert built >= 5


def test_model_lists_follow_the_feature_maps_like_check_configuration():
    """``MODEL.DROPOUT_VALUES`` / ``ISOTROPY`` / ``CONV_LAYERS`` / ``Z_DOWN`` / ``YX_DOWN`` are fitted to ``MODEL.FEATURE_MAPS`` with
    the rules of ``check_configuration.py:2677-2790``: defaults written for five levels also serve three or six levels, uniform
    lists are broadcast, anything else must match (same error messages)."""
    import contextlib
    import io
    from biapy_b200.config import load_config
    from biapy_b200.models import build_model, model_kwargs_from_cfg

    def kwargs(model, patch=(128, 128, 128, 1)):
        c = load_config({"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": patch}, "MODEL": model})
        return model_kwargs_from_cfg(c, [1], ["pred0"], ["ce_sigmoid"]), c
    k, c = kwargs({"FEATURE_MAPS": [4, 8, 16, 32, 64, 128]})                     # six levels on five-level defaults
    assert k["drop_values"] == [0.0] * 6 and k["isotropy"] == [True] * 6 and k["conv_layers"] == [2] * 6
    assert k["z_down"] == [2] * 5 and k["yx_down"] == [2] * 5
    with contextlib.redirect_stdout(io.StringIO()):
        m = build_model(c, [1], ["pred0"], ["ce_sigmoid"], "cpu")[0]
    assert len(m.down_path) == 5 and len(m.up_paths[0]) == 5
    k, _ = kwargs({"FEATURE_MAPS": [8, 16, 32], "CONV_LAYERS": [3]})
    assert k["conv_layers"] == [3, 3, 3] and k["drop_values"] == [0.0] * 3 and k["isotropy"] == [True] * 3
    k, _ = kwargs({"FEATURE_MAPS": [8, 16, 32], "CONV_LAYERS": [], "ISOTROPY": [False, True, True], "DROPOUT_VALUES": [0.1, 0.2, 0.3],
                   "Z_DOWN": [1, 2], "YX_DOWN": [2, 2]})
    assert k["conv_layers"] == [2, 2, 2] and k["isotropy"] == [False, True, True] and k["drop_values"] == [0.1, 0.2, 0.3]
    assert k["z_down"] == [1, 2]
    for model, msg in (({"FEATURE_MAPS": [8, 16, 32], "CONV_LAYERS": [2, 3]}, "'MODEL.FEATURE_MAPS' and 'MODEL.CONV_LAYERS' lengths must be equal"),
                       ({"FEATURE_MAPS": [8, 16, 32], "CONV_LAYERS": [0]}, "'MODEL.CONV_LAYERS' values must be greater than or equal to 1"),
                       ({"FEATURE_MAPS": [8, 16, 32], "DROPOUT_VALUES": [0.1, 0.2]}, "'MODEL.FEATURE_MAPS' and 'MODEL.DROPOUT_VALUES' lengths must be equal"),
                       ({"FEATURE_MAPS": [8, 16, 32], "DROPOUT_VALUES": [0.1, 1.2]}, "'MODEL.DROPOUT_VALUES' not in \\[0, 1\\] range"),
                       ({"FEATURE_MAPS": [8, 16, 32], "Z_DOWN": [2, 2, 2]}, "'MODEL.FEATURE_MAPS' length minus one and 'MODEL.Z_DOWN' length must be equal")):
        with pytest.raises(ValueError, match=msg):
            kwargs(model)
    # the patch must halve evenly at every level and keep more than two voxels (check_configuration.py:3183-3201)
    with pytest.raises(ValueError, match="not divisible by the downsampling factor at level 4 of the unet"):
        kwargs({"FEATURE_MAPS": [4, 8, 16, 32, 64, 128]}, patch=(32, 32, 32, 1))
    with pytest.raises(ValueError, match="not divisible by the downsampling factor at level 2 of the resunet"):
        kwargs({"ARCHITECTURE": "resunet", "FEATURE_MAPS": [8, 16, 32, 64]}, patch=(20, 36, 36, 1))      # 20 -> 10 -> 5: odd
    k, _ = kwargs({"FEATURE_MAPS": [8, 16, 32], "Z_DOWN": [1, 1]}, patch=(5, 36, 36, 1))       # Z not down-sampled: any depth in z
    assert k["z_down"] == [1, 1]


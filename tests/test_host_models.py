"""CPU tests of the host-side model mirror: constructor surface, state_dict layout (checkpoint compatibility with
the reference classes, via the golden fixtures) and loud failure without a GPU."""
import contextlib
import glob
import io
import json
import os

import numpy as np
import pytest
import torch

from biapy_b200 import _lib
from biapy_b200.models import build_model
from biapy_b200.models.attention_unet import Attention_U_Net
from biapy_b200.models.resunet import ResUNet
from biapy_b200.models.unet import U_Net

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLS = {"unet": U_Net, "resunet": ResUNet, "attention_unet": Attention_U_Net}


def _build(arch, kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return CLS[arch](**kw)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz"))))
def test_state_dict_matches_reference_layout(path):
    z = np.load(path)
    kw = json.loads(str(z["kwargs_json"]))
    m = _build(str(z["arch"]), kw)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    m.load_state_dict(sd, strict=True)          # same keys, same shapes as the reference class
    assert list(m.state_dict().keys()) == list(sd.keys())   # and the same order


def test_cfg2_model_has_reference_parameter_count():
    m = _build("resunet", dict(image_shape=(128, 128, 128, 2), activation="silu", feature_maps=[16, 32, 64, 128, 256],
                               drop_values=[0] * 5, normalization="gn", k_size=3, yx_down=[2] * 4, z_down=[2] * 4,
                               isotropy=[True] * 5, larger_io=False, conv_layers=[2] * 5, output_channels=[1]))
    assert sum(p.numel() for p in m.parameters()) == 6_694_065 or abs(sum(p.numel() for p in m.parameters()) - 6.694e6) < 2e3
    keys = list(m.state_dict().keys())
    assert keys[0] == "down_path.0.block.0.block.0.weight"
    assert "up_paths.0.3.conv_block.shortcut.0.weight" in keys and "heads.0.bias" in keys
    assert tuple(m.state_dict()["up_paths.0.0.up.weight"].shape) == (256, 256, 2, 2, 2)


def test_cpu_input_fails_loudly():
    m = _build("unet", dict(image_shape=(16, 16, 1), activation="elu", feature_maps=[8, 16], drop_values=[0, 0],
                            normalization="in", yx_down=[2], z_down=[2], isotropy=True, larger_io=False,
                            conv_layers=[2, 2], output_channels=[1]))
    with pytest.raises(_lib.B200Error):
        m(torch.zeros(1, 1, 16, 16))


def test_build_model_from_cfg_dict():
    cfg = {"PROBLEM": {"NDIM": "3D"}, "DATA": {"PATCH_SIZE": (32, 32, 32, 2)},
           "MODEL": {"ARCHITECTURE": "resunet", "FEATURE_MAPS": [16, 32], "NORMALIZATION": "gn", "ACTIVATION": "SiLU",
                     "DROPOUT_VALUES": [0, 0], "Z_DOWN": [0], "YX_DOWN": [0], "ISOTROPY": [True, True], "CONV_LAYERS": [2, 2]}}
    with contextlib.redirect_stdout(io.StringIO()):
        model, name, srcs, imports, files, args, stride = build_model(cfg, [1], ["F"], ["ce_sigmoid"], "cpu")
    assert name == "ResUNet" and stride == [1, 1, 1]
    assert args["z_down"] == [2] and args["activation"] == "silu"
    with pytest.raises(NotImplementedError):
        build_model({"PROBLEM": {"NDIM": "2D"}, "DATA": {"PATCH_SIZE": (32, 32, 1)}, "MODEL": {"ARCHITECTURE": "unetr"}},
                    [1], ["F"], ["ce_sigmoid"], "cpu")


def test_batchnorm_modules_follow_the_reference_factory():
    """'bn' -> nn.BatchNorm{2,3}d(momentum), 'sync_bn' -> nn.SyncBatchNorm (reference blocks.py:2117-2120, 2155-2158)."""
    m = _build("unet", dict(image_shape=(16, 16, 1), feature_maps=[8, 16], drop_values=[0, 0], normalization="bn",
                            yx_down=[2], z_down=[2], larger_io=False, conv_layers=[2, 2]))
    bns = [mod for mod in m.modules() if isinstance(mod, torch.nn.BatchNorm2d)]
    assert bns and all(b.momentum == 0.1 and b.track_running_stats for b in bns)
    assert any(k.endswith("running_var") for k in m.state_dict())
    m = _build("resunet", dict(image_shape=(8, 16, 16, 1), feature_maps=[8, 16], drop_values=[0, 0], normalization="sync_bn",
                               yx_down=[2], z_down=[2], larger_io=False, conv_layers=[2, 2]))
    assert any(isinstance(mod, torch.nn.SyncBatchNorm) for mod in m.modules())


def test_unsupported_options_fail_loudly():
    """Options outside the hot path (SURVEY §8 out-of-scope rows) raise instead of silently doing something else."""
    base = dict(image_shape=(16, 16, 1), feature_maps=[8, 16], drop_values=[0, 0], normalization="gn",
                yx_down=[2], z_down=[2], larger_io=False, conv_layers=[2, 2])
    with pytest.raises(NotImplementedError):
        _build("unet", dict(base, contrast=True))
    with pytest.raises(NotImplementedError):
        _build("unet", dict(base, upsampling_factor=(2, 2)))
    with pytest.raises(ValueError):
        _build("unet", dict(base, upsample_layer="nearest"))


def test_zero_arena_carves_one_cleared_buffer():
    """ops.ZeroArena (host logic, CPU tensors): first pass measures, later passes hand out aligned zeroed slices of one buffer;
    an outgrown buffer is retired, not freed (a captured CUDA graph may still point into it)."""
    import torch
    from biapy_b200 import ops
    a = ops.ZeroArena()
    a.begin("cpu")
    t = ops.zeros(10, torch.float64, "cpu")                  # no buffer yet: falls back to torch.zeros, records the need
    assert t.shape == (10,) and a.buf is None and a.need == 256
    ops.zeros(3, torch.float32, "cpu")
    a.end()
    assert ops.ARENA is None and a.need == 512
    a.begin("cpu")
    assert a.buf is not None and a.buf.numel() == 4096
    x = ops.zeros(10, torch.float64, "cpu")
    y = ops.zeros(3, torch.float32, "cpu")
    assert x.data_ptr() == a.buf.data_ptr() and y.data_ptr() - x.data_ptr() == 256 and x.dtype == torch.float64
    x.fill_(7.0)
    y.fill_(1.0)
    big = ops.zeros(4096, torch.float32, "cpu")              # does not fit: falls back, raises the high-water mark
    assert big.data_ptr() != a.buf.data_ptr() and a.need > 4096
    a.end()
    old = a.buf
    a.begin("cpu")
    assert a.buf is not old and a._retired == [old]
    x2 = ops.zeros(10, torch.float64, "cpu")
    assert float(x2.abs().sum()) == 0.0                      # cleared by the single fill of begin()
    a.end()
    assert ops.zeros(4, torch.float32, "cpu").sum().item() == 0.0     # no arena active: plain torch.zeros

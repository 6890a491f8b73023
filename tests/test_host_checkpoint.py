"""Checkpoint interop with BiaPy (biapy/utils/misc.py:328-660): host logic, runs without a GPU."""
import contextlib
import io
import os

import pytest
import torch

KW = dict(image_shape=(32, 32, 1), activation="elu", feature_maps=[8, 16], drop_values=[0, 0], normalization="bn", k_size=3,
          yx_down=[2], z_down=[2], isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1])


def _model(seed):
    from biapy_b200.models.resunet import ResUNet
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        return ResUNet(**KW)


class _Opt:
    def __init__(self):
        self.sd = {"state": {}, "param_groups": [{"lr": 0.5}]}

    def state_dict(self):
        return self.sd

    def load_state_dict(self, sd):
        self.sd = sd


@pytest.mark.parametrize("ext", ["pth", "safetensors"])
def test_save_and_load_round_trip(tmp_path, ext):
    from biapy_b200.utils.misc import load_model_checkpoint, save_model
    if ext == "safetensors":
        pytest.importorskip("safetensors")
    a, b = _model(0), _model(1)
    cfg = {"PATHS": {"CHECKPOINT": str(tmp_path), "CHECKPOINT_FILE": ""},
           "MODEL": {"LOAD_CHECKPOINT_EPOCH": "last_on_train", "ITEMS_TO_LOAD_FROM_CHECKPOINT": ["model", "optimizer", "epoch"]}}
    opt = _Opt()
    for epoch in (3, 12):
        path = save_model(tmp_path, cfg, "3.7.0", "job", epoch, a, [opt], model_build_kwargs=KW, extension=ext)
    assert os.path.basename(str(path)) == f"job-checkpoint-12.{ext}"
    assert any((x - y).abs().max() > 0 for x, y in zip(a.state_dict().values(), b.state_dict().values()))
    opt2 = _Opt()
    opt2.sd = None
    epoch, used = load_model_checkpoint(cfg, "job", b, "cpu", optimizer=[opt2])
    assert used.endswith(f"job-checkpoint-12.{ext}")
    for (k, x), y in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.equal(x, y), k
    if ext == "pth":
        assert epoch == 12 and opt2.sd == opt.sd
        info_cfg, ver = load_model_checkpoint(cfg, "job", b, "cpu", just_extract_checkpoint_info=True)
        assert ver == "3.7.0" and info_cfg["MODEL"]["LOAD_CHECKPOINT_EPOCH"] == "last_on_train"
        raw = torch.load(used, weights_only=True)                    # the reference's dictionary layout
        assert set(raw) == {"model_build_kwargs", "model", "optimizer", "epoch", "cfg", "biapy_version"}
    else:
        assert epoch == 0                                            # safetensors files carry the weights only


def test_layout_variants_and_unmatched_layers(tmp_path):
    from biapy_b200.utils.misc import load_model_checkpoint
    a, b = _model(0), _model(1)
    sd = a.state_dict()
    cfg = {"PATHS": {"CHECKPOINT": str(tmp_path), "CHECKPOINT_FILE": str(tmp_path / "other.pth")}, "MODEL": {}}
    for key in ("model_state_dict", "state_dict", None):
        torch.save({key: sd} if key else dict(sd), tmp_path / "other.pth")
        load_model_checkpoint(cfg, "job", b, "cpu")
        assert all(torch.equal(x, y) for x, y in zip(sd.values(), b.state_dict().values()))
    bad = dict(sd)
    bad["heads.0.weight"] = torch.zeros(3, 8, 1, 1)
    bad["not.a.layer"] = torch.zeros(1)
    torch.save({"model": bad}, tmp_path / "other.pth")
    with pytest.raises(RuntimeError):
        load_model_checkpoint(cfg, "job", _model(2), "cpu")
    c = _model(2)
    before = c.state_dict()["heads.0.weight"].clone()
    load_model_checkpoint(cfg, "job", c, "cpu", skip_unmatched_layers=True)
    assert torch.equal(c.state_dict()["heads.0.weight"], before)
    assert torch.equal(c.state_dict()["down_path.0.shortcut.0.weight"], sd["down_path.0.shortcut.0.weight"])
    cfg["PATHS"]["CHECKPOINT_FILE"] = str(tmp_path / "missing.pth")
    with pytest.raises(FileNotFoundError):
        load_model_checkpoint(cfg, "job", c, "cpu")


@pytest.mark.reference
def test_reference_written_checkpoint_loads_here(tmp_path):
    """A file written by the reference's own save_model for the reference's own ResUNet loads strictly into ours (and back)."""
    from types import SimpleNamespace
    import collections
    from pathlib import Path
    from oracle import ref_loader
    from biapy_b200.utils.misc import load_model_checkpoint, save_model
    R = ref_loader.load()
    fns = ref_loader._functions_from_source("biapy/utils/misc.py", ["save_model", "save_on_master", "cfg_to_plain_dict"],
                                            dict(torch=torch, Path=Path, collections=collections, is_main_process=lambda: True))
    torch.manual_seed(5)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_model = R.resunet.ResUNet(**KW)
    cfg = {"PATHS": {"CHECKPOINT": str(tmp_path), "CHECKPOINT_FILE": ""}, "MODEL": {"LOAD_CHECKPOINT_EPOCH": "last_on_train"}}
    fns["save_model"](tmp_path, cfg, "3.7.0", "ref", 7, ref_model, [], model_build_kwargs=None)
    ours = _model(9)
    load_model_checkpoint(cfg, "ref", ours, "cpu")
    for (k, x), (k2, y) in zip(ref_model.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and torch.equal(x, y), k
    save_model(tmp_path, cfg, "3.7.0", "ours", 1, ours, [])
    with contextlib.redirect_stdout(io.StringIO()):
        back = R.resunet.ResUNet(**KW)
    back.load_state_dict(torch.load(tmp_path / "ours-checkpoint-1.pth", weights_only=True)["model"], strict=True)

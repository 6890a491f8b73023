"""GPU parity of the device test-time augmentation (b200_orient_apply / b200_orient_reduce through
biapy_b200.data.post_processing.ensemble_predictions) against the golden fixtures of the reference's
ensemble_predictions and the numpy oracle: bit-exact for float32 predictions."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import port_tta

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
AO = {2: (0, 3, 1, 2), 3: (0, 4, 1, 2, 3)}
AOB = {2: (0, 2, 3, 1), 3: (0, 2, 3, 4, 1)}


def tta_image(shape, name):
    return np.random.default_rng(sum(map(ord, "tta" + name))).standard_normal(shape).astype(np.float32)


def _pred_func(nd):
    """The toy network of the fixtures, evaluated on the host from the device batch: the transforms under test are the
    device kernels on both sides of it."""
    def f(batch):
        assert batch.is_cuda
        return torch.from_numpy(port_tta.toy_pred_func(batch.cpu().numpy())).cuda().permute(AO[nd])
    return f


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tta_*.npz"))))
def test_tta_matches_reference_golden(path):
    from biapy_b200.data.post_processing.post_processing import ensemble_predictions
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    nd = meta["ndim"]
    img = tta_image(tuple(meta["shape"]), os.path.basename(path)[4:-4])
    out = ensemble_predictions(img, _pred_func(nd), AOB[nd], AO[nd], "cuda", nd, batch_size_value=meta["batch"], mode=meta["mode"],
                               group=meta["group"])
    assert out.is_cuda and out.dtype == torch.float32 and out.shape[0] == 1
    got = out.permute(AOB[nd])[0].cpu().numpy()
    assert np.array_equal(got, z["out"])


@pytest.mark.parametrize("shape,nd", [((16, 40, 24, 2), 3), ((33, 65, 3), 2), ((8, 64, 64, 1), 3)])
@pytest.mark.parametrize("mode", ["mean", "min", "max"])
def test_tta_matches_oracle(shape, nd, mode):
    from biapy_b200.data.post_processing.post_processing import ensemble_predictions
    img = np.random.default_rng(3).standard_normal(shape).astype(np.float32)
    ref = port_tta.ensemble_predictions(img, port_tta.toy_pred_func, nd, 4, mode, "auto")
    out = ensemble_predictions(torch.from_numpy(img).cuda(), _pred_func(nd), AOB[nd], AO[nd], "cuda", nd, batch_size_value=4, mode=mode)
    assert np.array_equal(out.permute(AOB[nd])[0].cpu().numpy(), ref)


def test_orientation_kernels_against_numpy():
    """b200_orient_apply == AxisTransform.apply (with front padding), for every orientation and storage dtype."""
    from biapy_b200.data.post_processing import post_processing as pp
    from biapy_b200.data.post_processing.tta import build_axis_transform_group
    rng = np.random.default_rng(0)
    for nd, shape in ((3, (5, 7, 7, 3)), (2, (9, 9, 2))):
        img = rng.standard_normal(shape).astype(np.float32)
        for dtype in (torch.float32, torch.bfloat16, torch.float16):
            x = torch.from_numpy(img).cuda().to(dtype)
            for t in build_axis_transform_group(nd):
                got = pp.orient_apply(x[None], t, (0,) * nd, "constant")[0]
                ref = port_tta.apply(x.float().cpu().numpy(), t.perm, t.sign)
                assert np.array_equal(got.float().cpu().numpy(), ref), (nd, dtype, t)
    # padding modes of _pad_for_orientations: reflect, edge (pad >= dim) and constant
    img = rng.standard_normal((4, 3, 9, 1)).astype(np.float32)
    x = torch.from_numpy(img).cuda()
    ident = build_axis_transform_group(3, "none")[0]
    for mode in ("reflect", "edge", "constant"):
        pad = (0, 2, 0) if mode == "reflect" else (0, 6, 0)
        got = pp.orient_apply(x[None], ident, pad, mode)[0].cpu().numpy()
        ref = np.pad(img, [(p, 0) for p in pad] + [(0, 0)], mode=mode)
        assert np.array_equal(got, ref), mode


def test_tta_model_pipeline_and_errors():
    """TTA around the real engine forward: equals the oracle ensemble of the same model evaluated per orientation on the CPU."""
    import contextlib
    import io
    from biapy_b200.data.post_processing.post_processing import ensemble_predictions
    from biapy_b200.models.unet import U_Net
    from oracle import port_models
    kw = dict(image_shape=(16, 16, 16, 1), activation="elu", feature_maps=[8, 16], drop_values=[0, 0], normalization="in", k_size=3,
              yx_down=[2], z_down=[2], isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = U_Net(**kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().set_engine(dtype=torch.float32).eval()
    img = np.random.default_rng(1).standard_normal((16, 16, 16, 1)).astype(np.float32)
    with torch.no_grad():
        out = ensemble_predictions(img, lambda b: m(b.permute(AO[3])), AOB[3], AO[3], "cuda", 3, batch_size_value=4, mode="mean")

        def cpu_pred(batch):
            y = port_models.forward("unet", sd, torch.from_numpy(batch).permute(AO[3]), training=False, **kw)
            return y.permute(AOB[3]).numpy()
        ref = port_tta.ensemble_predictions(img, cpu_pred, 3, 4, "mean", "auto")
    got = out.permute(AOB[3])[0].cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())      # fp32 engine vs ATen: the 1e-3 parity bar
    with pytest.raises(NotImplementedError):
        ensemble_predictions(img, lambda b: b, AOB[3], AO[3], "cuda", 3, tta_spec=object())
    with pytest.raises(ValueError):
        ensemble_predictions(img[0], lambda b: b, AOB[3], AO[3], "cuda", 3)
    with pytest.raises(AssertionError):
        ensemble_predictions(img, lambda b: b, AOB[3], AO[3], "cuda", 3, mode="median")


def test_sliding_window_inference_with_tta_matches_oracle_pipeline():
    """predict_volume(tta=True) == crop -> per-patch ensemble of sigmoid(model) -> merge, all on the CPU oracle."""
    import contextlib
    import io
    from biapy_b200.engine.inference import predict_volume
    from biapy_b200.models.resunet import ResUNet
    from oracle import port_models, port_stitch
    kw = dict(image_shape=(16, 16, 16, 2), activation="silu", feature_maps=[16, 32], drop_values=[0, 0], normalization="gn", k_size=3,
              yx_down=[2], z_down=[2], isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**kw)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda().set_engine(dtype=torch.float32).eval()
    vol = np.random.default_rng(5).standard_normal((24, 16, 28, 2)).astype(np.float32)
    patch, ov, pad = (16, 16, 16, 2), (0.25, 0.0, 0.25), (0, 0, 0)
    got = predict_volume(m, vol, patch, overlap=ov, padding=pad, batch_size=2, head_activations=["ce_sigmoid"], tta=True)
    patches, _ = port_stitch.crop_3d(vol, patch, ov, pad, "reflect")

    def cpu_pred(batch):
        with torch.no_grad():
            y = port_models.forward("resunet", sd, torch.from_numpy(batch).permute(AO[3]), training=False, **kw)
            return port_models.apply_head_activations(y, ["ce_sigmoid"], training=False).permute(AOB[3]).numpy()
    p = np.stack([port_tta.ensemble_predictions(patches[i], cpu_pred, 3, 2, "mean", "auto") for i in range(patches.shape[0])])
    ref = port_stitch.merge_3d(np.ascontiguousarray(p), (24, 16, 28, 1), ov, pad)
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-4

"""GPU test of the workflow surface: a YAML-configured `BiaPy` object trains and predicts through the B200 engine and agrees
with the CPU oracle pipeline (normalise -> crop -> forward -> sigmoid -> merge -> binarise) and with the bare Trainer."""
import contextlib
import copy
import io

import numpy as np
import pytest
import torch

from oracle import port_models, port_norm, port_stitch

pytestmark = pytest.mark.gpu

YAML = """
PROBLEM: {TYPE: SEMANTIC_SEG, NDIM: 3D}
DATA:
    PATCH_SIZE: (32, 32, 32, 1)
    NORMALIZATION: {TYPE: scale_range}
    TEST: {OVERLAP: "(0.25, 0.25, 0)", PADDING: "(4, 0, 2)"}
MODEL:
    ARCHITECTURE: resunet
    FEATURE_MAPS: [16, 32, 64]
    DROPOUT_VALUES: [0, 0, 0]
    ISOTROPY: [True, True, True]
    CONV_LAYERS: [2, 2, 2]
    Z_DOWN: [0, 0]
    YX_DOWN: [0, 0]
    NORMALIZATION: gn
    ACTIVATION: silu
TRAIN: {OPTIMIZER: ADAMW, LR: 1.E-3, BATCH_SIZE: 2, W_DECAY: 0.02}
"""
KW = dict(image_shape=(32, 32, 32, 1), activation="silu", feature_maps=[16, 32, 64], drop_values=[0, 0, 0], normalization="gn", k_size=3,
          yx_down=[2, 2], z_down=[2, 2], isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1])


def _biapy(dtype):
    from biapy_b200._biapy import BiaPy
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        return BiaPy(YAML, name="job", run_id=1, engine_dtype=dtype)


def test_yaml_configured_prediction_matches_oracle_pipeline():
    b = _biapy(torch.float32)
    assert type(b.workflow).__name__ == "Semantic_Segmentation_Workflow" and b.job_identifier == "job_1"
    sd = {k: v.detach().cpu().clone() for k, v in b.workflow.model.state_dict().items()}
    img = np.random.default_rng(3).integers(0, 256, (48, 40, 70, 1)).astype(np.uint8)
    pred, mask = b.workflow.process_test_sample(img)
    # oracle: the reference's numpy / ATen path
    x, info = port_norm.normalize_image(img.copy(), dict(b.workflow.test_norm_module))
    patch, ov, pad = (32, 32, 32, 1), (0.25, 0.25, 0.0), (4, 0, 2)
    patches, _ = port_stitch.crop_3d(x, patch, ov, pad, "reflect")
    with torch.no_grad():
        y = port_models.forward("resunet", sd, torch.from_numpy(patches).permute(0, 4, 1, 2, 3), training=False, **KW)
        p = port_models.apply_head_activations(y, ["ce_sigmoid"], training=False).permute(0, 2, 3, 4, 1).numpy()
    ref = port_stitch.merge_3d(np.ascontiguousarray(p), (48, 40, 70, 1), ov, pad)
    assert pred.shape == ref.shape and np.abs(pred - ref).max() < 1e-4
    assert b.workflow.current_sample["norm_info"]["per_channel_info"] == info["per_channel_info"]
    # after_merge_patches binarises with the Otsu threshold of the whole prediction (semantic_seg.py:429): exactly the oracle's
    # threshold for the engine's own prediction, and the oracle pipeline's mask wherever the 1e-4 difference cannot matter
    th = port_norm.threshold_otsu(pred)
    assert mask.dtype == np.uint8 and np.array_equal(mask, (pred > th).astype(np.uint8))
    th_ref = port_norm.threshold_otsu(ref)
    assert abs(float(th) - float(th_ref)) <= 1.01 / 256                # at most one histogram bin apart
    away = np.abs(ref - th) > 1e-3
    if th == th_ref:
        assert np.array_equal(mask[away], port_norm.binarize(ref, 2, None)[away])
    assert np.array_equal(b.predict(img), pred)
    # patches through predict_batches_in_test (the reference's per-batch entry point)
    pb = b.workflow.predict_batches_in_test(patches[:3])
    assert np.abs(pb - p[:3]).max() < 1e-4


def test_yaml_configured_training_equals_the_trainer():
    from biapy_b200.engine.train import Trainer
    from biapy_b200.models.resunet import ResUNet
    g = torch.Generator().manual_seed(1)
    X = torch.randn(4, 32, 32, 32, 1, generator=g).numpy()
    Y = (torch.rand(4, 32, 32, 32, 1, generator=g) < 0.3).float().numpy()
    b = _biapy(torch.float32)
    losses = b.train(X, Y, steps=3)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**dict(KW, output_channel_info=["pred0"], head_activations=["ce_sigmoid"]))
    m = m.cuda().set_engine(dtype=torch.float32)
    tr = Trainer(m, loss="bce", optimizer="adamw", lr=1e-3, betas=(0.9, 0.999), weight_decay=0.02)
    ref = []
    for s in range(3):
        lo = (s * 2) % 3
        ref.append(float(tr.step(X[lo:lo + 2], Y[lo:lo + 2]).item()))
    # same kernels on the same data; fp32 atomics (weight gradients) make the trajectories agree to rounding, not bit for bit
    assert all(abs(a - r) <= 1e-6 * max(1.0, abs(r)) for a, r in zip(losses, ref)) and losses[-1] < losses[0]


YAML_2D = """
PROBLEM: {TYPE: DENOISING, NDIM: 2D}
DATA:
    PATCH_SIZE: (64, 64, 1)
    NORMALIZATION: {TYPE: zero_mean_unit_variance, ZERO_MEAN_UNIT_VAR: {MEAN_VAL: [100.0], STD_VAL: [40.0]}}
    TEST: {OVERLAP: "(0.25, 0.5)", PADDING: "(8, 4)"}
MODEL:
    ARCHITECTURE: unet
    FEATURE_MAPS: [16, 32, 64]
    DROPOUT_VALUES: [0, 0, 0]
    ISOTROPY: [True, True, True]
    CONV_LAYERS: [2, 2, 2]
    Z_DOWN: [0, 0]
    YX_DOWN: [0, 0]
TRAIN: {BATCH_SIZE: 3}
"""
KW2D = dict(image_shape=(64, 64, 1), activation="elu", feature_maps=[16, 32, 64], drop_values=[0, 0, 0], normalization="in", k_size=3,
            yx_down=[2, 2], z_down=[2, 2], isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1])


def test_2d_denoising_workflow_matches_oracle_pipeline():
    """2D image through the denoising workflow: normalise (given mean / std) -> 2D crop -> U-Net -> linear head -> 2D merge ->
    undo the normalisation back to uint8, against the same chain on the CPU oracle."""
    from biapy_b200._biapy import BiaPy
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        b = BiaPy(YAML_2D, engine_dtype=torch.float32)
    assert type(b.workflow).__name__ == "Denoising_Workflow" and b.workflow.head_activations == ["linear"]
    sd = {k: v.detach().cpu().clone() for k, v in b.workflow.model.state_dict().items()}
    img = np.random.default_rng(5).integers(0, 256, (150, 170, 1)).astype(np.uint8)
    pred, restored = b.workflow.process_test_sample(img)
    x, info = port_norm.normalize_image(img.copy(), dict(b.workflow.test_norm_module))
    patch, ov, pad = (64, 64, 1), (0.25, 0.5), (8, 4)
    patches, _ = port_stitch.crop_2d(x[None], patch, ov, pad, "reflect")
    with torch.no_grad():
        y = port_models.forward("unet", sd, torch.from_numpy(patches).permute(0, 3, 1, 2), training=False, **KW2D)
    ref = port_stitch.merge_2d(np.ascontiguousarray(y.permute(0, 2, 3, 1).numpy()), (1, 150, 170, 1), ov, pad)[0]
    assert pred.shape == ref.shape and np.abs(pred - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
    back = port_norm.undo_image_norm(ref, info)
    assert restored.dtype == np.uint8 and restored.shape == back.shape
    # rounding to uint8 can flip where the two float predictions straddle x.5
    assert np.abs(restored.astype(np.int32) - back.astype(np.int32)).max() <= 1 and (restored != back).mean() < 1e-3


YAML_TRAIN = YAML.replace("TRAIN: {OPTIMIZER: ADAMW, LR: 1.E-3, BATCH_SIZE: 2, W_DECAY: 0.02}",
                          "TRAIN: {OPTIMIZER: ADAMW, LR: 2.E-3, BATCH_SIZE: 2, W_DECAY: 0.02, EPOCHS: 2, PATIENCE: 1,\n"
                          "        LR_SCHEDULER: {NAME: onecycle}}")


def test_epoch_loop_with_onecycle_follows_torch_cpu():
    """``Base_Workflow.train`` (prepare_optimizer -> train_one_epoch -> evaluate -> early stopping) on in-memory generators: the
    loss of every update follows torch CPU running the reference's recipe -- AdamW + ``OneCycleLR`` stepped after every update
    (``train_engine.py:166-174``), which also cycles beta1 -- and the validation loss is the eval-mode forward loss."""
    from biapy_b200._biapy import BiaPy
    assert "onecycle" in YAML_TRAIN
    g = torch.Generator().manual_seed(5)
    X = torch.randn(6, 32, 32, 32, 1, generator=g)
    Y = (torch.rand(6, 32, 32, 32, 1, generator=g) < 0.3).float()
    train_gen = [(X[i:i + 2].numpy(), Y[i:i + 2].numpy()) for i in (0, 2)]
    val_gen = [(X[4:6].numpy(), Y[4:6].numpy())]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        b = BiaPy(YAML_TRAIN, name="job", run_id=2, engine_dtype=torch.float32)
    wf = b.workflow
    sd = {k: v.detach().cpu().clone() for k, v in wf.model.state_dict().items()}
    with contextlib.redirect_stdout(io.StringIO()):
        hist = wf.train(train_gen, val_gen)
    # torch CPU: the reference's loop
    sd_r = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.AdamW(list(sd_r.values()), lr=2e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.02)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, 2e-3, epochs=2, steps_per_epoch=2)
    ref = []
    for epoch in range(2):
        tl, lr = [], None
        for xb, yb in train_gen:
            y = port_models.forward("resunet", sd_r, torch.from_numpy(xb).permute(0, 4, 1, 2, 3), training=True, **KW)
            loss = port_models.bce_with_logits_loss(y, torch.from_numpy(yb).permute(0, 4, 1, 2, 3))
            opt.zero_grad()
            loss.backward()
            lr = opt.param_groups[0]["lr"]
            opt.step()
            sched.step()
            tl.append(loss.item())
        with torch.no_grad():
            yv = port_models.forward("resunet", sd_r, torch.from_numpy(val_gen[0][0]).permute(0, 4, 1, 2, 3), training=False, **KW)
            vl = port_models.bce_with_logits_loss(yv, torch.from_numpy(val_gen[0][1]).permute(0, 4, 1, 2, 3)).item()
        ref.append((sum(tl) / len(tl), lr, vl))
    assert len(hist) == 2 and [h["epoch"] for h in hist] == [0, 1]
    for h, (tl, lr, vl) in zip(hist, ref):
        assert abs(h["train_loss"] - tl) < 1e-3 * max(1.0, abs(tl)), (h, tl)
        assert h["train_lr"] == pytest.approx(lr, rel=1e-9)
        assert abs(h["test_loss"] - vl) < 1e-3 * max(1.0, abs(vl)), (h, vl)
    assert wf.trainer.param_groups[0]["betas"][0] == pytest.approx(opt.param_groups[0]["betas"][0], rel=1e-9)
    assert wf.val_best_loss == pytest.approx(min(h["test_loss"] for h in hist))

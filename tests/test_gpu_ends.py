"""GPU parity of the kernels at the two ends of the inference path (biapy_b200/data/norm.py): normalize_image /
undo_image_norm against the reference's golden vectors (bit-exact once the statistics are fixed, 1e-6 when mean / std are
computed on the device) and the binarisation behind the merge against numpy."""
import copy
import json
import os

import numpy as np
import pytest
import torch

from oracle import port_norm

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_normalize_image_matches_reference_golden():
    from biapy_b200.data.norm import normalize_image, undo_image_norm
    for img, mod, y_ref, u_ref, info_ref, key in port_norm.golden_cases(GOLDEN):
        y, info = normalize_image(img.copy(), copy.deepcopy(mod))
        assert isinstance(y, np.ndarray) and y.dtype == np.float32
        stats_from_data = mod["type"] == "zero_mean_unit_variance" and "mean" not in mod
        if stats_from_data:
            # mean / std come from fp64 sums on the device (numpy: pairwise float32): same statistics to float32 rounding
            for k, ch in info["per_channel_info"].items():
                for q in ("mean", "std"):
                    assert abs(ch[q] - info_ref["per_channel_info"][k][q]) <= 2e-6 * max(1.0, abs(info_ref["per_channel_info"][k][q])), (key, q)
            assert np.abs(y - y_ref).max() <= 1e-5 * max(1.0, np.abs(y_ref).max()), key
        else:
            assert json.loads(json.dumps(info)) == info_ref, key
            assert np.array_equal(y, y_ref), key
        # with the reference's statistics handed back in, everything is bit-exact -- forward and inverse
        y2, info2 = normalize_image(torch.from_numpy(img).cuda(), dict(copy.deepcopy(mod), per_channel_info=info_ref["per_channel_info"]))
        assert y2.is_cuda and np.array_equal(y2.cpu().numpy(), y_ref), key
        u = undo_image_norm(y_ref.copy(), info_ref)
        assert u.dtype == u_ref.dtype and np.array_equal(u, u_ref), key


def test_normalize_image_errors_and_apply_norm_false():
    from biapy_b200.data.norm import normalize_image
    img = np.random.default_rng(0).integers(0, 255, (4, 8, 8, 1)).astype(np.uint8)
    with pytest.raises(AssertionError):                  # neither a percentile nor a value for the bounds (norm.py:446, 457)
        normalize_image(img, dict(type="div", percentile_clip=True, out_dtype="float32", per_lower_bound=-1, per_upper_bound=-1))
    with pytest.raises(NotImplementedError):
        normalize_image(img, dict(type="div", percentile_clip=False, out_dtype="uint8"))
    with pytest.raises(AssertionError):
        normalize_image(img, dict(type="bogus", percentile_clip=False, out_dtype="float32"))
    mod = dict(type="scale_range", percentile_clip=False, out_dtype="float32")
    y, info = normalize_image(img, mod, apply_norm=False)
    yr, ir = port_norm.normalize_image(img.copy(), mod, apply_norm=False)
    assert np.array_equal(y, yr) and json.loads(json.dumps(info)) == json.loads(json.dumps(ir))


@pytest.mark.parametrize("kind", ["float32", "uint16", "uint8"])
def test_percentile_bounds_from_the_data(kind):
    """``per_lower_bound`` / ``per_upper_bound``: exact order statistics from the device radix select, combined like
    ``np.percentile`` for numpy images (the oracle, pinned to the reference by the golden cases) and like the reference's
    ``torch_percentile`` (``kthvalue``, norm.py:475-497) for tensors -- bit-exact bounds either way."""
    from biapy_b200.data.norm import normalize_image
    rng = np.random.default_rng(11)
    if kind == "float32":
        img = (rng.standard_normal((6, 21, 24, 2)) * 40 - 5).astype(np.float32)
        img[0, 0, :5, 0] = img[1, 1, :5, 0]                               # ties
    elif kind == "uint16":
        img = rng.integers(0, 50000, (5, 17, 19, 3)).astype(np.uint16)
    else:
        img = rng.integers(0, 256, (40, 33, 2)).astype(np.uint8)
    mod = dict(type="scale_range", percentile_clip=True, out_dtype="float32", per_lower_bound=1.5, per_upper_bound=99.2)
    y, info = normalize_image(img.copy(), copy.deepcopy(mod))
    yr, ir = port_norm.normalize_image(img.copy(), copy.deepcopy(mod))
    assert json.loads(json.dumps(info)) == json.loads(json.dumps(ir))
    assert np.array_equal(y, yr)
    y2, info2 = normalize_image(torch.from_numpy(img).cuda(), copy.deepcopy(mod))
    for k in range(img.shape[-1]):
        ch = torch.from_numpy(img[..., k].astype(np.float32)).reshape(-1)
        n = ch.numel()
        lo = ch.kthvalue(1 + round(0.01 * 1.5 * (n - 1))).values.item()
        hi = ch.kthvalue(1 + round(0.01 * 99.2 * (n - 1))).values.item()
        got = info2["per_channel_info"][str(k)]
        assert (got["lower_bound_val"], got["upper_bound_val"]) == (lo, hi), k
        d = np.clip(img[..., k].astype(np.float32), lo, hi)
        want = (d - d.min()) / max(float(d.max()) - float(d.min()), 1e-6)
        assert np.abs(y2[..., k].cpu().numpy() - want).max() <= 1e-6, k


@pytest.mark.parametrize("shape", [(9, 33, 17, 1), (5, 12, 12, 2), (7, 9, 11, 5)])
def test_binarize_prediction(shape):
    from biapy_b200.data.norm import binarize_prediction
    rng = np.random.default_rng(4)
    pred = rng.random(shape).astype(np.float32)
    pred.flat[::7] = 0.5                                    # ties sit exactly on the threshold
    assert np.array_equal(binarize_prediction(pred, 2), port_norm.binarize(pred, 2))
    assert np.array_equal(binarize_prediction(pred, 2, threshold=0.37), port_norm.binarize(pred, 2, 0.37))
    if shape[-1] > 2:
        pred[..., 1] = pred[..., 3]                         # equal maxima: the first index wins
        got = binarize_prediction(torch.from_numpy(pred).cuda(), shape[-1])
        ref = port_norm.binarize(pred, shape[-1])
        assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), ref)
        assert binarize_prediction(pred, 300).dtype == np.uint16


def test_image_stats_kernel():
    from biapy_b200.data.norm import image_stats
    rng = np.random.default_rng(1)
    for dt in (np.uint8, np.uint16, np.float32):
        img = (rng.standard_normal((6, 40, 50, 3)) * 40 + 90).clip(0, 65000).astype(dt)
        img[..., 2] = rng.integers(0, 2, img.shape[:-1]).astype(dt)
        st = image_stats(torch.from_numpy(img).cuda())
        f = img.astype(np.float64)
        assert np.array_equal(st[:, 0], f.min((0, 1, 2))) and np.array_equal(st[:, 1], f.max((0, 1, 2)))
        assert np.allclose(st[:, 2], f.sum((0, 1, 2)), rtol=1e-12) and np.allclose(st[:, 3], (f * f).sum((0, 1, 2)), rtol=1e-12)
        assert st[:, 4].tolist() == [0.0, 0.0, 1.0]
        clip = np.array([[1, 20, 100], [0, 0, 0], [0, 0, 0]], np.float32)
        st2 = image_stats(torch.from_numpy(img).cuda(), clip)
        assert st2[0, 0] == max(f[..., 0].min(), 20) and st2[0, 1] == min(f[..., 0].max(), 100) and st2[1, 0] == st[1, 0]


def test_otsu_threshold_matches_the_oracle_bit_for_bit():
    """after_merge_patches' Otsu threshold (semantic_seg.py:429): device min / max + numpy-rule histogram, skimage's arithmetic on
    the host -- the same float32 threshold and the same 256 counts as the CPU oracle / the committed fixture, hence the same mask."""
    from biapy_b200 import _lib, ops
    from biapy_b200.data.norm import binarize_prediction, threshold_otsu
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "otsu_cases.npz"))
    for name, img in port_norm.otsu_cases().items():
        th = threshold_otsu(img)
        assert np.float32(th) == z["th." + name], (name, th, z["th." + name])
        assert np.array_equal(binarize_prediction(img, 2, threshold=None), port_norm.binarize(img, 2, None)), name
        if "counts." + name in z.files:
            dev = torch.from_numpy(img).cuda().reshape(-1)
            edges = torch.from_numpy(np.linspace(img.min(), img.max(), 257, dtype=np.float32)).cuda()
            counts = torch.empty(256, dtype=torch.int64, device="cuda")
            ops._launch("b200_edge_hist", ops._ptr(dev), dev.numel(), ops._ptr(edges), 256, ops._ptr(counts), _lib.stream_ptr())
            assert np.array_equal(counts.cpu().numpy(), z["counts." + name]), name
    # a sigmoid volume of the size the workflow produces
    big = torch.sigmoid(torch.randn(96, 160, 160, 1, generator=torch.Generator().manual_seed(0)) * 4).numpy()
    assert np.float32(threshold_otsu(big)) == port_norm.threshold_otsu(big)

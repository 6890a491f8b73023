"""GPU parity tests for the sliding-window stitch path (C ABI -> CUDA) against the reference's golden vectors
and the numpy oracle.  Integer bookkeeping and byte copies must be bit-exact; the float32 overlap-add is
bit-exact too because the kernel reproduces numpy's operation order."""
import glob
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import port_stitch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "stitch_s3d_*.npz"))))
def test_golden_3d(path):
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap, merge_3D_data_with_overlap
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    kw = dict(overlap=tuple(meta["overlap"]), padding=tuple(meta["padding"]), verbose=False)
    patches, coords = crop_3D_data_with_overlap(z["vol"], tuple(meta["patch"]), pad_type=meta["pad_type"], **kw)
    starts = np.array([[c.z_start, c.y_start, c.x_start] for c in coords], dtype=np.int64)
    assert np.array_equal(starts, z["starts"])
    assert patches.dtype == np.float32 and zlib.crc32(np.ascontiguousarray(patches).tobytes()) == int(z["patches_crc"])
    merged = merge_3D_data_with_overlap(z["pred"], tuple(meta["vshape"]), **kw)
    assert merged.dtype == np.float32 and np.array_equal(merged, z["merged"])       # bit-exact
    m16 = merge_3D_data_with_overlap(z["pred"].astype(np.float16), tuple(meta["vshape"]), **kw)
    assert m16.dtype == np.float16 and np.array_equal(m16, z["merged16"])
    rt = merge_3D_data_with_overlap(patches, tuple(meta["vshape"]), **kw)
    assert np.array_equal(rt, z["roundtrip"])
    # mask path: second array merged with the same weights
    a, b = merge_3D_data_with_overlap(z["pred"], tuple(meta["vshape"]), data_mask=z["pred"] * 2, **kw)
    assert np.array_equal(a, z["merged"]) and np.array_equal(b, port_stitch.merge_3d(z["pred"] * 2, meta["vshape"], meta["overlap"], meta["padding"]))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "stitch_s2d_*.npz"))))
def test_golden_2d(path):
    from biapy_b200.data.data_2D_manipulation import crop_data_with_overlap, merge_data_with_overlap
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    kw = dict(overlap=tuple(meta["overlap"]), padding=tuple(meta["padding"]), verbose=False)
    patches, coords = crop_data_with_overlap(z["vol"], tuple(meta["patch"]), pad_type=meta["pad_type"], **kw)
    starts = np.array([[c.y_start, c.x_start] for c in coords], dtype=np.int64)
    assert np.array_equal(starts, z["starts"])
    assert zlib.crc32(np.ascontiguousarray(patches).tobytes()) == int(z["patches_crc"])
    merged = merge_data_with_overlap(z["pred"], tuple(meta["vshape"]), **kw)
    assert np.array_equal(merged, z["merged"])


@pytest.mark.parametrize("shape,patch,ov,pad,mode", [
    ((37, 41, 29, 3), (16, 12, 10, 3), (0.3, 0.0, 0.6), (3, 2, 1), "reflect"),
    ((20, 20, 20, 1), (20, 10, 8, 1), (0.5, 0.25, 0.0), (0, 1, 2), "symmetric"),
    ((33, 18, 25, 2), (8, 8, 8, 2), (0.0, 0.0, 0.0), (2, 2, 2), "zeros"),
    ((15, 16, 17, 1), (6, 8, 6, 1), (0.2, 0.2, 0.2), (1, 1, 1), "edge"),
])
def test_random_vs_oracle(shape, patch, ov, pad, mode):
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap, merge_3D_data_with_overlap
    rng = np.random.default_rng(3)
    vol = rng.standard_normal(shape).astype(np.float32)
    patches, coords = crop_3D_data_with_overlap(vol, patch, overlap=ov, padding=pad, verbose=False, pad_type=mode)
    ref_patches, ref_starts = port_stitch.crop_3d(vol, patch, ov, pad, mode)
    assert np.array_equal(patches, ref_patches)
    pred = (patches + rng.standard_normal(patches.shape)).astype(np.float32)
    got = merge_3D_data_with_overlap(pred, shape, overlap=ov, padding=pad, verbose=False)
    assert np.array_equal(got, port_stitch.merge_3d(pred, shape, ov, pad))
    # uint8 crop (byte copies)
    v8 = rng.integers(0, 255, shape, dtype=np.uint8)
    p8, _ = crop_3D_data_with_overlap(v8, patch, overlap=ov, padding=pad, verbose=False, pad_type=mode)
    assert np.array_equal(p8, port_stitch.crop_3d(v8, patch, ov, pad, mode)[0])


def test_device_tensors_stay_on_device_and_bf16():
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap, merge_3D_data_with_overlap
    vol = torch.randn(40, 40, 40, 2, device="cuda")
    patches, coords = crop_3D_data_with_overlap(vol, (16, 16, 16, 2), overlap=(0.25,) * 3, verbose=False)
    assert patches.is_cuda and patches.shape[0] == len(coords)
    rt = merge_3D_data_with_overlap(patches, (40, 40, 40, 2), overlap=(0.25,) * 3, verbose=False)
    assert rt.is_cuda and (rt - vol).abs().max().item() < 2e-6
    rb = merge_3D_data_with_overlap(patches.bfloat16(), (40, 40, 40, 2), overlap=(0.25,) * 3, verbose=False)
    assert rb.dtype == torch.bfloat16 and (rb.float() - vol).abs().max().item() < 0.05


def test_full_size_cfg3_properties():
    """BASELINE cfg 3 grid (512^3, 128^3 patches, 25% overlap -> 216 patches) on the device: crop -> merge is the
    identity, a constant field stays constant, and the merge is linear."""
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap, merge_3D_data_with_overlap
    g = torch.Generator(device="cuda").manual_seed(1)
    vol = torch.randn(512, 512, 512, 1, device="cuda", generator=g)
    patches, coords = crop_3D_data_with_overlap(vol, (128, 128, 128, 1), overlap=(0.25,) * 3, verbose=False)
    assert patches.shape == (216, 128, 128, 128, 1)
    assert sorted({c.z_start for c in coords}) == [0, 77, 154, 231, 308, 384]
    rt = merge_3D_data_with_overlap(patches, (512, 512, 512, 1), overlap=(0.25,) * 3, verbose=False)
    assert (rt - vol).abs().max().item() < 2e-6
    ones = merge_3D_data_with_overlap(torch.ones_like(patches), (512, 512, 512, 1), overlap=(0.25,) * 3, verbose=False)
    assert (ones - 1).abs().max().item() < 1e-6
    del ones
    a = merge_3D_data_with_overlap(patches * 2.0, (512, 512, 512, 1), overlap=(0.25,) * 3, verbose=False)
    assert (a - 2 * rt).abs().max().item() < 4e-6

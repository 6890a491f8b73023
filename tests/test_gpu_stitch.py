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


def test_patch_range_crop_shard_and_slab_merge_equal_the_full_calls():
    """Building blocks of the sharded sliding-window inference: cropping a patch range (from the whole volume or from just the
    planes `planes_needed` names) gives the same bytes as the rows of the full crop; merging z slabs of the output gives the same
    bytes as the planes of the full merge, for every merge kernel variant the library holds."""
    import torch
    from biapy_b200.data import _stitch
    g = torch.Generator().manual_seed(3)
    vshape, patch, ov, pad = (45, 30, 34, 2), (16, 12, 16), (0.25, 0.5, 0.25), (2, 0, 3)
    vol = torch.randn(vshape, generator=g).cuda()
    axes = [_stitch.Axis(vshape[i], patch[i], pad[i], ov[i]) for i in range(3)]
    starts_c, starts_m, wins = [a.starts(0) for a in axes], [a.starts(1) for a in axes], [a.window() for a in axes]
    full = _stitch.crop_device(vol, patch, starts_c, pad, "reflect")
    n, n_yx = full.shape[0], axes[1].n * axes[2].n
    for first, end in ((0, n), (3, 11), (n - 5, n), (7, 8)):
        part = _stitch.crop_device(vol, patch, starts_c, pad, "reflect", patch_range=(first, end))
        assert torch.equal(part, full[first:end])
        z0, z1 = _stitch.planes_needed(vshape[0], patch[0], pad[0], starts_c[0], n_yx, (first, end), "reflect")
        shard = _stitch.VolumeShard(vol[z0:z1].contiguous(), z0, vshape[0])
        assert torch.equal(_stitch.crop_device(shard, patch, starts_c, pad, "reflect", patch_range=(first, end)), full[first:end])
    pred = torch.randn(full.shape[:4] + (1,), generator=g).cuda()
    ref = _stitch.merge_device(pred, vshape[:3], starts_m, wins, pad)
    for z0, z1 in ((0, 45), (0, 7), (7, 29), (44, 45)):
        slab = _stitch.merge_device(pred, vshape[:3], starts_m, wins, pad, z_range=(z0, z1))
        assert torch.equal(slab, ref[z0:z1])
    # fp16 predictions -> fp32 volume, the cheaper exchange format of the sharded path
    ref16 = _stitch.merge_device(pred.half(), vshape[:3], starts_m, wins, pad, out_dtype=torch.float32)
    assert torch.equal(_stitch.merge_device(pred.half(), vshape[:3], starts_m, wins, pad, out_dtype=torch.float32, z_range=(5, 20)), ref16[5:20])


@pytest.mark.parametrize("variant", ["slot", "cover", "plain"])
def test_merge_kernel_variants_bit_identical(variant):
    """The three overlap-add kernels (B200_MERGE_KERNEL) against the numpy oracle on a grid with padding and a 50 % overlap axis
    (three covering patches: the slot kernel's mask-walk branch).  The variant is read once per process, so each runs in a child."""
    import subprocess
    import sys
    code = (
        "import numpy as np, torch, sys\n"
        "sys.path.insert(0, '.')\n"
        "from biapy_b200.data.data_3D_manipulation import merge_3D_data_with_overlap\n"
        "from oracle import port_stitch\n"
        "rng = np.random.default_rng(5)\n"
        "for shape, patch, ov, pad in (((40, 36, 44, 1), (16, 16, 16), (0.25, 0.25, 0.25), (0, 0, 0)),\n"
        "                              ((33, 20, 50, 2), (12, 8, 16), (0.5, 0.25, 0.6), (1, 0, 2))):\n"
        "    n = port_stitch.crop_3d(np.zeros(shape, np.float32), patch + (shape[-1],), ov, pad)[0].shape[0]\n"
        "    pred = rng.standard_normal((n,) + patch + (shape[-1],)).astype(np.float32)\n"
        "    got = merge_3D_data_with_overlap(pred, shape, overlap=ov, padding=pad, verbose=False)\n"
        "    assert np.array_equal(got, port_stitch.merge_3d(pred, shape, ov, pad)), shape\n"
        "print('OK')\n")
    env = dict(os.environ, B200_MERGE_KERNEL=variant)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

"""CPU tests: the oracle restatement (oracle/port_*.py) against the golden vectors produced by the real
reference (oracle/make_golden.py).  Index bookkeeping must be bit-exact, merge output bit-exact (same numpy
operations in the same order), model outputs within 1e-5 (same ATen kernels, different graph plumbing)."""
import glob
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import port_models, port_stitch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _grids():
    with open(os.path.join(GOLDEN, "grids.json")) as f:
        return json.load(f)


def test_grid_known_answers_3d():
    g = _grids()
    for case in g["3d"]:
        plans, starts = port_stitch.crop_grid(case["shape"], case["patch"], case["overlap"], case["padding"])
        assert starts.shape[0] == case["n"]
        for ax, key in enumerate("zyx"):
            got = sorted({pl for pl in starts[:, ax].tolist()})
            assert got == case[key], (case, key)
        allc = np.stack([starts[:, 0], starts[:, 0] + case["patch"][0], starts[:, 1], starts[:, 1] + case["patch"][1],
                         starts[:, 2], starts[:, 2] + case["patch"][2]], axis=1).astype(np.int64)
        assert zlib.crc32(allc.tobytes()) == case["crc"]


def test_grid_docstring_counts():
    # reference docstring known answers: data_3D_manipulation.py:425-447, data_2D_manipulation.py:124-171
    n = lambda *a: port_stitch.crop_grid(*a)[1].shape[0]
    assert n((165, 768, 1024), (80, 80, 80), (0.5, 0.5, 0.5), (0, 0, 0)) == 2600
    assert n((165, 768, 1024), (80, 80, 80), (0, 0, 0), (0, 0, 0)) == 390
    assert n((512, 512, 512), (128, 128, 128), (0.25,) * 3, (0, 0, 0)) == 216
    # 2D: 165 images of 768x1024, 256x256 crops
    assert 165 * n((768, 1024), (256, 256), (0, 0), (0, 0)) == 1980
    assert 165 * n((768, 1024), (256, 256), (0.5, 0.5), (0, 0)) == 7920


def test_grid_known_answers_2d():
    g = _grids()
    for case in g["2d"]:
        plans, starts = port_stitch.crop_grid(case["shape"], case["patch"], case["overlap"], case["padding"])
        assert starts.shape[0] == case["n"]
        allc = np.stack([starts[:, 0], starts[:, 0] + case["patch"][0], starts[:, 1], starts[:, 1] + case["patch"][1]],
                        axis=1).astype(np.int64)
        assert zlib.crc32(allc.tobytes()) == case["crc"]


def test_float_truncation():
    for t in _grids()["trunc"]:
        assert port_stitch.AxisPlan(10 * t["P"], t["P"], 0, t["ov"]).step <= t["step"]
        assert int(t["P"] * (1 - t["ov"])) == t["step"]
    assert int(50 * (1 - 0.9)) == 4


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "stitch_s3d_*.npz"))))
def test_stitch_3d(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    patches, starts = port_stitch.crop_3d(z["vol"], meta["patch"], meta["overlap"], meta["padding"], meta["pad_type"])
    assert np.array_equal(starts, z["starts"])
    assert zlib.crc32(np.ascontiguousarray(patches).tobytes()) == int(z["patches_crc"])
    merged = port_stitch.merge_3d(z["pred"], meta["vshape"], meta["overlap"], meta["padding"])
    assert merged.dtype == np.float32 and np.array_equal(merged, z["merged"])
    merged16 = port_stitch.merge_3d(z["pred"].astype(np.float16), meta["vshape"], meta["overlap"], meta["padding"])
    assert merged16.dtype == np.float16 and np.array_equal(merged16, z["merged16"])
    rt = port_stitch.merge_3d(patches, meta["vshape"], meta["overlap"], meta["padding"])
    assert np.array_equal(rt, z["roundtrip"])
    assert np.abs(rt - z["vol"]).max() < 2e-6       # crop -> merge is the identity


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "stitch_s2d_*.npz"))))
def test_stitch_2d(path):
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    patches, starts = port_stitch.crop_2d(z["vol"], meta["patch"], meta["overlap"], meta["padding"], meta["pad_type"])
    assert np.array_equal(starts, z["starts"])
    assert zlib.crc32(np.ascontiguousarray(patches).tobytes()) == int(z["patches_crc"])
    merged = port_stitch.merge_2d(z["pred"], meta["vshape"], meta["overlap"], meta["padding"])
    assert np.array_equal(merged, z["merged"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "model_*.npz"))))
def test_model_port(path):
    z = np.load(path)
    kw = json.loads(str(z["kwargs_json"]))
    arch = str(z["arch"])
    sd = {k[3:]: torch.from_numpy(z[k].copy()).requires_grad_(z[k].dtype == np.float32 and "running_" not in k)
          for k in z.files if k.startswith("sd.")}
    x = torch.from_numpy(z["x"]).requires_grad_(True)
    y = port_models.forward(arch, sd, x, training=True, **kw)
    ref = torch.from_numpy(z["y"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-5 * max(1.0, ref.abs().max().item())
    (y * torch.from_numpy(z["gy"])).sum().backward()
    gx = torch.from_numpy(z["gx"])
    assert (x.grad - gx).abs().max() <= 1e-4 * max(1.0, gx.abs().max().item())
    # parameters in front of a norm layer have mathematically-zero gradients (pure rounding noise), so the
    # tolerance is tied to the largest gradient of the model, not to each tensor's own magnitude
    scale = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("grad."))
    for k in z.files:
        if k.startswith("grad."):
            g = torch.from_numpy(z[k])
            got = sd[k[5:]].grad
            assert got is not None, k
            assert (got - g).abs().max() <= 1e-4 * max(1.0, scale), k
    # BatchNorm fixtures: running statistics after the training step and the eval-mode forward on them
    after = [k for k in z.files if k.startswith("sd_after.") and "running_" in k]
    for k in after:
        assert (sd[k[9:]].detach() - torch.from_numpy(z[k])).abs().max() <= 1e-5, k
    if "y_eval" in z.files:
        with torch.no_grad():
            ye = port_models.forward(arch, {k: v.detach() for k, v in sd.items()}, x.detach(), training=False, **kw)
        ref_e = torch.from_numpy(z["y_eval"])
        assert (ye - ref_e).abs().max() <= 1e-5 * max(1.0, ref_e.abs().max().item())


# ------------------------------------------------------------------------------------- by-chunks tile grid (row a18)
def _chunk_cases():
    with open(os.path.join(GOLDEN, "chunks.json")) as f:
        return json.load(f)


def chunk_volume(case):
    """The seeded volume oracle/make_golden.py fed to the reference (not stored in the fixture)."""
    rng = np.random.default_rng(zlib.crc32(case["name"].encode()))
    return rng.standard_normal(tuple(case["shape"])).astype(np.float32)


def oracle_chunk_coords(g):
    from oracle import port_chunks  # noqa: F401
    rows = np.zeros((g.total_vols, 21), dtype=np.int64)
    for vid in range(g.total_vols):
        z, y, x, ext, real = g.patch_coords(vid)
        _, info = g.pad_to_add((z, y, x), ext)
        rows[vid] = [z, y, x] + ext + real + [v for ax in info for v in ax]
    return rows


@pytest.mark.parametrize("case", _chunk_cases(), ids=lambda c: c["name"])
def test_chunk_grid_oracle_matches_reference(case):
    from oracle import port_chunks
    g = port_chunks.ChunkGrid(case["shape"], case["crop"], case["padding"], case["z_start"], case["z_end"], case["patches_per_tile"])
    assert [g.step_z, g.step_y, g.step_x] == case["steps"]
    assert [g.vols_per_z, g.vols_per_y, g.vols_per_x] == case["vols"]
    assert [g.z_vol_start, g.z_vol_end] == case["z_vol"] and g.total_vols == case["total_vols"]
    assert len(g.tile_ids) == case["n_tiles"] and [g.tiles_per_z, g.tiles_per_y, g.tiles_per_x] == case["tiles"]
    assert zlib.crc32(np.asarray(g.tile_ids, np.int64).tobytes()) == case["tile_ids_crc"]
    assert g.tile_coords(g.tile_ids[-1]) == case["tile0"]
    rows = oracle_chunk_coords(g)
    fx = np.load(os.path.join(GOLDEN, f"chunks_{case['name']}.npz"))
    assert np.array_equal(rows, fx["coords"]) and zlib.crc32(rows.tobytes()) == case["coords_crc"]
    for key, d in case["deal"].items():
        world, rank = (int(v) for v in key.split(":"))
        vids = g.rank_patches(world, rank)
        assert len(vids) == d["n"] and vids[:6] == d["head"]
        assert zlib.crc32(np.asarray(vids, np.int64).tobytes()) == d["crc"]
        assert list(g.rank_workload(1, world, rank)) == d["workload"]


@pytest.mark.parametrize("case", [c for c in _chunk_cases() if c["arrays"]], ids=lambda c: c["name"])
def test_chunk_extract_insert_oracle_matches_reference(case):
    from oracle import port_chunks
    vol = chunk_volume(case)
    assert zlib.crc32(vol.tobytes()) == case["vol_crc"]
    g = port_chunks.ChunkGrid(case["shape"], case["crop"], case["padding"], case["z_start"], case["z_end"], case["patches_per_tile"])
    fx = np.load(os.path.join(GOLDEN, f"chunks_{case['name']}.npz"))
    for i, vid in enumerate(fx["pick"].tolist()):
        assert np.array_equal(g.extract(vol, vid)[0], fx["samples"][i])
    allp = np.stack([g.extract(vol, v)[0] for v in range(g.total_vols)])
    assert zlib.crc32(allp.tobytes()) == case["samples_crc"]
    # identity model: extract -> strip -> insert reproduces the volume, on any number of ranks
    for world in (1, 3):
        out = np.zeros(vol.shape, np.float32)
        for rank in range(world):
            port_chunks.predict_by_chunks(vol, g, lambda b: b, vol.shape[-1], world, rank, out=out)
        assert zlib.crc32(out.tobytes()) == case["out_crc"] and np.array_equal(out, vol)


# ------------------------------------------------------------------------------- test-time augmentation
def tta_image(shape, name):
    """The seeded image oracle/make_golden.py fed to the reference (not stored in the fixture)."""
    return np.random.default_rng(sum(map(ord, "tta" + name))).standard_normal(shape).astype(np.float32)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "tta_*.npz"))))
def test_tta_oracle_matches_reference(path):
    """oracle/port_tta.py against the reference's ensemble_predictions: orientation order and the ensemble, bit-exact."""
    from oracle import port_tta
    z = np.load(path)
    meta = json.loads(str(z["meta"]))
    name = os.path.basename(path)[4:-4]
    orients = port_tta.orientations(meta["ndim"], "full" if meta["group"] == "auto" else meta["group"])
    assert [list(p) for p, _ in orients] == z["perms"].tolist()
    assert [list(s) for _, s in orients] == z["signs"].tolist()
    out = port_tta.ensemble_predictions(tta_image(tuple(meta["shape"]), name), port_tta.toy_pred_func, meta["ndim"],
                                        meta["batch"], meta["mode"], meta["group"])
    assert out.dtype == np.float32 and np.array_equal(out, z["out"])


def test_tta_orientation_group_host_mirror():
    """biapy_b200's AxisTransform / build_axis_transform_group (host bookkeeping, no GPU) against the oracle enumeration."""
    from biapy_b200.data.post_processing.tta import AxisTransform, build_axis_transform_group
    from oracle import port_tta
    for nd in (2, 3):
        for level in ("full", "flips", "none"):
            got = build_axis_transform_group(nd, level)
            assert [(t.perm, t.sign) for t in got] == port_tta.orientations(nd, level)
            assert got[0].is_identity
            for t in got:
                assert (t.inverse.perm, t.inverse.sign) == port_tta.inverse(t.perm, t.sign)
                assert t.inverse.inverse == t
    assert len(build_axis_transform_group(2)) == 8 and len(build_axis_transform_group(3)) == 16
    with pytest.raises(ValueError):
        AxisTransform((0, 0), (1, 1))
    with pytest.raises(ValueError):
        build_axis_transform_group(4)


def test_tta_oracle_passes_the_reference_unit_tests():
    """The scalar-path unit tests the reference ships for its TTA (``tests/test_tta_equivariance.py``), held against the oracle
    and the host mirror of the transforms: inverse round trips (``:198-206``), group sizes and "Z never moves" (``:209-219``),
    rot90 == numpy (``:222-225``), an identity model comes back unchanged (``:529-546``), non-square inputs keep their shape
    (``:522-526``), ``TEST.AUGMENTATION_GROUP`` sets the number of forward passes (``:572-590``)."""
    from biapy_b200.data.post_processing.tta import AxisTransform, build_axis_transform_group
    from oracle import port_tta
    rng = np.random.default_rng(0)
    for ndim in (2, 3):
        arr = rng.normal(size=(5, 6, 7)[:ndim] + (3,))
        for perm, sign in port_tta.orientations(ndim, "full"):
            back = port_tta.apply(port_tta.apply(arr, perm, sign), *port_tta.inverse(perm, sign))
            assert np.array_equal(back, arr), (perm, sign)
    assert [len(port_tta.orientations(*a)) for a in ((2, "full"), (2, "flips"), (3, "full"), (3, "flips"), (3, "none"))] == [8, 4, 16, 8, 1]
    assert all(perm[0] == 0 for perm, _ in port_tta.orientations(3, "full"))
    assert port_tta.orientations(3, "full")[0] == ((0, 1, 2), (1, 1, 1))
    arr = rng.normal(size=(8, 8, 2))
    assert np.array_equal(port_tta.apply(arr, (1, 0), (-1, 1)), np.rot90(arr, 1, axes=(0, 1)))
    assert AxisTransform((1, 0), (-1, 1)) in build_axis_transform_group(2, "full")
    # identity model over the full group: the input comes back unchanged
    img = np.random.default_rng(1).normal(size=(32, 32, 1)).astype(np.float32)
    out = port_tta.ensemble_predictions(img, lambda b: b.astype(np.float32), 2, 8)
    assert out.shape == img.shape and np.allclose(out, img, atol=1e-5)
    # non-square input: padded for the rotations, cropped back
    img = np.random.default_rng(2).normal(size=(48, 64, 3)).astype(np.float32)
    out = port_tta.ensemble_predictions(img, lambda b: b.astype(np.float32), 2, 4)
    assert out.shape == (48, 64, 3) and np.allclose(out, img, atol=1e-5)
    for group, expected in (("none", 1), ("flips", 4), ("full", 8)):
        calls = []

        def pred_func(batch):
            calls.append(batch.shape[0])
            return batch.astype(np.float32)
        port_tta.ensemble_predictions(np.random.default_rng(2).normal(size=(16, 16, 1)).astype(np.float32), pred_func, 2, 1, "mean", group)
        assert sum(calls) == expected


# --------------------------------------------------------------------------- image normalisation at the ends
def test_norm_oracle_matches_reference():
    import copy
    from oracle import port_norm
    n = 0
    for img, mod, y_ref, u_ref, info_ref, key in port_norm.golden_cases(GOLDEN):
        y, info = port_norm.normalize_image(img.copy(), copy.deepcopy(mod))
        assert y.dtype == y_ref.dtype and np.array_equal(y, y_ref), key
        assert json.loads(json.dumps(info)) == info_ref, key
        u = port_norm.undo_image_norm(y.copy(), info)
        assert u.dtype == u_ref.dtype and np.array_equal(u, u_ref), key
        n += 1
    assert n == 32


def test_otsu_threshold_oracle_and_host_arithmetic():
    """oracle/port_norm.threshold_otsu (skimage's algorithm restated on np.histogram -- scikit-image itself is not installed, see
    its docstring) against the committed fixture, and the product's host half (`otsu_from_counts`: skimage's arithmetic on the
    counts the device histogram returns) against the oracle for the same histograms."""
    import numpy as np
    from biapy_b200.data.norm import otsu_from_counts
    from oracle import port_norm
    z = np.load(os.path.join(GOLDEN, "otsu_cases.npz"))
    for name, img in port_norm.otsu_cases().items():
        th = port_norm.threshold_otsu(img)
        assert th.dtype == np.float32 and th == z["th." + name], name
        if "counts." + name in z.files:
            counts, edges = np.histogram(img.reshape(-1), bins=256)
            assert np.array_equal(counts, z["counts." + name])
            got = otsu_from_counts(counts.astype(np.int64), np.linspace(img.min(), img.max(), 257, dtype=np.float32))
            assert got.dtype == np.float32 and got == th, name
    # hand-checkable case: two well separated clusters -> the threshold falls between them
    img = np.concatenate([np.full(100, 0.1, np.float32), np.full(50, 0.9, np.float32)])
    th = port_norm.threshold_otsu(img)
    assert 0.1 <= th < 0.9 and np.array_equal(port_norm.binarize(img, 2, None), (img > th).astype(np.uint8))

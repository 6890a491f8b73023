"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the REAL reference (container only).

Run:  python -m oracle.make_golden          (needs /root/reference; see oracle/ref_loader.py)

Every fixture stores seeded inputs, the reference-layout state_dict and the outputs of the unmodified
reference code (one documented GN patch, see ref_loader).  The fixtures travel to the GPU box, where
``/root/reference`` does not exist; tests compare both the oracle port and the CUDA engine against them.
Fixtures are kept small (few 100 KB each).
"""
from __future__ import annotations

import contextlib
import io
import json
import os

import numpy as np
import torch

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, arch, ctor kwargs, batch) -- small versions of the BASELINE configs + block-order / norm variants
MODEL_CASES = [
    ("resunet3d_gn_silu", "resunet", dict(image_shape=(16, 16, 16, 2), activation="silu", feature_maps=[16, 32],
                                           drop_values=[0, 0], normalization="gn", k_size=3, yx_down=[2], z_down=[2],
                                           isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1]), 2),
    ("resunet3d_gn_deep", "resunet", dict(image_shape=(16, 16, 16, 2), activation="silu", feature_maps=[8, 16, 32],
                                           drop_values=[0, 0, 0], normalization="gn", k_size=3, yx_down=[2, 2], z_down=[2, 2],
                                           isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1]), 1),
    ("resunet3d_preact_in", "resunet", dict(image_shape=(8, 16, 16, 1), activation="relu", feature_maps=[8, 16],
                                             drop_values=[0, 0], normalization="in", k_size=3, yx_down=[2], z_down=[1],
                                             isotropy=[False, True], larger_io=True, conv_layers=[2, 3], output_channels=[2],
                                             conv_block_order="norm_act_conv"), 1),
    ("unet2d_in_elu", "unet", dict(image_shape=(32, 32, 1), activation="elu", feature_maps=[8, 16, 32],
                                    drop_values=[0, 0, 0], normalization="in", k_size=3, yx_down=[2, 2], z_down=[2, 2],
                                    isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1]), 2),
    ("unet2d_gn_3ch", "unet", dict(image_shape=(32, 32, 3), activation="silu", feature_maps=[16, 32],
                                    drop_values=[0, 0], normalization="gn", k_size=3, yx_down=[2], z_down=[2],
                                    isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[2]), 2),
    ("attunet3d_in_elu", "attention_unet", dict(image_shape=(16, 16, 16, 1), activation="elu", feature_maps=[8, 16, 32],
                                                 drop_values=[0, 0, 0], normalization="in", k_size=3, yx_down=[2, 2], z_down=[2, 2],
                                                 isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1]), 2),
    ("unet3d_none_relu_upsampling", "unet", dict(image_shape=(8, 8, 8, 2), activation="relu", feature_maps=[8, 16],
                                                  drop_values=[0, 0], normalization="none", k_size=3, yx_down=[2], z_down=[2],
                                                  isotropy=[True] * 2, larger_io=False, conv_layers=[1, 2], output_channels=[1, 1],
                                                  output_channel_info=["A", "B"], upsample_layer="upsampling"), 1),
    ("resunet3d_none_relu", "resunet", dict(image_shape=(8, 8, 8, 1), activation="relu", feature_maps=[8, 16],
                                             drop_values=[0, 0], normalization="none", k_size=3, yx_down=[2], z_down=[2],
                                             isotropy=[True] * 2, larger_io=False, conv_layers=[2, 2], output_channels=[1]), 1),
    # BatchNorm (BiaPy's default normalisation): train-mode forward/backward with batch statistics, the running
    # statistics after that step ("sd_after.*") and an eval-mode forward on them ("y_eval")
    ("unet3d_bn_relu", "unet", dict(image_shape=(8, 16, 16, 1), activation="relu", feature_maps=[8, 16],
                                     drop_values=[0, 0], normalization="bn", k_size=3, yx_down=[2], z_down=[2],
                                     isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[1]), 3),
    ("resunet2d_bn_silu", "resunet", dict(image_shape=(32, 32, 2), activation="silu", feature_maps=[16, 32],
                                           drop_values=[0, 0], normalization="bn", k_size=3, yx_down=[2], z_down=[2],
                                           isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2, output_channels=[2]), 2),
]

STITCH_3D = [
    # (name, vol shape, patch, overlap, padding, pad_type)
    ("s3d_a", (24, 28, 20, 2), (12, 12, 12, 2), (0.25, 0.25, 0.25), (0, 0, 0), "reflect"),
    ("s3d_b", (21, 25, 19, 1), (8, 12, 10, 1), (0.1, 0.3, 0.5), (1, 0, 2), "reflect"),
    ("s3d_c", (14, 16, 14, 1), (8, 10, 8, 1), (0, 0, 0), (2, 2, 2), "zeros"),
    ("s3d_dup", (12, 14, 12, 1), (12, 8, 12, 1), (0.25, 0.5, 0.25), (0, 0, 0), "reflect"),   # axis == patch, ov>0
    ("s3d_hi", (26, 8, 8, 1), (10, 8, 8, 1), (0.7, 0, 0), (0, 0, 0), "symmetric"),
]
STITCH_2D = [
    ("s2d_a", (2, 50, 60, 3), (24, 24, 3), (0.25, 0.25), (0, 0), "reflect"),
    ("s2d_b", (1, 44, 38, 1), (20, 16, 1), (0.5, 0.1), (4, 2), "reflect"),
]
# index-only known answers (no tensors): shape, patch, overlap, padding -> n patches + per-axis starts
GRID_CASES_3D = [
    ((512, 512, 512), (128, 128, 128), (0.25, 0.25, 0.25), (0, 0, 0)),
    ((165, 768, 1024), (80, 80, 80), (0.5, 0.5, 0.5), (0, 0, 0)),
    ((165, 768, 1024), (80, 80, 80), (0, 0, 0), (0, 0, 0)),
    ((165, 768, 1024), (80, 80, 80), (0, 0, 0), (10, 10, 10)),
    ((100, 120, 90), (40, 40, 40), (0, 0, 0), (0, 0, 0)),
    ((100, 120, 90), (40, 40, 40), (0, 0, 0), (6, 6, 6)),
    ((100, 120, 90), (40, 40, 40), (0.25, 0.25, 0.25), (6, 6, 6)),
    ((97, 113, 89), (32, 48, 40), (0.1, 0.3, 0.5), (4, 0, 8)),
    ((64, 64, 64), (64, 64, 64), (0.25, 0.25, 0.25), (0, 0, 0)),
    ((200, 150, 100), (50, 100, 96), (0.9, 0.7, 0.35), (0, 0, 0)),
]
GRID_CASES_2D = [
    ((2048, 2048), (512, 512), (0.25, 0.25), (0, 0)),
    ((768, 1024), (256, 256), (0, 0), (0, 0)),
    ((768, 1024), (256, 256), (0.5, 0.5), (0, 0)),
    ((768, 1024), (256, 256), (0, 0), (64, 64)),
    ((300, 500), (128, 96), (0.33, 0.8), (16, 8)),
]


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def make_models(R, only=None):
    for name, arch, kw, batch in MODEL_CASES:
        if only and name not in only:
            continue
        cls = {"unet": R.unet.U_Net, "resunet": R.resunet.ResUNet, "attention_unet": R.attention_unet.Attention_U_Net}[arch]
        torch.manual_seed(0)
        with _quiet():
            m = cls(**kw)
        # non-trivial affine / bias values so that every parameter matters
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if p.ndim == 1:
                    p.add_(0.2 * torch.randn(p.shape, generator=g))
        m.train()  # exercise the training-mode graph
        sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}     # before the step (BatchNorm buffers change)
        shape = kw["image_shape"]
        x = torch.randn((batch, shape[-1]) + tuple(shape[:-1]), generator=g)
        x.requires_grad_(True)
        y = m(x)
        gy = torch.randn(y.shape, generator=g)
        (y * gy).sum().backward()
        out = {"x": x.detach().numpy(), "y": y.detach().numpy(), "gy": gy.numpy(), "gx": x.grad.numpy(),
               "kwargs_json": np.array(json.dumps(kw)), "arch": np.array(arch)}
        for k, v in sd0.items():
            out["sd." + k] = v.numpy()
        for k, p in m.named_parameters():
            out["grad." + k] = p.grad.numpy()
        if kw["normalization"] in ("bn", "sync_bn"):
            for k, v in m.state_dict().items():
                if "running_" in k or "num_batches_tracked" in k:
                    out["sd_after." + k] = v.numpy()
            m.eval()
            with torch.no_grad():
                out["y_eval"] = m(x.detach()).numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **out)
        print("model", name, tuple(y.shape), "params", sum(p.numel() for p in m.parameters()))


def make_stitch(R):
    rng = np.random.default_rng(7)
    for name, vshape, patch, ov, pad, pad_type in STITCH_3D:
        vol = rng.standard_normal(vshape).astype(np.float32)
        patches, coords = R.d3.crop_3D_data_with_overlap(vol, patch, overlap=ov, padding=pad, verbose=False, pad_type=pad_type)
        starts = np.array([[c.z_start, c.y_start, c.x_start] for c in coords], dtype=np.int64)
        pred = patches * 0.5 + rng.standard_normal(patches.shape).astype(np.float32)
        merged = R.d3.merge_3D_data_with_overlap(pred, vshape, overlap=ov, padding=pad, verbose=False)
        merged16 = R.d3.merge_3D_data_with_overlap(pred.astype(np.float16), vshape, overlap=ov, padding=pad, verbose=False)
        roundtrip = R.d3.merge_3D_data_with_overlap(patches, vshape, overlap=ov, padding=pad, verbose=False)
        np.savez_compressed(os.path.join(OUT, f"stitch_{name}.npz"), vol=vol, patches_crc=np.array(_crc(patches)),
                            starts=starts, pred=pred, merged=merged, merged16=merged16, roundtrip=roundtrip,
                            meta=np.array(json.dumps(dict(vshape=vshape, patch=patch, overlap=ov, padding=pad, pad_type=pad_type))))
        print("stitch", name, patches.shape, "roundtrip err", float(np.abs(roundtrip - vol).max()))
    for name, dshape, patch, ov, pad, pad_type in STITCH_2D:
        img = rng.standard_normal(dshape).astype(np.float32)
        patches, coords = R.d2.crop_data_with_overlap(img, patch, overlap=ov, padding=pad, verbose=False, pad_type=pad_type)
        starts = np.array([[c.y_start, c.x_start] for c in coords], dtype=np.int64)
        pred = patches * 0.5 + rng.standard_normal(patches.shape).astype(np.float32)
        merged = R.d2.merge_data_with_overlap(pred, dshape, overlap=ov, padding=pad, verbose=False)
        np.savez_compressed(os.path.join(OUT, f"stitch_{name}.npz"), vol=img, patches_crc=np.array(_crc(patches)),
                            starts=starts, pred=pred, merged=merged,
                            meta=np.array(json.dumps(dict(vshape=dshape, patch=patch, overlap=ov, padding=pad, pad_type=pad_type))))
        print("stitch", name, patches.shape)


def _crc(a: np.ndarray) -> int:
    import zlib
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def make_grids(R):
    grids = {"3d": [], "2d": []}
    for shape, patch, ov, pad in GRID_CASES_3D:
        dummy = np.broadcast_to(np.zeros((1,), np.uint8), shape + (1,))
        coords = R.d3.crop_3D_data_with_overlap(dummy, patch + (1,), overlap=ov, padding=pad, verbose=False, load_data=False)
        z = sorted({c.z_start for c in coords}); y = sorted({c.y_start for c in coords}); x = sorted({c.x_start for c in coords})
        first = [[c.z_start, c.y_start, c.x_start] for c in coords[:3]] + [[c.z_start, c.y_start, c.x_start] for c in coords[-3:]]
        import zlib
        allc = np.array([[c.z_start, c.z_end, c.y_start, c.y_end, c.x_start, c.x_end] for c in coords], dtype=np.int64)
        grids["3d"].append(dict(shape=shape, patch=patch, overlap=ov, padding=pad, n=len(coords), z=z, y=y, x=x,
                                edge_coords=first, crc=zlib.crc32(allc.tobytes())))
        print("grid3d", shape, patch, ov, pad, "->", len(coords))
    for shape, patch, ov, pad in GRID_CASES_2D:
        dummy = np.broadcast_to(np.zeros((1,), np.uint8), (1,) + shape + (1,))
        coords = R.d2.crop_data_with_overlap(dummy, patch + (1,), overlap=ov, padding=pad, verbose=False, load_data=False)
        y = sorted({c.y_start for c in coords}); x = sorted({c.x_start for c in coords})
        import zlib
        allc = np.array([[c.y_start, c.y_end, c.x_start, c.x_end] for c in coords], dtype=np.int64)
        grids["2d"].append(dict(shape=shape, patch=patch, overlap=ov, padding=pad, n=len(coords), y=y, x=x,
                                crc=zlib.crc32(allc.tobytes())))
        print("grid2d", shape, patch, ov, pad, "->", len(coords))
    # float-truncation probes (SURVEY 8a addendum)
    grids["trunc"] = [dict(P=P, ov=ov, step=int(P * (1 - ov))) for P, ov in [(50, 0.9), (100, 0.7), (96, 0.35), (128, 0.25), (80, 0.5)]]
    with open(os.path.join(OUT, "grids.json"), "w") as f:
        json.dump(grids, f, indent=0)


# by-chunks tile grid: (name, volume shape ZYXC, crop ZYXC, padding, z_start, z_end, patches_per_tile, keep arrays)
CHUNK_CASES = [
    ("c_a", (50, 70, 60, 2), (32, 32, 32, 2), (4, 6, 8), -1, -1, (1, 1, 1), True),
    ("c_b", (40, 33, 47, 1), (16, 24, 16, 1), (0, 0, 0), -1, -1, (2, 1, 3), True),     # no padding, ragged last tiles, grouped tiles
    ("c_c", (37, 41, 35, 1), (16, 16, 16, 1), (7, 7, 7), -1, -1, (1, 1, 1), True),      # step 2: last window shorter than the reflect reach
    ("c_d", (96, 64, 64, 1), (32, 32, 32, 1), (8, 8, 8), 20, 70, (1, 2, 2), False),      # Z range
    ("c_e", (512, 512, 512, 1), (128, 128, 128, 1), (16, 16, 16), -1, -1, (1, 1, 1), False),   # cfg-3 volume, by-chunks semantics
]


def make_chunks():
    """Fixtures from the reference's own `chunked_test_pair_data_generator` (grid, coordinates, extracted + reflect-padded
    samples, strip + insert with an identity model, DistributedSampler dealing)."""
    import zlib
    from torch.utils.data import DistributedSampler
    C = ref_loader.load_chunked()
    meta = []
    for name, shape, crop, pad, z0, z1, ppt, keep in CHUNK_CASES:
        rng = np.random.default_rng(zlib.crc32(name.encode()))
        small = keep or int(np.prod(shape)) <= 1 << 22
        vol = rng.standard_normal(shape).astype(np.float32) if small else np.zeros((1, 1, 1, 1), np.float32)
        X = vol if small else np.lib.stride_tricks.as_strided(vol, shape=shape, strides=(0, 0, 0, 0))
        with _quiet():
            g = C.cls(dict(X=X, Y=None, X_filename=name + ".zarr", X_dir="."), {}, "ZYXC", "ZYXC", crop, pad, out_dir="/tmp/_b200_golden",
                      z_start=z0, z_end=z1, patches_per_tile=ppt)
        coords = np.zeros((g.total_vols, 3 + 6 + 6 + 6), dtype=np.int64)
        out = np.zeros(shape, np.float32) if small else None
        samples = []
        for vid in range(g.total_vols):
            z, y, x, pe, pr = g._patch_coords(vid)
            if small:
                data, padinfo = g.extract_and_prepare_sample(z, y, x, pe)
            else:                       # coordinates + padding bookkeeping only (same arithmetic, no 512^3 copies)
                data, padinfo = None, None
                px = []
                for a, (s, e) in enumerate([(pe.z_start, pe.z_end), (pe.y_start, pe.y_end), (pe.x_start, pe.x_end)]):
                    pos = (z, y, x)[a]
                    step = (g.step_z, g.step_y, g.step_x)[a]
                    left = abs(pos * step - pad[a]) if pos * step - pad[a] < 0 else 0
                    right = crop[a] - (e - s) - left
                    px.append([max(left, pad[a]), max(right, pad[a])])
                padinfo = px
            coords[vid] = [z, y, x, pe.z_start, pe.z_end, pe.y_start, pe.y_end, pe.x_start, pe.x_end,
                           pr.z_start, pr.z_end, pr.y_start, pr.y_end, pr.x_start, pr.x_end,
                           padinfo[0][0], padinfo[0][1], padinfo[1][0], padinfo[1][1], padinfo[2][0], padinfo[2][1]]
            if small:
                samples.append(data)
                # base_workflow.py:2606-2614 (strip) + insert_patch_in_efficient_file (replace), identity "model"
                raw = data[padinfo[0][0]: data.shape[0] - padinfo[0][1], padinfo[1][0]: data.shape[1] - padinfo[1][1],
                           padinfo[2][0]: data.shape[2] - padinfo[2][1]]
                C.d3.insert_patch_in_efficient_file(out, raw, pr, data_axes_order="ZYXC", patch_axes_order="ZYXC", mode="replace")
        deal = {}
        for world in (1, 2, 3, 8):
            for rank in range(world):
                smp = DistributedSampler(g.tile_ids, num_replicas=world, rank=rank, shuffle=False)
                vids = [v for i in smp for v in g.patches_of_tile[g.tile_ids[int(i)]]]
                deal[f"{world}:{rank}"] = dict(crc=zlib.crc32(np.asarray(vids, np.int64).tobytes()), n=len(vids), head=vids[:6],
                                               workload=list(g.rank_workload(1, world, rank)))
        m = dict(name=name, shape=list(shape), crop=list(crop), padding=list(pad), z_start=z0, z_end=z1, patches_per_tile=list(ppt),
                 steps=[g.step_z, g.step_y, g.step_x], vols=[g.vols_per_z, g.vols_per_y, g.vols_per_x],
                 z_vol=[g.z_vol_start, g.z_vol_end], total_vols=g.total_vols, n_tiles=len(g.tile_ids),
                 tiles=[g.tiles_per_z, g.tiles_per_y, g.tiles_per_x], tile_ids_crc=zlib.crc32(np.asarray(g.tile_ids, np.int64).tobytes()),
                 tile0=[getattr(g.tile_coords(g.tile_ids[-1]), k) for k in ("z_start", "z_end", "y_start", "y_end", "x_start", "x_end")],
                 coords_crc=zlib.crc32(coords.tobytes()), deal=deal, arrays=bool(keep))
        if small:
            m["roundtrip_equal"] = bool(np.array_equal(out[g.z_vol_start * g.step_z: min(g.z_vol_end * g.step_z, shape[0])],
                                                       vol[g.z_vol_start * g.step_z: min(g.z_vol_end * g.step_z, shape[0])]))
            m["samples_crc"] = zlib.crc32(np.stack(samples).tobytes())
        if keep:
            pick = sorted({0, g.total_vols // 2, g.total_vols - 1, min(g.total_vols - 1, g.vols_per_x * g.vols_per_y - 1)})
            # the volume is NOT stored: tests regenerate it from the same seeded generator (vol_crc pins it)
            m["vol_crc"] = zlib.crc32(vol.tobytes())
            m["out_crc"] = zlib.crc32(out.tobytes())
            np.savez_compressed(os.path.join(OUT, f"chunks_{name}.npz"), coords=coords, pick=np.asarray(pick),
                                samples=np.stack([samples[i] for i in pick]))
        else:
            np.savez_compressed(os.path.join(OUT, f"chunks_{name}.npz"), coords=coords)
        meta.append(m)
        print("chunks", name, shape, crop, pad, "->", g.total_vols, "patches,", len(g.tile_ids), "tiles", m.get("roundtrip_equal"))
    with open(os.path.join(OUT, "chunks.json"), "w") as f:
        json.dump(meta, f, indent=0)


TTA_CASES = [
    # name, image shape (spatial..., C), ndim, mode, group, batch
    ("a", (6, 9, 7, 2), 3, "mean", "auto", 3),        # y != x: front padding (reflect) before the swaps
    ("b", (12, 9, 1), 2, "mean", "auto", 5),
    ("c", (5, 8, 8, 3), 3, "max", "full", 16),        # square in-plane: no padding
    ("d", (4, 3, 9, 1), 3, "min", "auto", 1),         # pad 6 >= dim 3: reflect degrades to edge (np.pad rule)
    ("e", (10, 7, 2), 2, "mean", "flips", 2),
    ("f", (3, 5, 6, 1), 3, "mean", "none", 4),
]


def tta_image(shape, name):
    return np.random.default_rng(sum(map(ord, "tta" + name))).standard_normal(shape).astype(np.float32)


def make_tta():
    """Scalar test-time augmentation: the reference's own `ensemble_predictions` driven by the fixed toy
    network of oracle/port_tta.py.  Stores the ensemble and the orientation list the reference enumerated."""
    from . import port_tta
    R = ref_loader.load_tta()
    for name, shape, nd, mode, group, bs in TTA_CASES:
        img = tta_image(shape, name)
        ao = {2: (0, 3, 1, 2), 3: (0, 4, 1, 2, 3)}[nd]
        aob = {2: (0, 2, 3, 1), 3: (0, 2, 3, 4, 1)}[nd]
        out = R.ensemble_predictions(img, lambda b: torch.from_numpy(port_tta.toy_pred_func(b)).permute(ao), aob, ao,
                                     torch.device("cpu"), nd, batch_size_value=bs, mode=mode, group=group)
        out = out.permute(aob)[0].numpy()
        grp = R.tta.build_axis_transform_group(nd, level=("full" if group == "auto" else group))
        np.savez_compressed(os.path.join(OUT, f"tta_{name}.npz"), out=out,
                            perms=np.array([t.perm for t in grp], np.int32), signs=np.array([t.sign for t in grp], np.int32),
                            meta=np.array(json.dumps(dict(shape=shape, ndim=nd, mode=mode, group=group, batch=bs))))
        print("tta", name, shape, mode, group, "->", out.shape, len(grp), "orientations")


NORM_MODULES = [
    dict(type="div", percentile_clip=False, out_dtype="float32"),
    dict(type="scale_range", percentile_clip=False, out_dtype="float32"),
    dict(type="zero_mean_unit_variance", percentile_clip=False, out_dtype="float32"),
    dict(type="zero_mean_unit_variance", percentile_clip=False, out_dtype="float32", mean=[0.5], std=[2.0]),
    dict(type="scale_range", percentile_clip=True, out_dtype="float32", lower_bound_val=[10.0], upper_bound_val=[200.0],
         per_lower_bound=-1, per_upper_bound=-1),
    dict(type="zero_mean_unit_variance", percentile_clip=True, out_dtype="float32", lower_bound_val=[-1.5], upper_bound_val=[2.5],
         per_lower_bound=-1, per_upper_bound=-1),
    # bounds from the data: np.percentile of each channel (norm.py:445-466)
    dict(type="scale_range", percentile_clip=True, out_dtype="float32", per_lower_bound=2.0, per_upper_bound=97.5),
    dict(type="zero_mean_unit_variance", percentile_clip=True, out_dtype="float32", per_lower_bound=0.7, per_upper_bound=-1,
         upper_bound_val=[150.0]),
]


def norm_images():
    rng = np.random.default_rng(2026)
    a = rng.integers(0, 256, (5, 9, 7, 2)).astype(np.uint8)
    b = rng.integers(0, 4000, (6, 8, 3)).astype(np.uint16)
    c = (rng.standard_normal((4, 6, 6, 2)) * 3 + 1).astype(np.float32)
    d = np.concatenate([rng.integers(0, 2, (4, 6, 6, 1)).astype(np.float32), c[..., :1]], -1)      # one binary channel
    return dict(a=a, b=b, c=c, d=d)


def make_norm():
    """Image normalisation at the ends: the reference's normalize_image / undo_image_norm on seeded images."""
    import copy
    R = ref_loader.load_norm()
    out = {}
    meta = []
    for name, img in norm_images().items():
        for i, m in enumerate(NORM_MODULES):
            y, info = R.normalize_image(img.copy(), copy.deepcopy(m))
            u = R.undo_image_norm(y.copy(), info)
            out[f"{name}{i}_y"], out[f"{name}{i}_u"] = y, u
            meta.append(dict(image=name, module=i, info=info))
    np.savez_compressed(os.path.join(OUT, "norm_cases.npz"), meta=np.array(json.dumps(meta)), **out)
    print("norm:", len(meta), "cases")


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_loader.load()
    import sys
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only == ["chunks"]:        # python -m oracle.make_golden chunks
        make_chunks()
        return
    if only == ["tta"]:           # python -m oracle.make_golden tta
        make_tta()
        return
    if only == ["norm"]:          # python -m oracle.make_golden norm
        make_norm()
        return
    if only:                      # add fixtures without rewriting the committed ones: python -m oracle.make_golden <model name>...
        make_models(R, only)
        return
    make_grids(R)
    make_stitch(R)
    make_models(R)
    make_chunks()
    make_tta()
    make_norm()
    with open(os.path.join(OUT, "PROVENANCE.txt"), "w") as f:
        f.write("generated by oracle/make_golden.py from /root/reference (BiaPy 3.7.0 @ 29539acd), "
                f"torch {torch.__version__}, numpy {np.__version__}; GN call patched as documented in oracle/ref_loader.py\n")


if __name__ == "__main__":
    main()

"""GPU parity of the by-chunks tile path through the C ABI: planner + extract / insert kernels against the
golden vectors produced by the reference's own `chunked_test_pair_data_generator` and against the CPU oracle."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from oracle import port_chunks

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "chunks.json")))


def _vol(case):
    rng = np.random.default_rng(zlib.crc32(case["name"].encode()))
    return rng.standard_normal(tuple(case["shape"])).astype(np.float32)


def _gen(case, vol):
    from biapy_b200.data.generators.chunked_test_pair_data_generator import chunked_test_pair_data_generator
    return chunked_test_pair_data_generator(dict(X=vol, Y=None, X_filename=case["name"]), None, "ZYXC", "ZYXC", tuple(case["crop"]),
                                            tuple(case["padding"]), z_start=case["z_start"], z_end=case["z_end"],
                                            patches_per_tile=tuple(case["patches_per_tile"]))


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["arrays"]], ids=lambda c: c["name"])
def test_chunk_extract_insert_bit_exact(case):
    vol = _vol(case)
    g = _gen(case, vol)
    fx = np.load(os.path.join(GOLDEN, f"chunks_{case['name']}.npz"))
    assert g.total_vols == case["total_vols"] and len(g.tile_ids) == case["n_tiles"]
    # single-tile API with the reference's signature
    for i, vid in enumerate(fx["pick"].tolist()):
        z, y, x, pe, pr = g._patch_coords(vid)
        data, pad = g.extract_and_prepare_sample(z, y, x, pe)
        assert np.array_equal(data, fx["samples"][i])
        assert [v for ax in pad[:3] for v in ax] == fx["coords"][vid, 15:21].tolist()
    # batched extraction of every tile: checksum of all samples as the reference produced them
    ids = list(range(g.total_vols))
    crc, out = 0, None
    for k in range(0, len(ids), 512):
        xb, pads, coords = g.extract_batch(ids[k:k + 512])
        crc = zlib.crc32(xb.cpu().numpy().tobytes(), crc)
        out = g.insert_batch(xb, pads, coords)                   # identity model
    assert crc == case["samples_crc"]
    res = out.cpu().numpy()
    assert zlib.crc32(res.tobytes()) == case["out_crc"] and np.array_equal(res, vol)


@pytest.mark.gpu
def test_chunk_fp16_volume_and_add_mode():
    case = CASES[0]
    vol = _vol(case).astype(np.float16)
    g = _gen(case, vol)
    og = port_chunks.ChunkGrid(case["shape"], case["crop"], case["padding"])
    ids = [0, 5, g.total_vols - 1]
    xb, pads, coords = g.extract_batch(ids)
    for j, vid in enumerate(ids):
        assert np.array_equal(xb[j].cpu().numpy(), og.extract(vol, vid)[0])
    out = torch.ones((g.z_dim, g.y_dim, g.x_dim, vol.shape[-1]), dtype=torch.float32, device="cuda")
    g.insert_batch(xb, pads, coords, mode="add", out=out)
    ref = np.ones(vol.shape, np.float32)
    for vid in ids:
        p, info, real = og.extract(vol, vid)
        port_chunks.strip_and_insert(ref, p.astype(np.float32), info, real, mode="add")
    assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.gpu
def test_predict_by_chunks_matches_oracle_loop():
    """Whole loop with a real network: tiles dealt to 3 'ranks' (run one after the other into the same output) == the
    oracle's loop driving the same engine forward patch by patch."""
    import contextlib, io
    from biapy_b200.engine.inference import predict_by_chunks
    from biapy_b200.models.unet import U_Net
    torch.manual_seed(3)
    with contextlib.redirect_stdout(io.StringIO()):
        model = U_Net(image_shape=(16, 16, 16, 1), activation="relu", feature_maps=[8, 16], drop_values=[0, 0], normalization="none",
                      k_size=3, yx_down=[2], z_down=[2], isotropy=[True] * 2, larger_io=False, conv_layers=[2] * 2,
                      output_channels=[1]).cuda().set_engine(dtype=torch.float32).eval()
    rng = np.random.default_rng(5)
    vol = rng.standard_normal((40, 36, 44, 1)).astype(np.float32)
    out = None
    for rank in range(3):
        out = predict_by_chunks(model, torch.from_numpy(vol).cuda(), (16, 16, 16, 1), padding=(2, 3, 4), batch_size=3,
                                head_activations=["ce_sigmoid"], rank=rank, world=3, out=out, reduce=False)

    def fn(b):
        with torch.no_grad():
            y = model(torch.from_numpy(b).cuda().permute(0, 4, 1, 2, 3))
        return torch.sigmoid(y).permute(0, 2, 3, 4, 1).cpu().numpy()

    ref = port_chunks.predict_by_chunks(vol, port_chunks.ChunkGrid(vol.shape, (16, 16, 16, 1), (2, 3, 4)), fn, 1)
    assert np.abs(out.cpu().numpy() - ref).max() < 1e-5


@pytest.mark.gpu
def test_chunk_errors_mirror_reference():
    vol = np.zeros((20, 20, 20, 1), np.float32)
    from biapy_b200.data.generators.chunked_test_pair_data_generator import chunked_test_pair_data_generator as G
    with pytest.raises(ValueError, match="Z Axis problem"):
        G(dict(X=vol, Y=None), None, "ZYXC", "ZYXC", (32, 16, 16, 1), (0, 0, 0))
    with pytest.raises(ValueError, match="Padding"):
        G(dict(X=vol, Y=None), None, "ZYXC", "ZYXC", (16, 16, 16, 1), (8, 0, 0))


class _CountingLazy:
    """A lazy (Z, Y, X, C) volume as zarr / h5py hand it over: `.shape`, `.dtype`, slice reads only -- and a record of how much
    was read at once, to show that the volume is never materialised."""

    def __init__(self, arr):
        self._a, self.shape, self.dtype, self.max_read, self.reads = arr, arr.shape, arr.dtype, 0, 0

    def __getitem__(self, key):
        out = self._a[key]
        self.max_read = max(self.max_read, out.nbytes)
        self.reads += 1
        return out


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["arrays"]][:3], ids=lambda c: c["name"])
def test_lazy_volume_tiles_equal_the_resident_path(case):
    """Tiles extracted from a lazy array-like (one halo read per tile, reflect padding by the same kernel) are the bytes the
    resident-volume path -- pinned to the reference's generator above -- extracts; an identity 'model' written back tile by tile
    into a memmap reproduces the volume."""
    vol = _vol(case)
    res = _gen(case, vol)
    lazy = _CountingLazy(vol)
    g = _gen(case, lazy)
    assert g.lazy is lazy and g.total_vols == res.total_vols
    ids = list(range(g.total_vols))
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        out = np.lib.format.open_memmap(os.path.join(tmp, "out.npy"), mode="w+", dtype=np.float32, shape=vol.shape)
        for k in range(0, len(ids), 7):
            xa, pa, ca = g.extract_batch(ids[k:k + 7])
            xb, pb, cb = res.extract_batch(ids[k:k + 7])
            assert torch.equal(xa, xb) and pa == pb and ca == cb
            g.write_batch(xa, pa, ca, out)
        assert np.array_equal(np.asarray(out), vol)
    tile_bytes = int(np.prod(case["crop"][:3])) * vol.shape[3] * 4
    assert lazy.max_read <= tile_bytes and lazy.max_read < vol.nbytes        # one tile with its halo at a time


@pytest.mark.gpu
def test_streaming_by_chunks_inference_equals_the_resident_call(tmp_path):
    """predict_by_chunks on a numpy.memmap volume, written into a memmap prediction, against the same call on the resident
    volume (which tests/test_gpu_workflow.py and tools/dist_check.py hold to the oracle / to world 1)."""
    import contextlib
    import io
    from biapy_b200.engine.inference import predict_by_chunks
    from biapy_b200.models.resunet import ResUNet
    kw = dict(image_shape=(32, 32, 32, 2), activation="silu", feature_maps=[16, 32, 64], drop_values=[0, 0, 0], normalization="gn",
              k_size=3, yx_down=[2, 2], z_down=[2, 2], isotropy=[True] * 3, larger_io=False, conv_layers=[2] * 3, output_channels=[1])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ResUNet(**kw).cuda().set_engine(dtype=torch.float32).eval()
    vol = np.random.default_rng(3).standard_normal((72, 64, 80, 2)).astype(np.float32)
    np.save(tmp_path / "vol.npy", vol)
    lazy = np.load(tmp_path / "vol.npy", mmap_mode="r")
    out = np.lib.format.open_memmap(tmp_path / "pred.npy", mode="w+", dtype=np.float32, shape=(72, 64, 80, 1))
    ref = predict_by_chunks(m, vol, (32, 32, 32, 2), padding=(4, 4, 4), batch_size=3, head_activations=["ce_sigmoid"])
    got = predict_by_chunks(m, lazy, (32, 32, 32, 2), padding=(4, 4, 4), batch_size=3, head_activations=["ce_sigmoid"], out=out)
    assert got is out and np.array_equal(np.asarray(out), ref)
    with pytest.raises(ValueError):
        predict_by_chunks(m, lazy, (32, 32, 32, 2), padding=(4, 4, 4), batch_size=3, head_activations=["ce_sigmoid"])

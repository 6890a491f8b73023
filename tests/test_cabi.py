"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares, and the host
entry points (planner, spline window) are bit-exact against the oracle and the golden vectors."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from biapy_b200 import _lib
from biapy_b200.data import _stitch
from oracle import port_stitch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "biapy_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-zA-Z0-9_]+)\s*\(", hdr))
    declared -= {"b200_dtype", "b200_act", "b200_pad_mode", "b200_conv_impl"}
    lib = _lib.lib()
    assert _lib.MISSING == []
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/biapy_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in biapy_b200/_lib.py"
    assert lib.b200_version() >= 100


def test_planner_matches_golden_grids():
    g = json.load(open(os.path.join(GOLDEN, "grids.json")))
    for case in g["3d"]:
        axes = [_stitch.Axis(case["shape"][i], case["patch"][i], case["padding"][i], case["overlap"][i]) for i in range(3)]
        assert int(np.prod([a.n for a in axes])) == case["n"]
        for a, key in zip(axes, "zyx"):
            assert sorted(set(a.starts(0).tolist())) == case[key]
    for case in g["2d"]:
        axes = [_stitch.Axis(case["shape"][i], case["patch"][i], case["padding"][i], case["overlap"][i]) for i in range(2)]
        assert int(np.prod([a.n for a in axes])) == case["n"]
        for a, key in zip(axes, "yx"):
            assert sorted(set(a.starts(0).tolist())) == case[key]


def test_planner_matches_oracle_randomised():
    rng = np.random.default_rng(0)
    for _ in range(3000):
        patch = int(rng.integers(4, 200))
        pad = int(rng.integers(0, patch // 2))
        dim = int(rng.integers(patch, 5 * patch + 7))
        ov = float(rng.choice([0, 0.1, 0.25, 0.3, 0.35, 0.5, 0.7, 0.9, rng.random() * 0.95]))
        try:
            ref = port_stitch.AxisPlan(dim, patch, pad, ov)
        except ZeroDivisionError:
            with pytest.raises(ZeroDivisionError):
                _stitch.Axis(dim, patch, pad, ov)
            continue
        a = _stitch.Axis(dim, patch, pad, ov)
        assert (a.step, a.n, a.last, a.core, a.ov_px) == (ref.step, ref.n, ref.last, ref.core, ref.ov_px)
        assert a.starts(0).tolist() == [ref.crop_start(i) for i in range(ref.n)]
        assert a.starts(1).tolist() == [ref.merge_start(i) for i in range(ref.n)]


def test_float_truncation_cases():
    # SURVEY 8a addendum: int((P-2p)*(1-ov)) in IEEE double, same operation order
    assert _stitch.Axis(500, 50, 0, 0.9).c.step <= 4
    c = _lib.AxisPlan()
    assert _lib.lib().b200_plan_axis(5000, 50, 0, 0.9, C.byref(c)) == 0
    # raw step before the per-block adjustment is 4 (not 5): n = ceil(5000/4) = 1250
    assert c.n == 1250
    assert _lib.lib().b200_plan_axis(10000, 100, 0, 0.7, C.byref(c)) == 0 and c.n == 334      # step 30
    assert _lib.lib().b200_plan_axis(9600, 96, 0, 0.35, C.byref(c)) == 0 and c.n == 155       # step 62


def test_overlap_range_error():
    with pytest.raises(ValueError):
        _stitch.Axis(100, 10, 0, 1.0)
    with pytest.raises(ValueError):
        _stitch.Axis(100, 10, 0, -0.1)


def test_spline_window_bit_exact():
    for size in (2, 3, 8, 51, 64, 128, 129, 512):
        for ov in (0, 1, 2, 7, 19, 51, 64, 100, 300):
            out = np.empty(size, np.float32)
            _lib.call("b200_spline_window_1d", size, ov, out.ctypes.data_as(C.POINTER(C.c_float)))
            assert np.array_equal(out, port_stitch.spline_window_1d(size, ov))


def test_coordinates_only_mode_needs_no_gpu():
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap
    from biapy_b200.data.data_2D_manipulation import crop_data_with_overlap
    c = crop_3D_data_with_overlap(np.zeros((165, 768, 1024, 1), np.uint8), (80, 80, 80, 1), overlap=(0.5, 0.5, 0.5),
                                  load_data=False, verbose=False)
    assert len(c) == 2600 and (c[-1].z_start, c[-1].z_end, c[-1].x_end) == (85, 165, 1024)
    c = crop_data_with_overlap(np.zeros((165, 768, 1024, 1), np.uint8), (256, 256, 1), overlap=(0.5, 0.5), load_data=False,
                               verbose=False)
    assert len(c) == 7920 and not hasattr(c[0], "z_start")


def test_reference_error_messages():
    from biapy_b200.data.data_3D_manipulation import crop_3D_data_with_overlap, merge_3D_data_with_overlap
    with pytest.raises(ValueError, match="data expected to be 4 dimensional"):
        crop_3D_data_with_overlap(np.zeros((4, 4, 4)), (2, 2, 2, 1), verbose=False)
    with pytest.raises(ValueError, match="greater than"):
        crop_3D_data_with_overlap(np.zeros((4, 4, 4, 1)), (8, 2, 2, 1), verbose=False)
    with pytest.raises(ValueError, match="Padding"):
        crop_3D_data_with_overlap(np.zeros((8, 8, 8, 1)), (4, 4, 4, 1), padding=(2, 0, 0), verbose=False)
    with pytest.raises(ValueError, match="overlap"):
        crop_3D_data_with_overlap(np.zeros((8, 8, 8, 1)), (4, 4, 4, 1), overlap=(1, 0, 0), verbose=False)
    with pytest.raises(AssertionError):
        merge_3D_data_with_overlap(np.zeros((4, 4, 4, 1)), (4, 4, 4, 1), verbose=False)


def test_chunk_planner_matches_golden_and_oracle():
    """Host-side by-chunks bookkeeping (b200_chunk_grid_plan / b200_chunk_patch_coords) vs the reference fixtures."""
    import zlib
    from oracle import port_chunks
    cases = json.load(open(os.path.join(GOLDEN, "chunks.json")))
    L3 = C.c_int64 * 3
    for case in cases:
        g = _lib.ChunkGrid()
        st = _lib.lib().b200_chunk_grid_plan(L3(*case["shape"][:3]), L3(*case["crop"][:3]), L3(*case["padding"]), case["z_start"],
                                             case["z_end"], C.byref(g))
        assert st == 0
        assert list(g.step) == case["steps"] and list(g.vols) == case["vols"]
        assert [g.z_vol_start, g.z_vol_end] == case["z_vol"] and g.total == case["total_vols"]
        rows = np.zeros((g.total, 21), dtype=np.int64)
        buf = (C.c_int64 * 27)()
        for vid in range(g.total):
            assert _lib.lib().b200_chunk_patch_coords(C.byref(g), vid, buf) == 0
            r = list(buf)
            rows[vid] = r[:15] + r[21:27]
        assert zlib.crc32(rows.tobytes()) == case["coords_crc"]
        assert _lib.lib().b200_chunk_patch_coords(C.byref(g), g.total, buf) != 0
    # randomised against the oracle (raw np.pad amounts included)
    rng = np.random.default_rng(1)
    for _ in range(300):
        crop = [int(rng.integers(4, 40)) for _ in range(3)]
        pad = [int(rng.integers(0, c // 2)) for c in crop]
        dim = [int(rng.integers(c, 4 * c + 5)) for c in crop]
        og = port_chunks.ChunkGrid(dim, crop, pad)
        g = _lib.ChunkGrid()
        assert _lib.lib().b200_chunk_grid_plan(L3(*dim), L3(*crop), L3(*pad), -1, -1, C.byref(g)) == 0
        assert g.total == og.total_vols
        buf = (C.c_int64 * 27)()
        for vid in rng.integers(0, g.total, size=min(8, g.total)).tolist():
            _lib.lib().b200_chunk_patch_coords(C.byref(g), vid, buf)
            z, y, x, ext, real = og.patch_coords(vid)
            raw, info = og.pad_to_add((z, y, x), ext)
            assert list(buf) == [z, y, x] + ext + real + [v for ax in raw for v in ax] + [v for ax in info for v in ax]

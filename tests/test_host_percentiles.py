"""Host half of the percentile clipping (``biapy_b200/data/norm.py``): given the digit histograms the device kernel
(``b200_select_hist``; its source runs under tests/test_simt_emulation.py) returns, the radix-select driver and the final
scalar arithmetic must reproduce ``np.percentile`` (numpy images) and ``kthvalue`` (tensors, the reference's
``torch_percentile``, norm.py:475-497) bit for bit.  The histogram here is a numpy stand-in with the kernel's key mapping."""
import numpy as np
import pytest
import torch

from biapy_b200.data import norm as N


def numpy_hist_fn(vals: np.ndarray):
    v = np.ascontiguousarray(vals)
    if v.dtype == np.float32:
        b = v.view(np.uint32)
        keys = np.where(b & 0x80000000, ~b, b | 0x80000000).astype(np.uint32)
    else:
        keys = v.astype(np.uint32)
    calls = []

    def h(shift, bits, prefix, has):
        calls.append((shift, bits, prefix, has))
        k = keys[(keys >> np.uint32(shift + bits)) == prefix] if has else keys
        return np.bincount(((k >> np.uint32(shift)) & np.uint32((1 << bits) - 1)).astype(np.int64), minlength=1 << bits).astype(np.uint32)
    h.calls = calls
    return h


QS = [0.1, 1, 2.5, 50, 75.3, 99, 99.8, 99.99]


@pytest.mark.parametrize("n", [2, 3, 17, 1000, 4097, 250_000, 3_000_000])
@pytest.mark.parametrize("kind", ["f32", "u8", "u16"])
def test_percentiles_match_numpy_and_kthvalue(n, kind):
    rng = np.random.default_rng(n)
    if kind == "f32":
        raw, dt = (rng.standard_normal(n) * 50).astype(np.float32), torch.float32
    elif kind == "u8":
        raw, dt = rng.integers(0, 256, n).astype(np.uint8), torch.uint8
    else:
        raw, dt = rng.integers(0, 60000, n).astype(np.uint16), torch.uint16
    a = raw.astype(np.float32)                      # the reference casts integer images to float32 first (norm.py:171-176)
    got = N.channel_percentiles(numpy_hist_fn(raw), N._SELECT_PLAN[dt], dt, n, QS, torch_rule=False)
    assert got == [float(np.percentile(a, q)) for q in QS]
    got = N.channel_percentiles(numpy_hist_fn(raw), N._SELECT_PLAN[dt], dt, n, QS, torch_rule=True)
    t = torch.from_numpy(a)
    assert got == [t.kthvalue(1 + round(0.01 * float(q) * (n - 1))).values.item() for q in QS]


def test_float32_virtual_index_rounding_of_large_channels():
    """Beyond 2^24 voxels numpy's float32 virtual index (n - 1) * q is no longer exact; the mirror keeps numpy's arithmetic."""
    n = 20_000_001
    raw = (np.random.default_rng(1).standard_normal(n) * 7).astype(np.float32)
    got = N.channel_percentiles(numpy_hist_fn(raw), N._SELECT_PLAN[torch.float32], torch.float32, n, [0.5, 99.5], torch_rule=False)
    assert got == [float(np.percentile(raw, 0.5)), float(np.percentile(raw, 99.5))]


def test_select_shares_passes_between_neighbouring_ranks_and_handles_specials():
    vals = np.array([3.5, -0.0, 0.0, -2.0, 1e-40, -1e-40, np.inf, -np.inf, 3.5, 7.25], dtype=np.float32)
    h = numpy_hist_fn(vals)
    keys = N.select_keys(h, N._SELECT_PLAN[torch.float32], list(range(len(vals))))
    order = [N._key_to_value(keys[r], torch.float32) for r in range(len(vals))]
    want = np.sort(vals)
    assert np.array_equal(np.array(order).view(np.uint32), want.view(np.uint32)) or np.array_equal(np.array(order), want)
    # two adjacent ranks inside one duplicate run cost one pass per digit, not two
    h2 = numpy_hist_fn(np.full(1000, 2.0, np.float32))
    N.select_keys(h2, N._SELECT_PLAN[torch.float32], [400, 401])
    assert len(h2.calls) == 3
    with pytest.raises(AssertionError):
        N.select_keys(numpy_hist_fn(vals), N._SELECT_PLAN[torch.float32], [len(vals)])
    with pytest.raises(NotImplementedError):
        N.channel_percentiles(numpy_hist_fn(vals), N._SELECT_PLAN[torch.float32], torch.float32, 2 ** 32, [50.0], torch_rule=False)


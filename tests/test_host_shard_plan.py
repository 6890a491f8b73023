"""Slab-exchange plan of the sharded sliding-window inference (biapy_b200/engine/dist.py), all ranks simulated in one process:
for several world sizes every rank's output slab, merged by the CPU oracle from the pieces the plan delivers, equals the
single-process merge bit for bit; and the traffic on BASELINE config[2]'s grid is what DESIGN.md states."""
import numpy as np
import pytest

from biapy_b200.data import _stitch
from biapy_b200.engine import dist as bd
from oracle import port_stitch


def _simulate(vshape, patch, ov, pad, world):
    axes = [_stitch.Axis(vshape[i], patch[i], pad[i], ov[i]) for i in range(3)]
    starts_m = [a.starts(1) for a in axes]
    n_yx = len(starts_m[1]) * len(starts_m[2])
    n = len(starts_m[0]) * n_yx
    truth = np.random.default_rng(1).standard_normal((n,) + tuple(patch) + (1,)).astype(np.float32)
    ref = port_stitch.merge_3d(truth, vshape, ov, pad)
    plan = bd.plan_slab_exchange(starts_m[0], n_yx, axes[0].core, pad[0], vshape[0], world)
    arrays = []
    for r in range(world):
        a = np.full_like(truth, np.nan)
        lo, hi = bd.deal_patch_range(n, r, world)
        a[lo:hi] = truth[lo:hi]
        arrays.append(a)
    for c, src, dst, a0, a1 in plan:
        assert bd.deal_patch_range(n, src, world)[0] <= c < bd.deal_patch_range(n, src, world)[1] and src != dst
        arrays[dst][c, a0:a1] = arrays[src][c, a0:a1]
    for r in range(world):
        z0, z1 = bd.slab_range(vshape[0], r, world)
        if z1 == z0:
            continue
        with np.errstate(invalid="ignore"):
            merged = port_stitch.merge_3d(arrays[r], vshape, ov, pad)
        assert np.array_equal(merged[z0:z1], ref[z0:z1]), (world, r)
    return plan, n


@pytest.mark.parametrize("world", [2, 3, 5, 8])
@pytest.mark.parametrize("pad", [(0, 0, 0), (2, 1, 0)])
def test_slab_exchange_delivers_every_piece(world, pad):
    _simulate((37, 18, 20, 1), (16, 12, 12), (0.25, 0.25, 0.25), pad, world)


def test_more_ranks_than_planes_and_no_overlap():
    _simulate((6, 12, 12, 1), (4, 8, 8), (0.0, 0.5, 0.0), (0, 0, 0), 8)


def test_cfg2_grid_traffic():
    """512^3 volume, 128^3 patches, 25 % overlap on 8 ranks: 216 patches, 27 per rank; a rank receives well under a quarter of what
    an all-gather of the predictions would deliver (7/8 of 216 patches)."""
    axes = [_stitch.Axis(512, 128, 0, 0.25) for _ in range(3)]
    sz = axes[0].starts(1)
    assert len(sz) == 6
    plan = bd.plan_slab_exchange(sz, 36, 128, 0, 512, 8)
    plane = 128 * 128 * 4
    recv = [sum((a1 - a0) * plane for c, s, d, a0, a1 in plan if d == r) for r in range(8)]
    allgather = 189 * 128 * plane
    assert max(recv) < 0.25 * allgather
    print("bytes received per rank (fp32 predictions):", recv, "all-gather:", allgather)


def test_allreduce_split_choice():
    """Trainer.enable_cuda_graph (world > 1): where the captured pass is cut so that the tail of the flat gradient can be reduced under
    the rest of backward.  Offsets in flat order = forward order; backward touches them from the last step down."""
    from biapy_b200.engine.train import choose_allreduce_split
    # ten parameters of growing size, parameter k first touched by backward step 10 * k (encoder first in the buffer, last in backward)
    sizes = [10, 20, 40, 80, 1000, 2000, 4000, 8000, 16000, 32000]
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += s
    first = {off: 10 * k + 5 for k, off in enumerate(offs)}
    total, n_steps = o, 100
    m, off = choose_allreduce_split(first, total, n_steps)
    assert off == offs[6] and off <= total // 20          # 3150 of 63150 elements stay in front (the largest boundary within 5 %)
    assert m == 65                                        # everything at / behind `off` is final once the steps >= 65 have run
    assert all(i >= m for o_, i in first.items() if o_ >= off)
    # nothing to hide the collective behind: the split would leave fewer than a tenth of the steps
    assert choose_allreduce_split({0: 0, 10: 2, 600: 50}, 1000, 100) is None
    # no parameter boundary inside the first 5 %
    assert choose_allreduce_split({0: 3, 600: 40}, 1000, 100) is None

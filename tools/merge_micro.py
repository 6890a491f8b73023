"""Times the spline overlap-add of BASELINE config[2]'s grid (216 patches of 128^3 into 512^3) for the kernel variant selected by
B200_MERGE_KERNEL / B200_MERGE_ROWS (read once per process: run one process per variant).  Prints one JSON line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from biapy_b200.data import _stitch  # noqa: E402

V, P = 512, 128
axes = [_stitch.Axis(V, P, 0, 0.25) for _ in range(3)]
starts, wins = [a.starts(1) for a in axes], [a.window() for a in axes]
out = {"kernel": os.environ.get("B200_MERGE_KERNEL", "slot"), "rows": os.environ.get("B200_MERGE_ROWS", "2")}
for name, dt in (("fp32", torch.float32), ("fp16", torch.float16)):
    pred = torch.rand(216, P, P, P, 1, device="cuda").to(dt)
    run = lambda: _stitch.merge_device(pred, (V, V, V), starts, wins, (0, 0, 0), out_dtype=torch.float32)
    ref = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    alg = 216 * P ** 3 * pred.element_size() + V ** 3 * 4
    out[name] = {"ms": round(ms, 3), "GB/s": round(alg / ms / 1e6, 1), "checksum": float(ref.double().sum())}
    del pred, ref
print(json.dumps(out))

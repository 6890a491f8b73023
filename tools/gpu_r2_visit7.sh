#!/bin/bash
set -x
mkdir -p gpurun_out
B200_TAPE_DEBUG=1 python tools/tape_debug.py 2>&1 | grep "\[tape\]" | grep -v "norm bwd" | cut -c1-120
for m in 0 1; do
B200_DBIAS_ANALYTIC=$m timeout 600 python bench.py --detail --no-cpu-baseline --no-infer --no-other-dtype > gpurun_out/bench_v7_$m.json 2> gpurun_out/bench_v7_$m.err
done
python - <<'PY'
import json
res = {}
for m in (0, 1):
    d = json.loads([l for l in open(f"gpurun_out/bench_v7_{m}.json") if l.startswith("{")][-1])
    print(m, "ms", round(d["ms_per_step"], 3))
    res[m] = d["roofline"]["all"]
keys = sorted(set(res[0]) | set(res[1]))
tot = [0, 0]
for k in keys:
    a = res[0].get(k, {}).get("ms_per_step", 0); b = res[1].get(k, {}).get("ms_per_step", 0)
    tot[0] += a; tot[1] += b
    if abs(a - b) > 0.008:
        print(f"  {k:70s} {a:.3f} -> {b:.3f}")
print("sum of events", tot)
PY
